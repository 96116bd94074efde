#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -m gpu -q -rf -p no:cacheprovider --timeout 600"
echo "=== kernels"; timeout 900 $P tests/test_kernels_gpu.py > gpurun_out/t_kernels.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_kernels.log
echo "=== unet"; timeout 900 $P tests/test_unet_gpu.py > gpurun_out/t_unet.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_unet.log
echo "=== sampler"; timeout 1500 $P tests/test_sampler_gpu.py > gpurun_out/t_sampler.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_sampler.log
echo "=== profile"; timeout 300 python tools/profile_pass.py 512 > gpurun_out/prof.txt 2> gpurun_out/prof_events.txt; echo "rc=$?"; cat gpurun_out/prof.txt
echo "=== bench default"; timeout 900 python bench.py --steps 1 --warmup 3 --no-train --no-cpu-baseline --no-variants > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], {k:round(v['ms'],3) for k,v in d['roofline']['families'].items()})"
echo "=== rest of the suites"; timeout 1500 $P tests/test_autoencoder_gpu.py tests/test_backward_gpu.py tests/test_train_gpu.py tests/test_trainer_gpu.py > gpurun_out/t_rest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_rest.log
