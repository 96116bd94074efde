#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -m gpu -q -rf -p no:cacheprovider --timeout 300"
echo "=== linattn"; timeout 600 $P tests/test_kernels_gpu.py -k "linear_attention_prenorm" > gpurun_out/t_la.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_la.log
echo "=== timing"; timeout 120 python tools/bench_linattn.py 512 32; timeout 120 python tools/bench_linattn.py 512 16
echo "=== unet"; timeout 900 $P tests/test_unet_gpu.py tests/test_sampler_gpu.py > gpurun_out/t_unet.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_unet.log
echo "=== bench"; timeout 900 python bench.py --steps 1 --warmup 3 --no-train --no-cpu-baseline --no-variants > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], {k:round(v['ms'],3) for k,v in d['roofline']['families'].items()})"
