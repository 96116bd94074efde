#!/bin/bash
# GPU box: parity suites (separate processes), then a short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
P="python -m pytest -m gpu -q -rf -p no:cacheprovider --timeout 900 -x"
echo "=== kernels"; timeout 900 $P tests/test_kernels_gpu.py > gpurun_out/t_kernels.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_kernels.log
echo "=== unet"; timeout 900 $P tests/test_unet_gpu.py > gpurun_out/t_unet.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/t_unet.log
echo "=== sampler"; timeout 1500 $P tests/test_sampler_gpu.py > gpurun_out/t_sampler.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/t_sampler.log
echo "=== bench"; timeout 1200 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
