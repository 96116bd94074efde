#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -m gpu -q -rf -p no:cacheprovider --timeout 300"
echo "=== unet + kernels"; timeout 1200 $P tests/test_unet_gpu.py tests/test_kernels_gpu.py tests/test_train_gpu.py > gpurun_out/t_unet.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_unet.log
echo "=== errors vs oracle (bf16)"; timeout 300 python tools/diag_unet.py 2>&1 | tail -6
echo "=== v2 32x32"; timeout 300 python tools/prof_linattn.py 512 32 2>&1 | grep -i "linattn\|fold"
