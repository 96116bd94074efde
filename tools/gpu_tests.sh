#!/bin/bash
# GPU box: run the parity suites in separate processes (a trapped kernel poisons its CUDA context only).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
P="python -m pytest -m gpu -q -rf -p no:cacheprovider --timeout 600"
echo "=== fp32 path"; timeout 1200 $P tests -k "not tcgen05 and not bf16" > gpurun_out/t_fp32.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_fp32.log
echo "=== tcgen05 conv kernels"; timeout 600 $P tests/test_kernels_gpu.py -k "tcgen05 or (conv_transpose and bf16)" > gpurun_out/t_tc.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_tc.log
echo "=== bf16 path"; timeout 1200 $P tests -k "bf16 and not tcgen05 and not conv_transpose" > gpurun_out/t_bf16.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_bf16.log
