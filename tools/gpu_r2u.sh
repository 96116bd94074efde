#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "linear_attention" -p no:cacheprovider > gpurun_out/t_la.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_la.log
echo "=== v2 32x32"; timeout 300 python tools/prof_linattn.py 512 32 2>&1 | grep linattn
echo "=== v2 16x16"; timeout 300 python tools/prof_linattn.py 512 16 2>&1 | grep linattn
LDM_LA2_TRACE=1 timeout 300 python tools/run_linattn.py 512 32 1 > gpurun_out/la2_trace.txt 2>&1; echo "rc=$?"; grep -c LA2TRACE gpurun_out/la2_trace.txt
