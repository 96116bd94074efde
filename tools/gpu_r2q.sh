#!/bin/bash
# GPU box: linattn_tc2_kernel first run -- parity tests of the fused block, forced fallback, micro-benchmark
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "linear_attention" -p no:cacheprovider > gpurun_out/t_la.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_la.log
echo "=== forced fallback"
LDM_LA2_FORCE_FALLBACK=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "to_out_fused" -p no:cacheprovider 2>&1 | tail -3
echo "=== v1 only"
LDM_LINATTN_V1=1 timeout 300 python tools/bench_linattn.py 512 32 2>&1 | tail -1
echo "=== v2"
timeout 300 python tools/bench_linattn.py 512 32 2>&1 | tail -1
timeout 300 python tools/bench_linattn.py 512 16 2>&1 | tail -1
