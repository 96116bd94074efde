#!/bin/bash
echo "--- early release"; timeout 300 python tools/bench_conv.py 512 64 64 32 3 2>&1 | tail -1
echo "--- + MMA waits for the accumulator read (bit 5)"; LDM_HALO_DEBUG=32 timeout 300 python tools/bench_conv.py 512 64 64 32 3 2>&1 | tail -1
timeout 300 python tools/bench_conv.py 512 128 128 16 3 2>&1 | tail -1
P="python -m pytest -m gpu -q -rf -p no:cacheprovider --timeout 300"
timeout 1200 $P tests/test_unet_gpu.py tests/test_kernels_gpu.py 2>&1 | tail -2
timeout 900 python bench.py --steps 1 --warmup 3 --no-train --no-cpu-baseline --no-variants --n-steps 200 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step']/200, {k:round(v['ms'],3) for k,v in d['roofline']['families'].items()})"
