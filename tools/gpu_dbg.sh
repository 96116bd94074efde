#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 600 > gpurun_out/dbg_full.log 2>&1
grep -E "passed|failed" gpurun_out/dbg_full.log | tail -1
grep -n "^E " gpurun_out/dbg_full.log | head -5 | cut -c1-250
