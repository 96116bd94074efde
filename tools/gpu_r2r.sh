#!/bin/bash
mkdir -p gpurun_out
echo "=== v2 32x32"; timeout 300 python tools/prof_linattn.py 512 32 2>&1 | tail -8
echo "=== v2 16x16"; timeout 300 python tools/prof_linattn.py 512 16 2>&1 | tail -8
echo "=== v1 32x32"; LDM_LINATTN_V1=1 timeout 300 python tools/prof_linattn.py 512 32 2>&1 | tail -8
