#!/bin/bash
mkdir -p gpurun_out
LDM_LA2_TRACE=1 timeout 300 python tools/run_linattn.py 512 32 1 > gpurun_out/la2_trace.txt 2>&1; echo "rc=$?"; grep -c LA2TRACE gpurun_out/la2_trace.txt
