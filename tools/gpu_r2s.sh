#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:linattn_tc2 -s 2 -c 1 -o gpurun_out/la2 -f python tools/run_linattn.py 512 32 > gpurun_out/ncu_la2.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_la2.log; ls -la gpurun_out/la2.ncu-rep
