#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_halo -s 7 -c 1 -o gpurun_out/halo_xf -f python tools/one_pass.py 512 2 > gpurun_out/ncu_xf.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_xf.log
