#!/bin/bash
for d in 0 1 2 3; do echo "LDM_LA_DEBUG=$d"; LDM_LA_DEBUG=$d timeout 120 python tools/bench_linattn.py 512 32 | head -1; done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:linattn_tc -s 5 -c 1 -o gpurun_out/ncu_linattn_tc -f python tools/bench_linattn.py 512 32 > gpurun_out/ncu_la.log 2>&1; tail -2 gpurun_out/ncu_la.log
