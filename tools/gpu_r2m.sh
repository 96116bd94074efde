#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -m gpu -q -rf -p no:cacheprovider --timeout 600"
echo "=== all gpu tests"; timeout 2400 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 900 > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_all.log
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== batch1"; timeout 300 python tools/batch1.py 2>&1 | tail -3
echo "=== bench short"; timeout 900 python bench.py --steps 1 --warmup 3 --no-train --no-cpu-baseline --no-variants > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], {k:round(v['ms'],3) for k,v in d['roofline']['families'].items()})"
echo "=== ncu launch list (bench, T=4)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench_T4.csv python bench.py --steps 1 --warmup 1 --n-steps 4 --no-cpu-baseline --no-train --no-variants > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "=== ncu full: one UNet pass, conv / attention / GroupNorm kernels"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"conv_halo|conv_tc|linattn_tc|linattn_mma|gn_apply|gn_stats" -s 85 -c 85 -o gpurun_out/r02_ncu_full_unet_pass -f python tools/one_pass.py 512 2 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_full.log
echo "=== ncu full: update kernels"
timeout 600 ncu --set full --clock-control none -k regex:"p_sample|q_sample" -s 6 -c 4 -o gpurun_out/r02_ncu_full_update_kernels -f python tools/ew_kernels.py 256 > gpurun_out/ncu_ew.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/ncu_ew.log
timeout 120 python tools/ew_kernels.py 256
