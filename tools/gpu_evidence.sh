#!/bin/bash
# GPU box: the ncu evidence behind profiles/r02_* (outputs kept small: reports are exported to CSV and deleted).
mkdir -p gpurun_out
echo "=== ncu launch list (bench, T=4)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench_T4.csv python bench.py --steps 1 --warmup 1 --n-steps 4 --no-cpu-baseline --no-train --no-variants > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
python tools/launch_summary.py gpurun_out/r02_launches_bench_T4.csv > gpurun_out/r02_launches_bench_T4.txt; head -30 gpurun_out/r02_launches_bench_T4.txt
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"
echo "=== ncu metrics: one UNet pass, conv / attention / GroupNorm kernels"
timeout 1500 ncu --metrics $M --clock-control none -k regex:"conv_halo|conv_tc|linattn_tc|linattn_mma|gn_apply|gn_stats|pool_gn|attn_kernel|initial_conv|final_conv" -c 400 --csv --log-file gpurun_out/r02_ncu_unet_pass_kernels.csv python tools/one_pass.py 512 2 > gpurun_out/ncu_pass.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_pass.log
echo "=== ncu metrics: update kernels"
timeout 600 ncu --metrics $M --clock-control none -k regex:"p_sample|q_sample" -s 6 -c 4 --csv --log-file gpurun_out/r02_ncu_update_kernels.csv python tools/ew_kernels.py 256 > gpurun_out/ncu_ew.log 2>&1; echo "rc=$?"
timeout 120 python tools/ew_kernels.py 256
echo "=== ncu --set full: halo conv (mode 0 / mode 1), attention"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"conv_halo" -s 7 -c 4 -o gpurun_out/tmp_halo -f python tools/one_pass.py 512 2 > gpurun_out/ncu_full1.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/tmp_halo.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_conv_halo.csv 2>/dev/null; mv gpurun_out/tmp_halo.ncu-rep gpurun_out/r02_halo.ncu-rep
timeout 900 ncu --set full --clock-control none -k regex:"linattn_tc2" -s 3 -c 1 -o gpurun_out/tmp_la -f python tools/one_pass.py 512 2 > gpurun_out/ncu_full2.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/tmp_la.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_linattn_tc.csv 2>/dev/null; rm -f gpurun_out/tmp_la.ncu-rep
echo "=== skip-family timing (true in-graph cost per family)"
for m in 0 4 8 64 32 128 1; do LDM_SKIP_FAM=$m LDM_BENCH_ALLOW_NONFINITE=1 timeout 300 python bench.py --steps 1 --warmup 2 --no-train --no-cpu-baseline --no-variants --n-steps 400 2> gpurun_out/bench.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('LDM_SKIP_FAM=$m ms/timestep', round(d['ms_per_step']/400, 4))"; done | tee gpurun_out/r02_skip_family_ms.txt
du -sh gpurun_out
