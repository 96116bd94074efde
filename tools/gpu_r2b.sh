#!/bin/bash
mkdir -p gpurun_out
echo "=== bits"; timeout 300 python tools/dbg_gn_bits.py 2>&1 | tail -20
echo "=== timing knobs: halo 64->64 @32, B=512"
for d in 0 1 2 4 8 15; do LDM_EPI_DEBUG=$d timeout 120 python tools/bench_conv_gn.py 512 64 64 32 3 8 1 1 0 2>&1 | tail -1; done
echo "=== halo nvar=2 B=256"; timeout 120 python tools/bench_conv_gn.py 256 64 64 32 3 8 1 1 0 2 2>&1 | tail -1
echo "=== to_out 128->64 @32 (1x1, G=1, residual)"
for d in 0 1 2 4 8 15; do LDM_EPI_DEBUG=$d timeout 120 python tools/bench_conv_gn.py 512 128 64 32 1 1 0 0 1 2>&1 | tail -1; done
echo "=== to_out 128->128 @16"
for d in 0 1 4 8; do LDM_EPI_DEBUG=$d timeout 120 python tools/bench_conv_gn.py 512 128 128 16 1 1 0 0 1 2>&1 | tail -1; done
echo "=== conv1 64->128 @16"
for d in 0 1 4 8; do LDM_EPI_DEBUG=$d timeout 120 python tools/bench_conv_gn.py 512 64 128 16 3 8 1 1 0 2>&1 | tail -1; done
echo "=== ncu halo fused"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_halo -s 30 -c 1 -o gpurun_out/ncu_halo_gn -f python tools/bench_conv_gn.py 512 64 64 32 3 8 1 1 0 > gpurun_out/ncu_halo_gn.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_halo_gn.log
