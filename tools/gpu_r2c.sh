#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -m gpu -q -rf -p no:cacheprovider --timeout 600"
echo "=== conv kernels (old + fused)"; timeout 900 $P tests/test_kernels_gpu.py -k "conv" > gpurun_out/t_conv.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/t_conv.log
echo "=== bits"; timeout 300 python tools/dbg_gn_bits.py 2>&1 | tail -20
echo "=== timing: halo 64->64 @32, B=512"
for d in 0 1 4 8; do LDM_EPI_DEBUG=$d timeout 120 python tools/bench_conv_gn.py 512 64 64 32 3 8 1 1 0 2>&1 | tail -1; done
echo "=== halo nvar=2 B=256"; timeout 120 python tools/bench_conv_gn.py 256 64 64 32 3 8 1 1 0 2 2>&1 | tail -1
echo "=== halo 128->64"; timeout 120 python tools/bench_conv_gn.py 512 128 64 32 3 8 1 1 0 2>&1 | tail -1
echo "=== to_out 128->64 @32 (1x1, G=1, residual)"
for d in 0 1 4 8; do LDM_EPI_DEBUG=$d timeout 120 python tools/bench_conv_gn.py 512 128 64 32 1 1 0 0 1 2>&1 | tail -1; done
echo "=== to_out 128->128 @16"; timeout 120 python tools/bench_conv_gn.py 512 128 128 16 1 1 0 0 1 2>&1 | tail -1
echo "=== to_out 128->256 @8"; timeout 120 python tools/bench_conv_gn.py 512 128 256 8 1 1 0 0 1 2>&1 | tail -1
echo "=== to_out 128->512 @4"; timeout 120 python tools/bench_conv_gn.py 512 128 512 4 1 1 0 0 1 2>&1 | tail -1
echo "=== conv1 64->128 @16"; timeout 120 python tools/bench_conv_gn.py 512 64 128 16 3 8 1 1 0 2>&1 | tail -1
echo "=== conv1 128->256 @8"; timeout 120 python tools/bench_conv_gn.py 512 128 256 8 3 8 1 1 0 2>&1 | tail -1
echo "=== conv1 256->512 @4"; timeout 120 python tools/bench_conv_gn.py 512 256 512 4 3 8 1 1 0 2>&1 | tail -1
echo "=== unet"; timeout 900 $P tests/test_unet_gpu.py > gpurun_out/t_unet.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_unet.log
echo "=== sampler"; timeout 1500 $P tests/test_sampler_gpu.py > gpurun_out/t_sampler.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/t_sampler.log
echo "=== profile"; timeout 300 python tools/profile_pass.py 512 > gpurun_out/prof.txt 2> gpurun_out/prof_events.txt; echo "rc=$?"; cat gpurun_out/prof.txt
echo "=== bench"; timeout 900 python bench.py --steps 1 --warmup 3 --no-train --no-cpu-baseline --no-variants > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], {k:round(v['ms'],3) for k,v in d['roofline']['families'].items()})"
