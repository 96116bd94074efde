#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -m gpu -q -rf -p no:cacheprovider --timeout 600"
echo "=== conv kernels (old + fused)"; timeout 900 $P tests/test_kernels_gpu.py -k "conv" > gpurun_out/t_conv.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_conv.log
echo "=== bits"; timeout 300 python tools/dbg_gn_bits.py 2>&1 | grep -v "differing 0\|equal True stats equal True" | tail
echo "=== timing"
timeout 120 python tools/bench_conv_gn.py 512 64 64 32 3 8 1 1 0 2>&1 | tail -1
CONV_RES=1 timeout 120 python tools/bench_conv_gn.py 512 64 64 32 3 1 0 0 0 2>&1 | tail -1
timeout 120 python tools/bench_conv_gn.py 512 128 64 32 3 8 1 1 0 2>&1 | tail -1
timeout 120 python tools/bench_conv_gn.py 512 128 64 32 1 1 0 0 1 2>&1 | tail -1
CONV_RES=1 timeout 120 python tools/bench_conv_gn.py 512 128 128 16 3 1 0 0 0 2>&1 | tail -1
echo "=== unet"; timeout 900 $P tests/test_unet_gpu.py > gpurun_out/t_unet.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_unet.log
echo "=== sampler"; timeout 1500 $P tests/test_sampler_gpu.py > gpurun_out/t_sampler.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_sampler.log
echo "=== profile"; timeout 300 python tools/profile_pass.py 512 > gpurun_out/prof.txt 2> gpurun_out/prof_events.txt; echo "rc=$?"; cat gpurun_out/prof.txt
echo "=== profile nofuse"; LDM_NO_GN_FUSE=1 timeout 300 python tools/profile_pass.py 512 > gpurun_out/prof_nofuse.txt 2> gpurun_out/prof_nofuse_events.txt; echo "rc=$?"; cat gpurun_out/prof_nofuse.txt
echo "=== bench nofuse"; LDM_NO_GN_FUSE=1 timeout 900 python bench.py --steps 1 --warmup 3 --no-train --no-cpu-baseline --no-variants > gpurun_out/bench_nofuse.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_nofuse.json')); print(d['value'], d['e2e']['value'], {k:round(v['ms'],3) for k,v in d['roofline']['families'].items()})"
echo "=== bench default"; timeout 900 python bench.py --steps 1 --warmup 3 --no-train --no-cpu-baseline --no-variants > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], {k:round(v['ms'],3) for k,v in d['roofline']['families'].items()})"
