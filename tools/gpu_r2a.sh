#!/bin/bash
# GPU box, round 2 step A: fused GroupNorm epilogue -- kernel parity first, then the UNet / sampler suites, a warm per-launch
# profile and a short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
P="python -m pytest -m gpu -q -rf -p no:cacheprovider --timeout 600"
echo "=== conv kernels (old + fused)"; timeout 900 $P tests/test_kernels_gpu.py -k "conv" > gpurun_out/t_conv.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/t_conv.log
echo "=== unet"; timeout 900 $P tests/test_unet_gpu.py > gpurun_out/t_unet.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_unet.log
echo "=== sampler"; timeout 1500 $P tests/test_sampler_gpu.py > gpurun_out/t_sampler.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/t_sampler.log
echo "=== profile"; timeout 300 python tools/profile_pass.py 512 > gpurun_out/prof.txt 2> gpurun_out/prof_events.txt; echo "rc=$?"; cat gpurun_out/prof.txt
echo "=== profile (no fuse)"; LDM_NO_GN_FUSE=1 timeout 300 python tools/profile_pass.py 512 > gpurun_out/prof_nofuse.txt 2> gpurun_out/prof_nofuse_events.txt; echo "rc=$?"; cat gpurun_out/prof_nofuse.txt
echo "=== bench"; timeout 900 python bench.py --steps 1 --warmup 3 --no-train --no-cpu-baseline --no-variants > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
