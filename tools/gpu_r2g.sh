#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -m gpu -q -rf -p no:cacheprovider --timeout 600"
echo "=== unet"; timeout 900 $P tests/test_unet_gpu.py tests/test_autoencoder_gpu.py tests/test_trainer_gpu.py > gpurun_out/t_unet.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_unet.log
for m in 0 4 8 64 1 32 77; do echo "=== skip mask $m"; LDM_SKIP_FAM=$m timeout 300 python bench.py --steps 1 --warmup 2 --no-train --no-cpu-baseline --no-variants --n-steps 400 2> gpurun_out/bench.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms/timestep', d['ms_per_step']/400)"; done
