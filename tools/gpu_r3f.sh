#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -m gpu -q -rf -p no:cacheprovider --timeout 300"
echo "=== unet"; timeout 1200 $P tests/test_unet_gpu.py > gpurun_out/t_unet.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/t_unet.log
echo "=== profile"; timeout 300 python tools/profile_pass.py 512 > gpurun_out/prof.txt 2> gpurun_out/prof_events.txt; echo "rc=$?"; cat gpurun_out/prof.txt
echo "=== bench"; timeout 900 python bench.py --steps 1 --warmup 3 --no-train --no-cpu-baseline --no-variants --n-steps 200 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['ms_per_step']/200, {k:round(v['ms'],3) for k,v in d['roofline']['families'].items()})"
