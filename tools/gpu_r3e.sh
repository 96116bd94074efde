#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -m gpu -q -rf -p no:cacheprovider --timeout 300"
echo "=== unet + sampler + kernels"; timeout 1200 $P tests/test_unet_gpu.py tests/test_sampler_gpu.py tests/test_kernels_gpu.py > gpurun_out/t_unet.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_unet.log
echo "=== errors vs oracle"; timeout 300 python tools/diag_unet.py 2>&1 | grep "bf16 impl=0"
echo "=== profile"; timeout 300 python tools/profile_pass.py 512 > gpurun_out/prof.txt 2> gpurun_out/prof_events.txt; echo "rc=$?"; cat gpurun_out/prof.txt
echo "=== bench"; timeout 900 python bench.py --steps 1 --warmup 3 --no-train --no-cpu-baseline --no-variants > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], {k:round(v['ms'],3) for k,v in d['roofline']['families'].items()})"
echo "=== bench, LDM_NO_CONV_XFORM=1"; LDM_NO_CONV_XFORM=1 timeout 900 python bench.py --steps 1 --warmup 3 --no-train --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], {k:round(v['ms'],3) for k,v in d['roofline']['families'].items()})"
