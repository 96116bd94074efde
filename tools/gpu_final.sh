#!/bin/bash
# GPU box: the whole -m gpu suite, the evidence pass, then the default bench line
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_pytest_gpu.log
bash tools/gpu_evidence.sh
echo "=== default bench"; timeout 1500 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "rc=$?"; tail -2 gpurun_out/r02_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline'] and d['cpu_baseline']['value']); print(d['variants']); print(d['train'] and d['train']['ms_per_step'], d['train_batch256'] and d['train_batch256']['ms_per_step'])"
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_reference.json 2>/dev/null; echo "rc=$?"; cut -c1-300 gpurun_out/r02_bench_reference.json
