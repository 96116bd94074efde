#!/bin/bash
mkdir -p gpurun_out
echo "=== full bench"; SECONDS=0; timeout 1200 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$? elapsed ${SECONDS}s"; grep -E "bench:" gpurun_out/bench_full.err | tail -5; python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); print(d['value'], d['e2e']['value']); print(json.dumps(d['variants'], indent=1)); print(json.dumps(d['gpu_library_baseline'], indent=1)); print(d['train']); print(d['train_batch256']); print(d['cpu_baseline'])"
