#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -m gpu -q -rf -p no:cacheprovider --timeout 900 -s"
echo "=== train"; timeout 1500 $P tests/test_train_gpu.py > gpurun_out/t_train.log 2>&1; echo "rc=$?"; grep -E "worst|graphed|passed|failed|Error" gpurun_out/t_train.log | tail -20
echo "=== sampler"; timeout 1500 $P tests/test_sampler_gpu.py > gpurun_out/t_sampler.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_sampler.log
echo "=== full bench"; /usr/bin/time -v timeout 1200 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?"; grep -E "Elapsed|bench:" gpurun_out/bench_full.err | tail -5; python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); print(d['value'], d['e2e']['value']); print(json.dumps(d['variants'], indent=1)); print(json.dumps(d['gpu_library_baseline'], indent=1)); print(d['train']); print(d['train_batch256']); print(d['cpu_baseline'])"
