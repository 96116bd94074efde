#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s -p no:cacheprovider --timeout 800 > gpurun_out/r02_two_gpu_test.log 2>&1; echo rc=$?; tail -6 gpurun_out/r02_two_gpu_test.log
echo "=== 2-GPU bench (train legs)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 --no-cpu-baseline --no-variants --n-steps 50 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; tail -3 gpurun_out/bench_n2.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['train']['ms_per_step'], d['train_batch256']['ms_per_step'])"
echo "=== 1-GPU train for comparison"
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants --n-steps 50 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['train']['ms_per_step'], d['train_batch256']['ms_per_step'])"
