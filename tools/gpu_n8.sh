#!/bin/bash
# 8 GPUs of one box: the bench line (sampling + both training legs) under torchrun, as the driver launches it
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 1 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "rc=$?"; tail -2 gpurun_out/r02_bench_n8.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['train']['ms_per_step'], d['train']['value'], d['train_batch256']['ms_per_step'], d['train_batch256']['value'])"
