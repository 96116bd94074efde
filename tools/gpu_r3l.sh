#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_train_gpu.py tests/test_trainer_gpu.py -m gpu -q -p no:cacheprovider --timeout 800 > gpurun_out/t_mg.log 2>&1; echo rc=$?; tail -3 gpurun_out/t_mg.log
for sp in 2 1 4; do
echo "=== 2-GPU train, LDM_TRAIN_SPLIT=$sp"
LDM_TRAIN_SPLIT=$sp timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 --no-cpu-baseline --no-variants --n-steps 20 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['train']['ms_per_step'], d['train_batch256']['ms_per_step'])"
done
echo "=== 1-GPU"; timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants --n-steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['train']['ms_per_step'], d['train_batch256']['ms_per_step'])"
