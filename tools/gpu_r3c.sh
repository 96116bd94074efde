#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -m gpu -q -rf -p no:cacheprovider --timeout 300"
timeout 1500 $P tests/test_trainer_gpu.py tests/test_train_gpu.py tests/test_backward_gpu.py tests/test_unet_gpu.py > gpurun_out/t_train.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_train.log
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants --n-steps 50 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['train']['ms_per_step'], d['train_batch256']['ms_per_step'], d['train']['cuda_graphs'])"
