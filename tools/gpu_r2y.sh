#!/bin/bash
mkdir -p gpurun_out
P="python -m pytest -m gpu -q -rf -p no:cacheprovider --timeout 300"
timeout 1500 $P tests/test_trainer_gpu.py tests/test_train_gpu.py > gpurun_out/t_train.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_train.log
