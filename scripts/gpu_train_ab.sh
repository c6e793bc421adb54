#!/bin/bash
# train tests (fail fast) + bench with train launch profile; usage: scripts/gpu_train_ab.sh <tag>
TAG=${1:-p1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_local.py tests/test_gpu_train.py tests/test_gpu_native.py -m gpu -q -x > gpurun_out/pytest_train_$TAG.log 2>&1
grep -E "passed|failed|Error|over tolerance|assert" gpurun_out/pytest_train_$TAG.log | cut -c1-220 | head
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --train-profile-out gpurun_out/train_launches_$TAG.csv > gpurun_out/bench_$TAG.log 2>&1
grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus"' gpurun_out/bench_$TAG.log
grep -o '"train": {"metric": "images_per_sec_train_512", "value": [0-9.]*, "unit": "images/s", "ms_per_step": [0-9.]*' gpurun_out/bench_$TAG.log
grep -o '"phases": {[^}]*}' gpurun_out/bench_$TAG.log
