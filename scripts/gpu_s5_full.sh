#!/bin/bash
# session 5: infer + train tests (fail fast), bench with per-launch tables; usage: scripts/gpu_s5_full.sh <tag>
TAG=${1:-s5b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_infer.py tests/test_gpu_train_local.py tests/test_gpu_train.py -m gpu -q -x > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest_exit=$?" >> gpurun_out/pytest_$TAG.log
grep -E "passed|failed|Error|over tolerance|assert|pytest_exit" gpurun_out/pytest_$TAG.log | cut -c1-220 | head
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/infer_launches_$TAG.csv --train-profile-out gpurun_out/train_launches_$TAG.csv > gpurun_out/bench_$TAG.log 2>&1
grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus"' gpurun_out/bench_$TAG.log
grep -o '"e2e": {"value": [0-9.]*' gpurun_out/bench_$TAG.log | head -1
grep -o '"frac": [0-9.]*' gpurun_out/bench_$TAG.log | head -1
grep -o '"train": {"metric": "images_per_sec_train_512", "value": [0-9.]*, "unit": "images/s", "ms_per_step": [0-9.]*' gpurun_out/bench_$TAG.log
grep -o '"phases": {[^}]*}' gpurun_out/bench_$TAG.log
