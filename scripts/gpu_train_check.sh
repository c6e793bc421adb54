#!/bin/bash
TAG=${1:-t1}
mkdir -p gpurun_out
make >/dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -s > gpurun_out/pytest_train_$TAG.log 2>&1; echo "pytest_exit=$?" >> gpurun_out/pytest_train_$TAG.log
grep -E "^\[|^    |passed|failed|Error|assert" gpurun_out/pytest_train_$TAG.log | cut -c1-200 | head -120
