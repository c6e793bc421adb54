#!/bin/bash
TAG=${1:-i1}
mkdir -p gpurun_out
make >/dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_infer.py -m gpu -x -q -s > gpurun_out/pytest_infer_$TAG.log 2>&1; echo "pytest_exit=$?" >> gpurun_out/pytest_infer_$TAG.log
grep -E "^\[|passed|failed|Error|assert|pytest_exit" gpurun_out/pytest_infer_$TAG.log | cut -c1-330 | head -40
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train --profile-out gpurun_out/infer_launches_$TAG.csv > gpurun_out/bench_infer_$TAG.log 2>&1
tail -c 1500 gpurun_out/bench_infer_$TAG.log
