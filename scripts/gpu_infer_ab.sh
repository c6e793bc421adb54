#!/bin/bash
# inference tests + bench A/B on an env switch; usage: scripts/gpu_infer_ab.sh <tag> <ENVVAR>
TAG=${1:-i1}; VAR=${2:-UNETB200_NO_WPREFETCH}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_infer.py -m gpu -x -q > gpurun_out/pytest_infer_$TAG.log 2>&1; echo "pytest_exit=$?" >> gpurun_out/pytest_infer_$TAG.log
grep -E "passed|failed|Error|assert|pytest_exit" gpurun_out/pytest_infer_$TAG.log | cut -c1-200 | head
for V in "" 1 "" 1; do
  if [ -n "$V" ]; then export $VAR=1; else unset $VAR; fi
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/bench_infer_${TAG}_$V.log 2>&1
  echo "$VAR=$V $(grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus"' gpurun_out/bench_infer_${TAG}_$V.log) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/bench_infer_${TAG}_$V.log | head -1)"
done
