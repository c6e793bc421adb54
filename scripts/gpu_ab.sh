#!/bin/bash
# GPU tests, then the bench with programmatic dependent launch on / off (A/B); usage: scripts/gpu_ab.sh <tag>
TAG=${1:-ab}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest_exit=$?" >> gpurun_out/pytest_gpu_$TAG.log
grep -E "passed|failed|Error|over tolerance|assert|pytest_exit" gpurun_out/pytest_gpu_$TAG.log | cut -c1-220 | head
for PDL in 1 0; do
  UNETB200_PDL=$PDL timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/infer_launches_${TAG}_pdl$PDL.csv --train-profile-out gpurun_out/train_launches_${TAG}_pdl$PDL.csv > gpurun_out/bench_${TAG}_pdl$PDL.log 2>&1
  echo "PDL=$PDL rc=$?"
  grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus"' gpurun_out/bench_${TAG}_pdl$PDL.log
  grep -o '"e2e": {"value": [0-9.]*' gpurun_out/bench_${TAG}_pdl$PDL.log | head -1
  grep -o '"train": {"metric": "images_per_sec_train_512", "value": [0-9.]*, "unit": "images/s", "ms_per_step": [0-9.]*' gpurun_out/bench_${TAG}_pdl$PDL.log
  grep -o '"phases": {[^}]*}' gpurun_out/bench_${TAG}_pdl$PDL.log
done
