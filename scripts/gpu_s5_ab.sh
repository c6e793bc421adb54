#!/bin/bash
# session 5: train tests (fail fast) + bench A/B of the fused BN finalize; usage: scripts/gpu_s5_ab.sh <tag>
TAG=${1:-s5a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_local.py tests/test_gpu_train.py -m gpu -q -x > gpurun_out/pytest_train_$TAG.log 2>&1
grep -E "passed|failed|Error|over tolerance|assert" gpurun_out/pytest_train_$TAG.log | cut -c1-220 | head
for V in tail notail; do
  if [ $V = notail ]; then export UNETB200_NO_BN_TAIL=1; fi
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --train-profile-out gpurun_out/train_launches_${TAG}_$V.csv > gpurun_out/bench_${TAG}_$V.log 2>&1
  echo "== $V"
  grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus"' gpurun_out/bench_${TAG}_$V.log
  grep -o '"train": {"metric": "images_per_sec_train_512", "value": [0-9.]*, "unit": "images/s", "ms_per_step": [0-9.]*' gpurun_out/bench_${TAG}_$V.log
  grep -o '"phases": {[^}]*}' gpurun_out/bench_${TAG}_$V.log
done
