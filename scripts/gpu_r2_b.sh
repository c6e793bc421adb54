#!/bin/bash
# round 2, call B: side-stream wgrads + new head_bwd_weight: parity, A/B timing; training-recipe exploration
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_local.py -m gpu -q -x -p no:cacheprovider -k "not 1024" > gpurun_out/pytest_r2b.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2b.log
for mode in aux noaux; do
  if [ $mode = noaux ]; then export UNETB200_NO_AUX=1; else unset UNETB200_NO_AUX; fi
  timeout 300 python bench.py --no-cpu-baseline --no-library-baseline --steps 20 --train-profile-out gpurun_out/train_launches_r2b_$mode.csv > gpurun_out/bench_r2b_$mode.json 2> gpurun_out/bench_r2b_$mode.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_r2b_$mode.json'))
t=d['train']; print('$mode', 'infer', round(d['value']), 'train ms', round(t['ms_per_step'],3), t['phases'], 'e2e', round(t['e2e']['value']))
"
done
unset UNETB200_NO_AUX
timeout 600 python scripts/explore_recipe.py > gpurun_out/explore_recipe_r2b.log 2>&1
echo "explore rc=$?"; grep -E "==|step  *[0-9]*00 |   [0-9.]* s" gpurun_out/explore_recipe_r2b.log | tail -50
