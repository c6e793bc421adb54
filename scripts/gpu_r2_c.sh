#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_local.py -m gpu -q -x -p no:cacheprovider -k "not 1024" > gpurun_out/pytest_r2c.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2c.log
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-library-baseline --steps 20 > gpurun_out/bench_r2c_$tag.json 2> gpurun_out/bench_r2c_$tag.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_r2c_$tag.json'))
t=d['train']; print('$tag', 'infer', round(d['value']), 'train ms', round(t['ms_per_step'],3), {k:round(v,3) for k,v in t['phases'].items()}, 'e2e', round(t['e2e']['value']))
"
}
run aux A=1
run aux_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
run aux_nopack UNETB200_NO_AUX_PACK=1
run aux_nopack_conn32 UNETB200_NO_AUX_PACK=1 CUDA_DEVICE_MAX_CONNECTIONS=32
run aux_nofuse UNETB200_NO_BN_FUSE=1
run noaux UNETB200_NO_AUX=1
timeout 900 python scripts/explore_recipe.py > gpurun_out/explore_recipe_r2c.log 2>&1
echo "explore rc=$?"; grep -E "==|step |   [0-9.]* s" gpurun_out/explore_recipe_r2c.log | tail -50
EXPLORE_BATCH=8 timeout 900 python scripts/explore_recipe.py > gpurun_out/explore_recipe_r2c_b8.log 2>&1
echo "explore b8 rc=$?"; grep -E "==|step |   [0-9.]* s" gpurun_out/explore_recipe_r2c_b8.log | tail -50
