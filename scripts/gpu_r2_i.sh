#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_infer.py tests/test_gpu_train.py tests/test_gpu_native.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_r2i.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2i.log
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-library-baseline --steps 20 --profile-out gpurun_out/infer_launches_r2i_$tag.csv > gpurun_out/bench_r2i_$tag.json 2> gpurun_out/bench_r2i_$tag.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r2i_$tag.json') if l.startswith('{')][-1])
t=d['train']; print('$tag', 'infer', round(d['value']), 'roof', round(d['roofline']['frac'],4), 'train ms', round(t['ms_per_step'],3), {k:round(v,3) for k,v in t['phases'].items()})
"
}
run st256 A=1
run regstore UNETB200_TC_REGSTORE=1
run st256_again A=1
