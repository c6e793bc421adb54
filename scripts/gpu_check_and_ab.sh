#!/bin/bash
# Development loop on one GPU: parity tests of the kernels touched most often, then bench.py A/B runs that differ only in
# an environment switch (per-launch CUDA-event tables land in gpurun_out/).  Edit the `run` lines for the switch under test.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_infer.py tests/test_gpu_train.py tests/test_gpu_native.py tests/test_gpu_train_local.py -m gpu -q -x -p no:cacheprovider -k "not 1024" > gpurun_out/pytest_ab.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ab.log
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-library-baseline --steps 20 --profile-out gpurun_out/infer_launches_ab_$tag.csv --train-profile-out gpurun_out/train_launches_ab_$tag.csv > gpurun_out/bench_ab_$tag.json 2> gpurun_out/bench_ab_$tag.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_ab_$tag.json') if l.startswith('{')][-1])
t=d['train']; print('$tag', 'infer', round(d['value']), 'roof', round(d['roofline']['frac'],4), 'train ms', round(t['ms_per_step'],3), {k:round(v,3) for k,v in t['phases'].items()})
"
}
run default A=1
run stemreg UNETB200_STEM_REGSTORE=1
run default2 A=1
