#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_local.py tests/test_gpu_infer.py -m gpu -q -x -p no:cacheprovider -k "not 1024" > gpurun_out/pytest_r2g.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2g.log
timeout 300 python bench.py --no-cpu-baseline --no-library-baseline --steps 20 --train-profile-out gpurun_out/train_launches_r2g.csv > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err
python -c "
import json
d=json.load(open('gpurun_out/bench_r2g.json'))
t=d['train']; print('infer', round(d['value']), 'e2e', round(d['e2e']['value']), 'train ms', round(t['ms_per_step'],3), {k:round(v,3) for k,v in t['phases'].items()}, 'e2e', round(t['e2e']['value']), round(t['e2e']['fp32_frames']['value']))
"
grep -E "pack_dgrad|head_bwd|maxpool_bwd" gpurun_out/train_launches_r2g.csv
