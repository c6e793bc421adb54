#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_r2.json
timeout 1200 python -m pytest tests/test_gpu_trained.py tests/test_gpu_train.py tests/test_gpu_reference_loop.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_r2d.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_r2d.log
timeout 300 python bench.py --no-cpu-baseline --no-library-baseline --steps 20 --train-profile-out gpurun_out/train_launches_r2d.csv > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err
python -c "
import json
d=json.load(open('gpurun_out/bench_r2d.json'))
t=d['train']; print('infer', round(d['value']), 'e2e', round(d['e2e']['value']), 'train ms', round(t['ms_per_step'],3), {k:round(v,3) for k,v in t['phases'].items()}, 'e2e', round(t['e2e']['value']), round(t['e2e']['fp32_frames']['value']))
"
timeout 300 python scripts/ab_two_lane.py > gpurun_out/ab_two_lane_r2d.log 2>&1; cat gpurun_out/ab_two_lane_r2d.log | tail -6
B=64 timeout 300 python scripts/ab_two_lane.py > gpurun_out/ab_two_lane_b64_r2d.log 2>&1; cat gpurun_out/ab_two_lane_b64_r2d.log | tail -4
