#!/bin/bash
# round 2, call A: full GPU test suite with printed distances, bench (both arms), train launch table
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_r2a.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_r2a.log
tail -5 gpurun_out/pytest_r2a.log
timeout 600 python bench.py --profile-out gpurun_out/infer_launches_r2a.csv --train-profile-out gpurun_out/train_launches_r2a.csv > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r2a.json 2> gpurun_out/bench_ref_r2a.err
echo "ref rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_r2a.json'))
print('infer', d['value'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], 'train ms', d['train']['ms_per_step'], 'sustained', d['sustained'])
print(json.dumps(d.get('library_baseline'), indent=1))
"
