#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_trained.py -m gpu -q -s -p no:cacheprovider -k "fp16 or 200_steps or oracle_reaches" > gpurun_out/pytest_r2f.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_r2f.log
timeout 300 python bench.py --no-cpu-baseline --no-library-baseline --no-train --steps 20 > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err
python -c "
import json
d=json.load(open('gpurun_out/bench_r2f.json'))
print('infer', round(d['value']), 'e2e', round(d['e2e']['value']), 'fp16', d.get('fp16_operands'))
"
