#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_r2.json
timeout 1500 python -m pytest tests/test_gpu_trained.py tests/test_gpu_native.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_r2e.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_r2e.log
timeout 300 python bench.py --no-cpu-baseline --no-library-baseline --no-train --steps 20 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err
python -c "
import json
d=json.load(open('gpurun_out/bench_r2e.json'))
print('infer', round(d['value']), 'e2e', round(d['e2e']['value']), 'fp16', d.get('fp16_operands'))
"
