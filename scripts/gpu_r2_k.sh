#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-library-baseline --no-train --steps 20 --profile-out gpurun_out/infer_launches_r2k_$tag.csv > gpurun_out/bench_r2k_$tag.json 2> gpurun_out/bench_r2k_$tag.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r2k_$tag.json') if l.startswith('{')][-1])
print('$tag', 'infer', round(d['value']), 'roof', round(d['roofline']['frac'],4))
"
  grep -E "decoder.blocks.4.conv1|decoder.blocks.3.conv1.0.weight.up" gpurun_out/infer_launches_r2k_$tag.csv
}
run default A=1
run noparstage UNETB200_NO_PARITY_STAGE=1
