#!/bin/bash
# N GPUs: bench only (final state), plus N=1 on the same box for the all-reduce cost
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 20 > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err
echo "bench N=$N rc=$?"
timeout 300 python bench.py --no-cpu-baseline --no-library-baseline --steps 20 > gpurun_out/bench_r2_1gpu_samebox8.json 2> gpurun_out/bench_r2_1gpu_samebox8.err
python - <<PY
import json
for f in ["gpurun_out/bench_r2_${N}gpu.json", "gpurun_out/bench_r2_1gpu_samebox8.json"]:
    d = json.loads([l for l in open(f) if l.startswith("{")][-1]); t = d["train"]
    print(f, "infer", round(d["value"]), "e2e", round(d["e2e"]["value"]), "| train ms", round(t["ms_per_step"], 3), "img/s", round(t["value"]), "e2e", round(t["e2e"]["value"]),
          {k: round(v, 3) for k, v in t["phases"].items()}, t.get("dp_parity_rel_l2"))
PY
