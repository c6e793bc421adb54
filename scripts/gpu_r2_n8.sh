#!/bin/bash
# N GPUs of one box: bench.py under torchrun (inference replicas + DDP train step, DP parity inside the line), then the
# BASELINE configs[3] / [4] sweep batch-sharded over all GPUs.  usage: scripts/gpu_r2_n8.sh <N>
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
nvidia-smi topo -m > gpurun_out/topo_r2_${N}gpu.txt 2>&1
timeout 900 $TR bench.py --gpus $N --steps 20 > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err
echo "bench N=$N rc=$?"
timeout 900 $TR scripts/configs_sweep.py > gpurun_out/configs_sweep_r2_${N}gpu.jsonl 2> gpurun_out/configs_sweep_r2_${N}gpu.err
echo "sweep N=$N rc=$?"; tail -3 gpurun_out/configs_sweep_r2_${N}gpu.jsonl
python - <<PY
import json
line = [l for l in open("gpurun_out/bench_r2_${N}gpu.json") if l.startswith("{")][-1]
d = json.loads(line); t = d["train"]
print("N=$N infer", round(d["value"]), "e2e", round(d["e2e"]["value"]), "fp32 frames", round(d["e2e"]["fp32_frames"]["value"]), "| train ms", round(t["ms_per_step"], 3),
      "img/s", round(t["value"]), "e2e", round(t["e2e"]["value"]), round(t["e2e"]["fp32_frames"]["value"]), {k: round(v, 3) for k, v in t["phases"].items()},
      t.get("dp_parity_rel_l2"), t.get("dp_grads_identical_across_ranks"))
PY
