#!/bin/bash
# multi-GPU: DP parity test, bench at N GPUs with the high-priority and the normal-priority all-reduce, N = 1 on the same box
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_dp_r2_${N}gpu.log 2>&1
echo "dp pytest rc=$?"; tail -3 gpurun_out/pytest_dp_r2_${N}gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 20 > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err
echo "bench N=$N rc=$?"
UNETB200_DP_NORMAL_PRIORITY=1 timeout 600 $TR bench.py --gpus $N --steps 20 > gpurun_out/bench_r2_${N}gpu_normalprio.json 2> gpurun_out/bench_r2_${N}gpu_normalprio.err
echo "bench N=$N normal priority rc=$?"
timeout 300 python bench.py --no-cpu-baseline --no-library-baseline --steps 20 > gpurun_out/bench_r2_1gpu_samebox.json 2> gpurun_out/bench_r2_1gpu_samebox.err
python - <<PY
import json
for tag in ["${N}gpu", "${N}gpu_normalprio", "1gpu_samebox"]:
    try:
        d = json.load(open(f"gpurun_out/bench_r2_{tag}.json"))
        t = d["train"]
        print(tag, "infer", round(d["value"]), "e2e", round(d["e2e"]["value"]), "| train ms", round(t["ms_per_step"], 3), "img/s", round(t["value"]),
              "e2e", round(t["e2e"]["value"]), {k: round(v, 3) for k, v in t["phases"].items()}, t.get("dp_parity_rel_l2"), t.get("dp_grads_identical_across_ranks"))
    except Exception as e:
        print(tag, "failed", e)
PY
