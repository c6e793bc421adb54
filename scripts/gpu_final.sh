#!/bin/bash
# Round-end evidence on one GPU: full GPU test suite, smoke, default bench (with CPU baseline), reference arm, clocks,
# per-launch CUDA-event tables.  usage: scripts/gpu_final.sh <tag>
TAG=${1:-s5}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks_$TAG.csv &
SMI=$!
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest_exit=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke_exit=$?" >> gpurun_out/smoke_$TAG.log
timeout 900 python bench.py --profile-out gpurun_out/infer_launches_$TAG.csv --train-profile-out gpurun_out/train_launches_$TAG.csv > gpurun_out/bench_$TAG.log 2>&1; echo "bench_exit=$?" >> gpurun_out/bench_$TAG.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; echo "ref_exit=$?" >> gpurun_out/bench_ref_$TAG.log
kill $SMI
tail -2 gpurun_out/pytest_gpu_$TAG.log gpurun_out/smoke_$TAG.log
tail -c 600 gpurun_out/bench_$TAG.log; echo; tail -c 900 gpurun_out/bench_ref_$TAG.log
