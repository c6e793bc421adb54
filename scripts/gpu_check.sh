#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, ncu launch list + full captures of three representative conv launches.
# usage: scripts/gpu_check.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks_$TAG.csv &
SMI=$!
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest_exit=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke_exit=$?" >> gpurun_out/smoke_$TAG.log
timeout 600 python bench.py --profile-out gpurun_out/infer_launches_events_$TAG.csv > gpurun_out/bench_$TAG.log 2>&1; echo "bench_exit=$?" >> gpurun_out/bench_$TAG.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1
kill $SMI
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/ncu_launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
for S in 60 1 24; do
  ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s $S -c 1 -o gpurun_out/prof_igemm_s${S}_$TAG -f $CMD > gpurun_out/ncu_full_s${S}_$TAG.log 2>&1
done
tail -3 gpurun_out/pytest_gpu_$TAG.log gpurun_out/smoke_$TAG.log gpurun_out/bench_$TAG.log
