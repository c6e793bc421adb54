#!/bin/bash
# ncu evidence for the inference step (1 GPU): launch list (device time per launch) + full captures of the three conv kernels.
# usage: scripts/gpu_ncu.sh <tag>      (run under gpurun; reads nothing outside the repo)
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
# skip the first forward pass's launches (plan building / first-touch), then list 3 forward passes' worth of launches
ncu --metrics gpu__time_duration.sum --clock-control none -s 70 -c 200 --csv --log-file gpurun_out/ncu_launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
echo "launch list rc=$?"
for K in wconv_kernel tconv_kernel igemm_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 12 -c 3 -f -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu_full_${K}_$TAG.log 2>&1
  echo "$K rc=$?"
done
ls -la gpurun_out/*.ncu-rep
