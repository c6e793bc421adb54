#!/bin/bash
# ncu evidence, 1 GPU: inference launch list + full captures (wconv / tconv), train-step launch list + full captures
# (xwgrad, bn_bwd_apply).  usage: scripts/gpu_ncu_s5.sh <tag>
TAG=${1:-s5}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain infer run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 70 -c 212 --csv --log-file gpurun_out/ncu_launches_infer_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
echo "infer launch list rc=$?"
for K in wconv_kernel tconv_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 12 -c 3 -f -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu_full_${K}_$TAG.log 2>&1
  echo "$K rc=$?"
done
TCMD="python scripts/train_loop.py 5"
$TCMD > gpurun_out/plain_train_$TAG.log 2>&1 || { echo "plain train run failed"; tail -5 gpurun_out/plain_train_$TAG.log; exit 1; }
# 3 steps skipped (first-touch, plan building), then one full step's launches
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1300 -c 420 --csv --log-file gpurun_out/ncu_launches_train_$TAG.csv $TCMD > gpurun_out/ncu2_$TAG.log 2>&1
echo "train launch list rc=$?"
for K in xwgrad_kernel bn_bwd_apply_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 60 -c 3 -f -o gpurun_out/prof_${K}_$TAG $TCMD > gpurun_out/ncu_full_${K}_$TAG.log 2>&1
  echo "$K rc=$?"
done
ls -la gpurun_out/*_$TAG.ncu-rep
