#!/bin/bash
# ncu evidence, 1 GPU (round 2): launch lists of the batch-32 inference step and the batch-16 train step (every command
# first exits 0 WITHOUT ncu), full captures of the top kernels.  usage: scripts/gpu_ncu_r2.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
ICMD="python scripts/infer_loop.py 6"
$ICMD > gpurun_out/plain_infer_$TAG.log 2>&1 || { echo "plain infer run failed"; tail -5 gpurun_out/plain_infer_$TAG.log; exit 1; }
# 2 weight-prep launches + 2 steps x 55 skipped, then 4 steps
ncu --metrics $M --clock-control none -s 112 -c 220 --csv --log-file gpurun_out/ncu_launches_infer_$TAG.csv $ICMD > gpurun_out/ncu1_$TAG.log 2>&1
echo "infer launch list rc=$?"
for K in wconv_kernel tconv_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 24 -c 4 -f -o gpurun_out/prof_${K}_$TAG $ICMD > gpurun_out/ncu_full_${K}_$TAG.log 2>&1
  echo "$K rc=$?"
done
TCMD="python scripts/train_loop.py 5"
$TCMD > gpurun_out/plain_train_$TAG.log 2>&1 || { echo "plain train run failed"; tail -5 gpurun_out/plain_train_$TAG.log; exit 1; }
# 3 steps skipped (first-touch, plan building), then one full step's launches
ncu --metrics $M --clock-control none -s 1240 -c 412 --csv --log-file gpurun_out/ncu_launches_train_$TAG.csv $TCMD > gpurun_out/ncu2_$TAG.log 2>&1
echo "train launch list rc=$?"
for K in xwgrad_kernel bn_bwd_apply_kernel head_bwd_weight_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 60 -c 2 -f -o gpurun_out/prof_${K}_$TAG $TCMD > gpurun_out/ncu_full_${K}_$TAG.log 2>&1
  echo "$K rc=$?"
done
ls -la gpurun_out/*_$TAG.ncu-rep
