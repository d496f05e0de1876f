#!/bin/bash
# GPU call 6 (round 2): ncu of the few-channel kernels (exact kernel names)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c6
mkdir -p $O
for L in fprop:h0:few_fprop_kernel:p3 wgrad:h0.wgrad:few_wgrad_kernel:p1 transpose:d_h4:few_transpose_kernel:p3; do
  K=$(echo $L | cut -d: -f1); NM=$(echo $L | cut -d: -f2); KR=$(echo $L | cut -d: -f3); I=$(echo $L | cut -d: -f4)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KR -c 1 -o $O/prof_${KR} -f \
    python tools/layer_bench.py --one $K:$NM --impl $I --reps 1 > $O/ncu_$KR.log 2>&1
  tail -2 $O/ncu_$KR.log
done
ls -la $O
