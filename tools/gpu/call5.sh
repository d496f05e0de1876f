#!/bin/bash
# GPU call 5 (round 2): few-channel kernels — parity incl. the transposed conv, timing, ncu of few_fprop / few_wgrad
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c5
mkdir -p $O
timeout 900 python -m pytest tests/test_p16_gpu.py -x -q -m gpu -k "few" > $O/pytest_few.txt 2>&1
tail -30 $O/pytest_few.txt
timeout 600 python tools/layer_bench.py --impl x1,p3,p1 --only fprop,wgrad,transpose --reps 10 > $O/layer_few.txt 2>&1
grep -E "h0|d_h4" $O/layer_few.txt
for L in fprop:h0 wgrad:h0.wgrad transpose:d_h4; do
  N=$(echo $L | tr ':' '_')
  I=p3; [ "$L" = "wgrad:h0.wgrad" ] && I=p1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:few_ -c 1 -o $O/prof_few_$N -f \
    python tools/layer_bench.py --one $L --impl $I --reps 1 > $O/ncu_$N.log 2>&1
  tail -2 $O/ncu_$N.log
done
ls -la $O
