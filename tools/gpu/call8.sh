#!/bin/bash
# GPU call 8 (round 2): few_fprop with the TMA-store epilogue
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c8
mkdir -p $O
timeout 900 python -m pytest tests/test_p16_gpu.py -x -q -m gpu -k few > $O/pytest_few.txt 2>&1
tail -12 $O/pytest_few.txt | cut -c1-300
timeout 600 python tools/layer_bench.py --impl p3,p1 --only fprop,wgrad,transpose --reps 10 > $O/layer_few.txt 2>&1
grep -E "h0|d_h4" $O/layer_few.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:few_fprop_kernel -c 1 -o $O/prof_few_fprop -f \
    python tools/layer_bench.py --one fprop:h0 --impl p3 --reps 1 > $O/ncu_few_fprop.log 2>&1
tail -2 $O/ncu_few_fprop.log
