#!/bin/bash
# GPU call 2 (round 2): parity of the 16-bit planes kernels (tc_h16.cu / planes.cu), per-layer timing of the new modes.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c2
mkdir -p $O
timeout 900 python -m pytest tests/test_p16_gpu.py -x -q -m gpu > $O/pytest_p16.txt 2>&1
tail -25 $O/pytest_p16.txt
timeout 600 python -m pytest tests/test_bf16x3_gpu.py -x -q -m gpu > $O/pytest_bf16x3.txt 2>&1
tail -5 $O/pytest_bf16x3.txt
timeout 600 python tools/layer_bench.py --impl x1,bf3,p3,p1 --only fprop,transpose,wgrad --reps 10 > $O/layer_p16.txt 2>&1
cat $O/layer_p16.txt
timeout 600 python tools/layer_bench.py --model inception2 --impl x1,bf3,p3,p1 --reps 5 > $O/layer_p16_inc2.txt 2>&1
cat $O/layer_p16_inc2.txt
