#!/bin/bash
# GPU call 11 (round 2): h16 kernel with the TMA-store epilogue: parity + per-layer timing, with / without
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c11
mkdir -p $O
timeout 900 python -m pytest tests/test_p16_gpu.py tests/test_bf16x3_gpu.py -x -q -m gpu > $O/pytest_p16.txt 2>&1
tail -12 $O/pytest_p16.txt | cut -c1-300
timeout 600 python tools/layer_bench.py --impl p3,p1 --only fprop,wgrad,transpose --reps 10 > $O/layer_tma.txt 2>&1
cat $O/layer_tma.txt | cut -c1-160
CTX_TMA_STORE=0 timeout 600 python tools/layer_bench.py --impl p3,p1 --only fprop,transpose --reps 10 > $O/layer_notma.txt 2>&1
cat $O/layer_notma.txt | cut -c1-160
timeout 600 python tools/layer_bench.py --model inception2 --impl p3,p1 --reps 5 > $O/layer_tma_inc2.txt 2>&1
cat $O/layer_tma_inc2.txt | cut -c1-160
timeout 1500 python -m pytest tests/test_model_gpu.py -x -q -m gpu > $O/pytest_model.txt 2>&1
tail -5 $O/pytest_model.txt | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_train.json 2> $O/bench_train.err
tail -c 600 $O/bench_train.json; tail -5 $O/bench_train.err
