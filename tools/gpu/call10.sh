#!/bin/bash
# GPU call 10 (round 2): few kernels (deferred store drain; small wgrad CTAs) -> parity, timing, trainer bench
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c10
mkdir -p $O
timeout 900 python -m pytest tests/test_p16_gpu.py -x -q -m gpu -k few > $O/pytest_few.txt 2>&1
tail -12 $O/pytest_few.txt | cut -c1-300
timeout 600 python tools/layer_bench.py --impl p3,p1 --only fprop,wgrad,transpose --reps 10 > $O/layer_few.txt 2>&1
grep -E "h0|d_h4" $O/layer_few.txt | cut -c1-160
timeout 1500 python -m pytest tests/test_model_gpu.py -x -q -m gpu > $O/pytest_model.txt 2>&1
tail -5 $O/pytest_model.txt | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_train.json 2> $O/bench_train.err
tail -c 1500 $O/bench_train.json; tail -5 $O/bench_train.err
