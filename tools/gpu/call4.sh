#!/bin/bash
# GPU call 4 (round 2): few-channel image-side kernels (tc_few.cu): parity, per-layer timing, model parity, bench.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c4
mkdir -p $O
timeout 900 python -m pytest tests/test_p16_gpu.py -x -q -m gpu  > $O/pytest_few.txt 2>&1
tail -30 $O/pytest_few.txt
timeout 600 python tools/layer_bench.py --impl x1,p3,p1 --only fprop,wgrad --reps 10 > $O/layer_few.txt 2>&1
head -8 $O/layer_few.txt
timeout 600 python tools/grad_check.py --modes mixed > $O/grad_check.txt 2>&1
tail -45 $O/grad_check.txt
timeout 1500 python -m pytest tests/test_model_gpu.py -x -q -m gpu > $O/pytest_model.txt 2>&1
tail -15 $O/pytest_model.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_train.json 2> $O/bench_train.err
tail -c 1800 $O/bench_train.json; tail -5 $O/bench_train.err
