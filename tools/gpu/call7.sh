#!/bin/bash
# GPU call 7 (round 2): async few-channel kernels + fused loss + scaled filter planes + rounded dense operands
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c7
mkdir -p $O
timeout 900 python -m pytest tests/test_p16_gpu.py tests/test_bf16x3_gpu.py -x -q -m gpu > $O/pytest_p16.txt 2>&1
tail -30 $O/pytest_p16.txt | cut -c1-300
timeout 600 python tools/layer_bench.py --impl x1,p3,p1 --only fprop,wgrad,transpose --reps 10 > $O/layer_few.txt 2>&1
grep -E "h0|d_h4" $O/layer_few.txt
timeout 600 python tools/grad_check.py --modes mixed > $O/grad_check.txt 2>&1
tail -45 $O/grad_check.txt
timeout 1500 python -m pytest tests/test_model_gpu.py -x -q -m gpu > $O/pytest_model.txt 2>&1
tail -15 $O/pytest_model.txt | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_train.json 2> $O/bench_train.err
tail -c 1800 $O/bench_train.json; tail -5 $O/bench_train.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
tail -2 $O/ncu_bench.log | cut -c1-300
