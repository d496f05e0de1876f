#!/bin/bash
# GPU call 3 (round 2): the trainer on fp16 planes (forward 3 products, backward 1 product): parity suite, gradient table, bench.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c3
mkdir -p $O
timeout 900 python -m pytest tests/test_p16_gpu.py tests/test_bf16x3_gpu.py -x -q -m gpu > $O/pytest_p16.txt 2>&1
tail -5 $O/pytest_p16.txt
timeout 600 python tools/grad_check.py --modes fp32,mixed > $O/grad_check.txt 2>&1
cat $O/grad_check.txt
timeout 1500 python -m pytest tests/test_model_gpu.py -x -q -m gpu > $O/pytest_model.txt 2>&1
tail -15 $O/pytest_model.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_train.json 2> $O/bench_train.err
tail -c 3500 $O/bench_train.json; tail -5 $O/bench_train.err
CTX_BWD_P16=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_train_nop1.json 2> $O/bench_train_nop1.err
tail -c 600 $O/bench_train_nop1.json
timeout 900 python bench.py --workload inception2 --batch 256 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_inception2_b256.json 2> $O/bench_inception2_b256.err
tail -c 2500 $O/bench_inception2_b256.json; tail -5 $O/bench_inception2_b256.err
