#!/bin/bash
# GPU call 13a (round 2, 1 GPU): full GPU test suite + gradient table + bench lines after the graph / rounding changes
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c13
mkdir -p $O
timeout 2400 python -m pytest tests -x -q -m gpu > $O/pytest_all.txt 2>&1
tail -8 $O/pytest_all.txt | cut -c1-400
timeout 600 python tools/grad_check.py --modes mixed > $O/grad_check.txt 2>&1
tail -42 $O/grad_check.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err
tail -c 1200 $O/bench_1gpu.json
timeout 900 python bench.py --workload inception2 --batch 256 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_inception2_b256.json 2> $O/bench_inception2_b256.err
tail -c 400 $O/bench_inception2_b256.json; tail -3 $O/bench_inception2_b256.err
timeout 900 python bench.py --workload striker6 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_striker6.json 2> $O/bench_striker6.err
tail -c 400 $O/bench_striker6.json; tail -3 $O/bench_striker6.err
