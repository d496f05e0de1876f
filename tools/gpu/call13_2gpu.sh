#!/bin/bash
# GPU call 13b (round 2, 2 GPUs): data-parallel correctness + weak / strong scaling lines at N = 2; graph-capture diagnostics
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c13
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -k "device_bank_equals or prefetch_is_transparent" -W "always" > $O/pytest_graphdiag.txt 2>&1
grep -A3 "UserWarning" $O/pytest_graphdiag.txt | head -30 | cut -c1-1200
timeout 600 $TR --nproc-per-node 2 --master-port 29511 tools/dp_check.py > $O/dp_check.txt 2>&1
tail -5 $O/dp_check.txt | cut -c1-300
timeout 600 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_2gpu.json 2> $O/bench_2gpu.err
tail -c 400 $O/bench_2gpu.json; tail -3 $O/bench_2gpu.err | cut -c1-300
timeout 600 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --global-batch 256 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_2gpu_strong256.json 2> $O/bench_2gpu_strong256.err
tail -c 400 $O/bench_2gpu_strong256.json; tail -3 $O/bench_2gpu_strong256.err | cut -c1-300
