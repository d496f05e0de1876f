#!/bin/bash
# GPU call 13 (round 2, 2 GPUs): data-parallel correctness + weak / strong scaling lines at N = 2
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c13
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29511 tools/dp_check.py > $O/dp_check.txt 2>&1
tail -5 $O/dp_check.txt | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_1gpu.json 2> $O/bench_1gpu.err
tail -c 300 $O/bench_1gpu.json
timeout 600 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_2gpu.json 2> $O/bench_2gpu.err
tail -c 400 $O/bench_2gpu.json; tail -3 $O/bench_2gpu.err | cut -c1-300
timeout 600 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --global-batch 256 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_2gpu_strong256.json 2> $O/bench_2gpu_strong256.err
tail -c 400 $O/bench_2gpu_strong256.json; tail -3 $O/bench_2gpu_strong256.err | cut -c1-300
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_bf16x3_gpu.py -x -q -m gpu > $O/pytest_model.txt 2>&1
tail -5 $O/pytest_model.txt | cut -c1-400
timeout 600 python tools/grad_check.py --modes mixed > $O/grad_check.txt 2>&1
tail -42 $O/grad_check.txt
