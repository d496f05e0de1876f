#!/bin/bash
# GPU call 12 (round 2): CUDA-graph replay of the training step
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c12
mkdir -p $O
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu -k "cuda_graph or adam_training or uint8_feeds or trainer" > $O/pytest_graph.txt 2>&1
tail -25 $O/pytest_graph.txt | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_train.json 2> $O/bench_train.err
tail -c 700 $O/bench_train.json; tail -5 $O/bench_train.err
CTX_GRAPH=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_train_nograph.json 2> $O/bench_train_nograph.err
tail -c 300 $O/bench_train_nograph.json
timeout 600 python bench.py --batch 32 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_train_b32.json 2> $O/bench_train_b32.err
tail -c 300 $O/bench_train_b32.json; tail -3 $O/bench_train_b32.err
timeout 1500 python -m pytest tests/test_model_gpu.py tests/test_bf16x3_gpu.py -x -q -m gpu > $O/pytest_model.txt 2>&1
tail -5 $O/pytest_model.txt | cut -c1-400
