#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c34
mkdir -p $O
timeout 300 python -m pytest tests/test_model_gpu.py -x -q -m gpu -k "stream_schedule or graph_replay" > $O/pytest.txt 2>&1
tail -15 $O/pytest.txt | cut -c1-250
timeout -s TERM -k 10 60 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "rc=$?"; tail -c 300 $O/bench.json
