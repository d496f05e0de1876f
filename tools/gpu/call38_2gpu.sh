#!/bin/bash
# GPU call 38 (round 2, 2 GPUs): does the data-parallel stall of the context-encoder chain go away with more hardware queues?
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c38
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
CTX_DP_CTX_STREAM=1 CUDA_DEVICE_MAX_CONNECTIONS=32 timeout -k 5 45 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_2gpu_ctx_conn32.json 2> $O/bench_2gpu_ctx_conn32.err; echo "conn32 rc=$?"
tail -c 400 $O/bench_2gpu_ctx_conn32.json | cut -c1-400
python - <<'PY'
import json
t=open("gpurun_out/r2c38/bench_2gpu_ctx_conn32.json").read().strip().splitlines()
if t:
    d=json.loads(t[-1]); print("n=%d ms=%.3f value=%.0f e2e=%.0f"%(d["n_gpus"],d["ms_per_step"],d["value"],d["e2e"]["value"]))
else: print("EMPTY (stalled)")
PY
