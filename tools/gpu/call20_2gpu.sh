#!/bin/bash
# GPU call 20 (round 2, 2 GPUs): data-parallel step as a CUDA graph with orderly shutdown; every command under a short timeout
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c20
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
date +%T
timeout 120 $TR --nproc-per-node 2 --master-port 29511 tools/dp_check.py > $O/dp_check.txt 2>&1; echo "dp_check rc=$?"
tail -4 $O/dp_check.txt | cut -c1-300
date +%T
timeout 100 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_2gpu_graph.json 2> $O/bench_2gpu_graph.err; echo "weak rc=$?"
date +%T
timeout 100 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --global-batch 256 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_2gpu_strong256_graph.json 2> $O/bench_2gpu_strong256_graph.err; echo "strong rc=$?"
date +%T
CTX_DP_GRAPH=0 timeout 100 $TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --global-batch 256 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_2gpu_strong256_eager.json 2> $O/bench_2gpu_strong256_eager.err; echo "strong eager rc=$?"
date +%T
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c20/bench_*.json")):
    t=open(f).read().strip().splitlines()
    if not t: print(f,"EMPTY"); continue
    d=json.loads(t[-1]); print("%-40s n=%d B/gpu=%s ms=%.3f value=%.0f e2e=%.0f"%(f.split("/")[-1],d["n_gpus"],d["config"].get("triples_per_gpu"),d["ms_per_step"],d["value"],d["e2e"]["value"]))
PY
tail -3 $O/*.err | cut -c1-200
