#!/bin/bash
# GPU call 16 (round 2, 2 GPUs): data-parallel step as a CUDA graph (NCCL all-reduces captured) vs eager
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c16
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29511 tools/dp_check.py > $O/dp_check.txt 2>&1
tail -4 $O/dp_check.txt | cut -c1-300
for V in graph eager; do
  E=1; [ $V = eager ] && E=0
  CTX_DP_GRAPH=$E timeout 600 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_2gpu_$V.json 2> $O/bench_2gpu_$V.err
  tail -3 $O/bench_2gpu_$V.err | cut -c1-400
  CTX_DP_GRAPH=$E timeout 600 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --global-batch 256 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_2gpu_strong256_$V.json 2> $O/bench_2gpu_strong256_$V.err
  tail -3 $O/bench_2gpu_strong256_$V.err | cut -c1-400
  CTX_DP_GRAPH=$E timeout 600 $TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --global-batch 64 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_2gpu_strong64_$V.json 2> $O/bench_2gpu_strong64_$V.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c16/bench_*.json")):
    t=open(f).read().strip().splitlines()
    if not t: print(f,"EMPTY"); continue
    d=json.loads(t[-1]); print("%-60s n=%d B/gpu=%s ms=%.3f value=%.0f e2e=%.0f enqueue=%.3f"%(f.split("/")[-1],d["n_gpus"],d["config"].get("triples_per_gpu"),d["ms_per_step"],d["value"],d["e2e"]["value"],d.get("host_enqueue_ms_per_step",-1)))
PY
