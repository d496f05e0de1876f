#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c24
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for V in 0 1; do
CTX_DP_FINE=$V timeout 100 $TR --nproc-per-node 2 --master-port $((29600 + V)) bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_2gpu_fine$V.json 2> $O/bench_2gpu_fine$V.err; echo "fine=$V rc=$?"
CTX_DP_FINE=$V timeout 100 $TR --nproc-per-node 2 --master-port $((29610 + V)) bench.py --gpus 2 --batch 32 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_2gpu_b32_fine$V.json 2> $O/bench_2gpu_b32_fine$V.err; echo "fine=$V b32 rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c24/bench_*.json")):
    t=open(f).read().strip().splitlines()
    if not t: print(f,"EMPTY"); continue
    d=json.loads(t[-1]); print("%-36s n=%d B/gpu=%s ms=%.3f value=%.0f e2e=%.0f %s"%(f.split("/")[-1],d["n_gpus"],d["config"].get("triples_per_gpu"),d["ms_per_step"],d["value"],d["e2e"]["value"],d["scaling"]))
PY
