#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c25
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for V in 2 4 8 16; do
NCCL_MAX_NCHANNELS=$V NCCL_MIN_NCHANNELS=1 timeout 100 $TR --nproc-per-node 2 --master-port $((29600 + V)) bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_2gpu_nch$V.json 2> $O/bench_2gpu_nch$V.err; echo "nch=$V rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c25/bench_*.json")):
    t=open(f).read().strip().splitlines()
    if not t: print(f,"EMPTY"); continue
    d=json.loads(t[-1]); print("%-36s n=%d B/gpu=%s ms=%.3f value=%.0f e2e=%.0f %s"%(f.split("/")[-1],d["n_gpus"],d["config"].get("triples_per_gpu"),d["ms_per_step"],d["value"],d["e2e"]["value"],d["scaling"]))
PY
