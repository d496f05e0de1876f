#!/bin/bash
# GPU call 23 (round 2, 8 GPUs): BASELINE configs[2] (B = 1024 over 8 GPUs), weak scaling at 256 / GPU, fixed global batch 256 at 4 and 8 GPUs
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c23
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { # name nproc args...
  local name=$1 np=$2; shift 2
  date +%T
  timeout 110 $TR --nproc-per-node $np --master-port $((29600 + RANDOM % 300)) bench.py --gpus $np --steps 20 --warmup 5 --no-cpu-baseline "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"
}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
run bench_8gpu_weak256 8
run bench_8gpu_b128_configs2 8 --batch 128
run bench_8gpu_global256 8 --global-batch 256
run bench_4gpu_global256 4 --global-batch 256
run bench_4gpu_weak256 4
timeout 100 $TR --nproc-per-node 8 --master-port 29955 tools/dp_check.py > $O/dp_check8.txt 2>&1; echo "dp_check rc=$?"
tail -4 $O/dp_check8.txt | cut -c1-200
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c23/bench_*.json")):
    t=open(f).read().strip().splitlines()
    if not t: print(f,"EMPTY"); continue
    d=json.loads(t[-1]); print("%-36s n=%d B/gpu=%s ms=%.3f value=%.0f e2e=%.0f %s"%(f.split("/")[-1],d["n_gpus"],d["config"].get("triples_per_gpu"),d["ms_per_step"],d["value"],d["e2e"]["value"],d["scaling"]))
PY
for f in $O/*.err; do if grep -q "Traceback\|Error" $f; then echo "== $f"; grep -v "^\*\*\*\|OMP_NUM" $f | head -30 | cut -c1-200; fi; done
