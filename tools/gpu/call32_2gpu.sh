#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c32
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 100 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "weak rc=$?"
timeout 100 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > $O/bench_2gpu_ref.json 2> $O/bench_2gpu_ref.err; echo "ref rc=$?"
timeout 100 $TR --nproc-per-node 2 --master-port 29514 tools/dp_check.py > $O/dp_check.txt 2>&1; echo "dp rc=$?"; tail -4 $O/dp_check.txt | cut -c1-200
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c32/bench_*.json")):
    t=open(f).read().strip().splitlines()
    if not t: print(f,"EMPTY"); continue
    d=json.loads(t[-1]); print("%-20s n=%s ms=%.3f value=%.0f clocks=%s impl=%s"%(f.split("/")[-1],d["n_gpus"],d["ms_per_step"],d["value"],d.get("clocks"),d.get("impl")))
PY
pgrep -a nvidia-smi | head -3
