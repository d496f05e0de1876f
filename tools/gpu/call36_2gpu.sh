#!/bin/bash
# GPU call 36 (round 2, 2 GPUs): data parallel with the context encoder's chain on its own streams x finer all-reduce buckets
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c36
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
CTX_DP_FINE=1 timeout 100 $TR --nproc-per-node 2 --master-port 29511 tools/dp_check.py > $O/dp_check_fine1.txt 2>&1; echo "dp rc=$?"; tail -4 $O/dp_check_fine1.txt | cut -c1-200
timeout 100 $TR --nproc-per-node 2 --master-port 29512 tools/dp_check.py > $O/dp_check_fine0.txt 2>&1; echo "dp rc=$?"; tail -4 $O/dp_check_fine0.txt | cut -c1-200
for F in 0 1; do for C in 1 0; do
CTX_DP_FINE=$F CTX_BWD_CTX_STREAM=$C timeout 100 $TR --nproc-per-node 2 --master-port $((29520 + 2*F + C)) bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_2gpu_fine${F}_ctx${C}.json 2> $O/bench_2gpu_fine${F}_ctx${C}.err; echo "fine=$F ctx=$C rc=$?"
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c36/bench_*.json")):
    t=open(f).read().strip().splitlines()
    if not t: print(f,"EMPTY"); continue
    d=json.loads(t[-1]); print("%-36s n=%d ms=%.3f value=%.0f e2e=%.0f"%(f.split("/")[-1],d["n_gpus"],d["ms_per_step"],d["value"],d["e2e"]["value"]))
PY
