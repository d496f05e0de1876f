#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c33
mkdir -p $O
timeout 400 python -m pytest tests/test_model_gpu.py -x -q -m gpu > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt | cut -c1-300
for i in 1 2 3; do
timeout -s TERM -k 10 60 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_$i.json 2> $O/bench_$i.err; echo "rc=$?"
done
timeout -s TERM -k 10 60 python bench.py --batch 32 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_b32.json 2> $O/bench_b32.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c33/bench_*.json")):
    t=open(f).read().strip().splitlines()
    if not t: print(f,"EMPTY"); continue
    d=json.loads(t[-1]); print("%-20s ms=%.3f value=%.0f e2e=%.0f"%(f.split("/")[-1],d["ms_per_step"],d["value"],d["e2e"]["value"]))
PY
