#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c31
mkdir -p $O
for i in 1 2 3; do
timeout -s TERM -k 10 60 python bench.py --no-cpu-baseline > $O/bench_$i.json 2> $O/bench_$i.err; echo "rc=$?"
done
timeout -s TERM -k 10 60 python bench.py --workload inception --no-cpu-baseline > $O/bench_inc.json 2> $O/bench_inc.err; echo "rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c31/bench_*.json")):
    t=open(f).read().strip().splitlines()
    if not t: print(f,"EMPTY"); continue
    d=json.loads(t[-1]); print("%-20s ms=%.3f value=%.0f clocks=%s share=%s"%(f.split("/")[-1],d["ms_per_step"],d["value"],d["clocks"],(d.get("roofline") or {}).get("share_of_step")))
PY
pgrep -a nvidia-smi | head -3; echo "orphans above (none expected)"
