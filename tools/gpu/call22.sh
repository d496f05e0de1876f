#!/bin/bash
# GPU call 22 (round 2): full GPU suite on the current build; repeated bench runs to catch an intermittent stall
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c22
mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt | cut -c1-300
for i in 1 2 3 4 5 6; do
  date +%T
  timeout 90 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_$i.json 2> $O/bench_$i.err; echo "run $i rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c22/bench_*.json")):
    t=open(f).read().strip().splitlines()
    if not t: print(f,"EMPTY"); continue
    d=json.loads(t[-1]); print("%-20s ms/step %.3f value %.0f e2e %.0f (plain %.0f)"%(f.split("/")[-1],d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["value_unmodified_call_site"]))
PY
for f in $O/bench_*.err; do if [ -s $f ]; then echo "== $f"; head -60 $f | cut -c1-200; fi; done
