#!/bin/bash
# GPU call 35 (round 2): final build — full GPU suite, smoke, every single-GPU bench line
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c35
mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt | cut -c1-200
timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
timeout 300 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"
timeout 100 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_train.json 2> $O/bench_train.err
timeout 100 python bench.py --batch 32 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_b32.json 2> $O/bench_b32.err
timeout 100 python bench.py --workload inception2 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_inception2.json 2> $O/bench_inception2.err
timeout 100 python bench.py --workload striker6 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_striker6.json 2> $O/bench_striker6.err
timeout 100 python bench.py --workload inception --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_inception.json 2> $O/bench_inception.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c35/bench_*.json")):
    t=open(f).read().strip().splitlines()
    if not t: print(f,"EMPTY"); continue
    d=json.loads(t[-1])
    r=d.get("roofline") or {}
    print("%-28s ms=%.3f value=%.0f e2e=%.0f frac=%s share=%s clocks=%s"%(f.split("/")[-1],d["ms_per_step"],d["value"],(d.get("e2e") or {}).get("value",0),r.get("frac"),r.get("share_of_step"),(d.get("clocks") or {}).get("sm_mhz")))
PY
