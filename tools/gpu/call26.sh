#!/bin/bash
# GPU call 26 (round 2): 128-wide three-product tiles as CTA pairs (on / off)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c26
mkdir -p $O
timeout 400 python -m pytest tests/test_p16_gpu.py tests/test_model_gpu.py -x -q -m gpu > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt | cut -c1-300
for V in 1 0; do
  echo "== CTX_H16_PAIR128=$V"
  CTX_H16_PAIR128=$V timeout 200 python tools/layer_bench.py --only fprop,transpose --impl p3 --reps 10 2>&1 | grep -E "h1 |h2 |h3 |d_h1 |d_h2 |d_h3 " | cut -c1-110
  CTX_H16_PAIR128=$V timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_pair$V.json 2> $O/bench_pair$V.err
  CTX_H16_PAIR128=$V timeout 200 python bench.py --workload striker6 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_striker_pair$V.json 2> $O/bench_striker_pair$V.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c26/bench_*.json")):
    t=open(f).read().strip().splitlines()
    if not t: print(f,"EMPTY"); continue
    d=json.loads(t[-1]); print("%-36s ms=%.3f value=%.0f e2e=%.0f"%(f.split("/")[-1],d["ms_per_step"],d["value"],d["e2e"]["value"]))
PY
