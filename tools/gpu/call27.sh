#!/bin/bash
# GPU call 27 (round 2): context encoder's backward chain on its own streams (on / off)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c27
mkdir -p $O
timeout 400 python -m pytest tests/test_model_gpu.py tests/test_inception_gpu.py -x -q -m gpu > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt | cut -c1-300
for V in 1 0; do
  CTX_BWD_CTX_STREAM=$V timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_ctxpar$V.json 2> $O/bench_ctxpar$V.err
  CTX_BWD_CTX_STREAM=$V timeout 200 python bench.py --batch 32 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_b32_ctxpar$V.json 2> $O/bench_b32_ctxpar$V.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c27/bench_*.json")):
    t=open(f).read().strip().splitlines()
    if not t: print(f,"EMPTY"); continue
    d=json.loads(t[-1]); print("%-36s ms=%.3f value=%.0f e2e=%.0f"%(f.split("/")[-1],d["ms_per_step"],d["value"],d["e2e"]["value"]))
PY
