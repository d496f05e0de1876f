#!/bin/bash
# GPU call 29 (round 2): Adam slice of the encoders' dense layers under the conv backward (on / off)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c29
mkdir -p $O
timeout 400 python -m pytest tests/test_model_gpu.py -x -q -m gpu > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt | cut -c1-300
for V in 1 0 1 0; do
  CTX_ADAM_DENSE_EARLY=$V timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_ade${V}_$RANDOM.json 2> $O/err.txt
done
CTX_ADAM_DENSE_EARLY=1 timeout 200 python bench.py --batch 32 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_b32_ade1.json 2> $O/err.txt
CTX_ADAM_DENSE_EARLY=0 timeout 200 python bench.py --batch 32 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_b32_ade0.json 2> $O/err.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c29/bench_*.json")):
    t=open(f).read().strip().splitlines()
    if not t: print(f,"EMPTY"); continue
    d=json.loads(t[-1]); print("%-36s ms=%.3f value=%.0f e2e=%.0f"%(f.split("/")[-1],d["ms_per_step"],d["value"],d["e2e"]["value"]))
PY
