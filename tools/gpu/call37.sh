#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c37
mkdir -p $O
timeout 200 python -m pytest tests/test_model_gpu.py -x -q -m gpu -k "b32 or graph_replay or stream_schedule or session_run or golden" > $O/pytest.txt 2>&1
tail -2 $O/pytest.txt | cut -c1-200
timeout 150 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c37/bench_default.json").read().strip().splitlines()[-1])
print("ms=%.3f value=%.0f e2e=%.0f cpu=%.0f clocks=%s launches=%s"%(d["ms_per_step"],d["value"],d["e2e"]["value"],d["cpu_baseline"]["value"],d["clocks"],d["gpu_launches"]))
PY
