#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c14
mkdir -p $O
timeout 300 python tools/graph_debug.py > $O/graph_debug.txt 2>&1
cat $O/graph_debug.txt | cut -c1-400
timeout 600 python tools/layer_bench.py --impl p1 --only transpose,fprop --reps 10 > $O/layer_p1.txt 2>&1
grep -E "dgrad" $O/layer_p1.txt | cut -c1-120
timeout 600 python -m pytest tests/test_p16_gpu.py -x -q -m gpu > $O/pytest_p16.txt 2>&1
tail -3 $O/pytest_p16.txt | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_train.json 2> $O/bench_train.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c14/bench_train.json").read().strip().splitlines()[-1])
print("ms/step %.3f value %.0f e2e %.0f"%(d["ms_per_step"], d["value"], d["e2e"]["value"]))
for k,v in d["kernel_breakdown"].items(): print("  %-24s %8.1f TF/s %7.3f ms %3d calls"%(k,v["tflops"],v["ms"],v["calls"]))
PY
