#!/bin/bash
# GPU call 18 (round 2): dense GEMM epilogue (vector reductions for +=), lower split-K threshold, prefetched += in h16 epilogue
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c18
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dump-launches $O/launch_profile.jsonl > $O/bench_train.json 2> $O/bench_train.err
timeout 600 python bench.py --batch 32 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_train_b32.json 2> $O/bench_train_b32.err
python - <<'PY'
import json
for f in ("bench_train","bench_train_b32"):
    d=json.loads(open("gpurun_out/r2c18/%s.json"%f).read().strip().splitlines()[-1])
    print(f,"ms/step %.3f value %.0f e2e %.0f"%(d["ms_per_step"], d["value"], d["e2e"]["value"]))
for L in open("gpurun_out/r2c18/launch_profile.jsonl"):
    d=json.loads(L)
    if d["kind"]=="gemm" or d["dims"][0]==256: print("%-10s impl %d %-40s %8.2f GF %8.1f us %7.1f TF/s"%(d["kind"],d["impl"],d["dims"],d["gflop"],d["us"],d["tflops"]))
PY
timeout 600 python tools/layer_bench.py --only gemm --impl x3,x1 --reps 10 > $O/layer_gemm.txt 2>&1
cat $O/layer_gemm.txt | cut -c1-140
