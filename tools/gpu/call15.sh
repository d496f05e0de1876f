#!/bin/bash
# GPU call 15 (round 2): full GPU test suite, bench with per-launch dump, ncu launch list of the bench, ncu --set full of h16_kernel
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c15
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.txt 2>&1
tail -3 $O/pytest_all.txt | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dump-launches $O/launch_profile.jsonl > $O/bench_train.json 2> $O/bench_train.err
tail -c 600 $O/bench_train.json; tail -3 $O/bench_train.err | cut -c1-300
python - <<'PY'
import json
for L in open("gpurun_out/r2c15/launch_profile.jsonl"):
    d=json.loads(L); print("%-10s impl %d %-40s %8.2f GF %8.1f us %7.1f TF/s"%(d["kind"],d["impl"],d["dims"],d["gflop"],d["us"],d["tflops"]))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
wc -l $O/launches_bench.csv
for L in fprop:h2:p3 transpose:d_h2:p3 transpose:d_h3:p3 fprop:d_h2.dgrad:p1 transpose:h1.dgrad:p1 wgrad:d_h2.wgrad:p1 fprop:h1:p3; do
  K=$(echo $L | cut -d: -f1,2); I=$(echo $L | cut -d: -f3); N=$(echo $L | tr ':' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:h16_kernel -c 1 -o $O/prof_$N -f \
    python tools/layer_bench.py --batch 256 --one $K --impl $I --reps 1 > $O/ncu_$N.log 2>&1
  tail -1 $O/ncu_$N.log | cut -c1-200
done
ls -la $O
