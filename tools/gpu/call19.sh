#!/bin/bash
# GPU call 19 (round 2): cluster split-K dense GEMMs (on / off), few_transpose gather, tightened gradient bars
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c19
mkdir -p $O
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "gemm" > $O/pytest_gemm.txt 2>&1
tail -3 $O/pytest_gemm.txt | cut -c1-300
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt | cut -c1-300
timeout 300 python tools/layer_bench.py --only gemm --impl x3,x1 --reps 10 > $O/layer_gemm_cluster.txt 2>&1
CTX_TC_CLUSTER_K=0 timeout 300 python tools/layer_bench.py --only gemm --impl x3,x1 --reps 10 > $O/layer_gemm_nocluster.txt 2>&1
paste -d'\n' $O/layer_gemm_cluster.txt $O/layer_gemm_nocluster.txt | cut -c1-140
timeout 300 python tools/layer_bench.py --one "transpose:d_h4" --impl p3 --reps 10 2>&1 | tail -2 | cut -c1-200
for V in 1 0; do
  CTX_TC_CLUSTER_K=$V timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_train_ck$V.json 2> $O/bench_train_ck$V.err
  CTX_TC_CLUSTER_K=$V timeout 300 python bench.py --batch 32 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_train_b32_ck$V.json 2> $O/bench_train_b32_ck$V.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c19/bench_*.json")):
    t=open(f).read().strip().splitlines()
    if not t: print(f,"EMPTY"); continue
    d=json.loads(t[-1]); print("%-40s ms/step %.3f value %.0f e2e %.0f"%(f.split("/")[-1],d["ms_per_step"], d["value"], d["e2e"]["value"]))
PY
