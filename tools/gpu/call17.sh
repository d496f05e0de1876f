#!/bin/bash
# GPU call 17 (round 2): grad_planes lane fix check, dense-GEMM ncu captures (tc_igemm), Inception2 3x3 layers on h16_kernel under ncu
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c17
mkdir -p $O
timeout 900 python -m pytest tests/test_p16_gpu.py tests/test_model_gpu.py -x -q -m gpu > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_train.json 2> $O/bench_train.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c17/bench_train.json").read().strip().splitlines()[-1])
print("ms/step %.3f value %.0f e2e %.0f"%(d["ms_per_step"], d["value"], d["e2e"]["value"]))
PY
timeout 600 python tools/layer_bench.py --only gemm --impl x3,x1 --reps 10 > $O/layer_gemm.txt 2>&1
cat $O/layer_gemm.txt | cut -c1-140
for L in gemm:hz_lin.wgrad:x1 gemm:h4_lin.wgrad:x1 gemm:hz_lin.dgrad:x1 gemm:hz_lin:x3 gemm:h4_lin:x3; do
  K=$(echo $L | cut -d: -f1,2); I=$(echo $L | cut -d: -f3); N=$(echo $L | tr ':' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_igemm_kernel -c 1 -o $O/prof_$N -f \
    python tools/layer_bench.py --batch 256 --one $K --impl $I --reps 1 > $O/ncu_$N.log 2>&1
  tail -1 $O/ncu_$N.log | cut -c1-200
done
for L in fprop:h0 transpose:d_h4; do
  N=$(echo $L | tr ':' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:h16_kernel -c 1 -o $O/prof_inc2_$N -f \
    python tools/layer_bench.py --model inception2 --batch 256 --one $L --impl p3 --reps 1 > $O/ncu_inc2_$N.log 2>&1
  tail -1 $O/ncu_inc2_$N.log | cut -c1-200
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:few_transpose_kernel -c 1 -o $O/prof_few_transpose -f \
    python tools/layer_bench.py --batch 256 --one "transpose:d_h4" --impl p3 --reps 1 > $O/ncu_few_transpose.log 2>&1
tail -2 $O/ncu_few_transpose.log | cut -c1-200
ls $O
