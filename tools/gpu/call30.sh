#!/bin/bash
# GPU call 30 (round 2): catch the intermittent bench stall with its Python stacks (faulthandler on SIGTERM)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c30
mkdir -p $O
for i in $(seq 1 24); do
  timeout -s TERM -k 10 40 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$i.json 2> $O/bench_$i.err
  rc=$?
  echo "run $i rc=$rc $(date +%T)"
  if [ $rc -ne 0 ]; then
    echo "== stalled run $i"; head -120 $O/bench_$i.err | cut -c1-220
    nvidia-smi --query-gpu=utilization.gpu,clocks.sm,power.draw --format=csv,noheader
    break
  fi
done
