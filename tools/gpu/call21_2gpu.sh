#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c21
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
CTX_DP_GRAPH=0 timeout 120 $TR --nproc-per-node 2 --master-port 29511 tools/dp_check.py > $O/dp_check_eager.txt 2>&1; echo "rc=$?"
tail -3 $O/dp_check_eager.txt | cut -c1-200
CTX_GRAPH=0 timeout 120 $TR --nproc-per-node 2 --master-port 29512 tools/dp_check.py > $O/dp_check_nograph_at_all.txt 2>&1; echo "rc=$?"
tail -3 $O/dp_check_nograph_at_all.txt | cut -c1-200
