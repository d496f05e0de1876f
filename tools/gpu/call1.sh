#!/bin/bash
# GPU call 1 (round 2): 16-bit operand probes, Inception2 full-size parity / layer bench / ncu, bench sanity.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
P=tools/h16_probe
{
echo "# correctness, small shapes: M N K BN stages nprod mn f16 pair splits reps [lbo sbo kstep]"
timeout 120 $P 512 256 640 256 3 1 0 1 0 1 2
timeout 120 $P 512 256 640 256 3 1 0 1 1 1 2
timeout 120 $P 512 256 640 256 2 3 0 1 1 1 2
timeout 120 $P 512 256 640 128 3 3 0 0 0 1 2
echo "# MN-major hypotheses (lbo sbo kstep)"
timeout 120 $P 512 256 640 256 3 1 1 1 0 1 2 8192 1024 2048
timeout 120 $P 512 256 640 256 3 1 1 1 0 1 2 1024 8192 2048
timeout 120 $P 512 256 640 256 3 1 1 1 0 1 2 8192 1024 1024
timeout 120 $P 512 256 640 256 3 1 1 1 0 1 2 8192 2048 2048
timeout 120 $P 512 256 640 256 3 1 1 1 0 1 2 16384 1024 2048
timeout 120 $P 512 256 640 128 3 1 1 1 0 1 2 8192 1024 2048
timeout 120 $P 512 256 640 256 3 1 1 1 1 1 2 8192 1024 2048
timeout 120 $P 512 256 640 256 3 1 1 0 1 1 2 8192 1024 2048
timeout 120 $P 512 256 640 256 2 3 1 1 1 1 2 8192 1024 2048
timeout 120 $P 512 256 1280 256 3 1 1 1 1 4 2 8192 1024 2048
echo "# throughput: dgrad-like K-major (d_h2.dgrad: M=32768 N=512 K=3200)"
timeout 300 $P 32768 512 3200 256 6 1 0 1 1 1 20
timeout 300 $P 32768 512 3200 256 4 1 0 1 1 1 20
timeout 300 $P 32768 512 3200 256 3 1 0 1 1 1 20
timeout 300 $P 32768 512 3200 256 4 1 0 1 0 1 20
timeout 300 $P 32768 512 3200 128 6 1 0 1 0 1 20
timeout 300 $P 32768 512 3200 256 3 3 0 1 1 1 20
timeout 300 $P 32768 512 3200 256 3 3 0 0 1 1 20
echo "# throughput: wgrad-like MN-major (d_h2.wgrad: M=25*128 -> 3328, N=512, K=32768 pixels), split-K"
timeout 300 $P 3328 512 32768 256 6 1 1 1 1 6 20
timeout 300 $P 3328 512 32768 256 6 1 1 1 1 12 20
timeout 300 $P 3328 512 32768 256 4 1 1 1 0 6 20
timeout 300 $P 3328 512 32768 256 3 3 1 1 1 6 20
echo "# h1-like (N=128): M=131072 N=128 K=1600"
timeout 300 $P 131072 128 1600 128 6 1 0 1 0 1 20
timeout 300 $P 131072 128 1600 128 4 3 0 1 0 1 20
} > $O/h16_probe.txt 2>&1
echo "probe done" ; tail -40 $O/h16_probe.txt

timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu -s -k "inception2_full_size or uint8_feeds" > $O/pytest_new.txt 2>&1
tail -15 $O/pytest_new.txt
timeout 600 python -m pytest tests/test_inception_gpu.py -x -q -m gpu -k "tf_checkpoint" > $O/pytest_ckpt.txt 2>&1
tail -3 $O/pytest_ckpt.txt

timeout 600 python tools/layer_bench.py --model inception2 --batch 256 --impl x3,x1,bf3 --reps 5 > $O/layer_inception2_b256.txt 2>&1
cat $O/layer_inception2_b256.txt
timeout 600 python tools/layer_bench.py --model inception2 --batch 32 --impl x3,x1,bf3 --reps 5 > $O/layer_inception2_b32.txt 2>&1
cat $O/layer_inception2_b32.txt

timeout 900 python bench.py --workload inception2 --batch 256 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_inception2_b256.json 2> $O/bench_inception2_b256.err
tail -c 3000 $O/bench_inception2_b256.json; tail -5 $O/bench_inception2_b256.err
timeout 900 python bench.py --workload inception2 --batch 25 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_inception2_b25.json 2> $O/bench_inception2_b25.err
tail -c 1500 $O/bench_inception2_b25.json; tail -5 $O/bench_inception2_b25.err

timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_train.json 2> $O/bench_train.err
tail -c 3000 $O/bench_train.json; tail -5 $O/bench_train.err
timeout 600 python bench.py --batch 32 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_train_b32.json 2> $O/bench_train_b32.err
tail -c 1500 $O/bench_train_b32.json

# ncu --set full: the two big 3x3 layers of Inception2 on the bf16-plane kernel (one launch each)
for L in fprop:h0 transpose:d_h4; do
  N=$(echo $L | tr ':' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:bf3_kernel -c 1 -o $O/prof_inc2_$N -f \
    python tools/layer_bench.py --model inception2 --batch 256 --one $L --impl bf3 --reps 1 > $O/ncu_$N.log 2>&1
  tail -2 $O/ncu_$N.log
done
ls -la $O
