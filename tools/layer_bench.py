"""Per-layer timing of the GEMM-shaped ops of the bench workload (ContextSkipNew 64x64x3, B triples), each op in
isolation: CUDA events around `reps` back-to-back launches after warm-up.  Prints TFLOP/s (algorithmic flops).

    python tools/layer_bench.py [--batch 256] [--impl x3,x1] [--only fprop,transpose,wgrad,gemm] [--reps 10]
    python tools/layer_bench.py --one fprop:h2 --impl x3 --reps 3      # a short ncu target
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imitation_from_observation_b200 import ops                                   # noqa: E402
from imitation_from_observation_b200._lib import IMPL_AUTO, IMPL_FAST, IMPL_SIMT  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--impl", default="x3,x1")
ap.add_argument("--only", default="")
ap.add_argument("--one", default="")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--model", default="skipnew", choices=["skipnew", "inception2"],
                help="skipnew = the bench workload (ContextSkipNew 64x64x3); inception2 = ContextAEInception2 on 8x8x2048 "
                     "Mixed_7c maps, 3x3 kernels, filters 1024/1024/512/512 (base.py:126)")
a = ap.parse_args()
B = a.batch
dev = "cuda"
# bf3 = bf16 (hi, lo) planes, three products (round 1's forward); p3 = fp16 (hi, lo) planes, three products; p1 = fp16 hi
# planes, one product (the backward pass: TF32-grade operands rounded to nearest)
IMPLS = {"x3": IMPL_AUTO, "x1": IMPL_FAST, "simt": IMPL_SIMT, "bf3": "bf3", "p3": "p3", "p1": "p1"}
PL = {"bf3": (torch.bfloat16, 3, 2), "p3": (torch.float16, 3, 2), "p1": (torch.float16, 1, 1)}   # dtype, nprod, planes


def planes_of(x, tag, transpose=False):
    dt, _, npl = PL[tag]
    return (ops.split_planes_transpose if transpose else ops.split_planes)(x, nplanes=npl, dtype=dt)


def rnd(*shape):
    return torch.randn(shape, device=dev) * 0.1


cases = []   # (kind, name, flops, fn(impl))


def add_fprop(name, N, H, C, Co, k=5, st=2):
    x, w, b = rnd(N, H, H, C), rnd(k, k, C, Co), rnd(Co)
    P = -(-H // st)
    y = torch.empty((N, P, P, Co), device=dev)
    bf = ops.conv2d_planes_eligible(N, (H, H), C, Co, k, st)
    few = ops.conv2d_few_eligible(N, (H, H), C, Co, k, st)
    cache = {}

    def run(impl):
        if impl in PL and few and PL[impl][0] == torch.float16:
            if impl not in cache:
                cache[impl] = (ops.conv2d_few_pack(w, PL[impl][2]), torch.empty((PL[impl][2], N, P, P, Co), device=dev,
                                                                                dtype=torch.float16))
            wimg, yp = cache[impl]
            return ops.conv2d_few(x, wimg, k, Co, bias=b, stride=st, act=True, out=y, out_planes=yp, nprod=PL[impl][1])
        if impl in PL:
            if not bf:
                return None
            if impl not in cache:
                cache[impl] = (planes_of(x, impl), planes_of(w, impl, True),
                               torch.empty((PL[impl][2], N, P, P, Co), device=dev, dtype=PL[impl][0]))
            xp, wp, yp = cache[impl]
            return ops.conv2d_planes(xp, wp, k, bias=b, stride=st, act=True, out=y, out_planes=yp, nprod=PL[impl][1])
        return ops.conv2d(x, w, b, stride=st, act=True, out=y, impl=impl)
    cases.append(("fprop", name, 2.0 * N * P * P * Co * k * k * C, run))


def add_im2col(name, N, H, C, k=5, st=2):
    x = rnd(N, H, H, C)
    ld = -(-(C * k * k) // 4) * 4
    P = -(-H // st)
    out = torch.empty((N * P * P, ld), device=dev)
    cases.append(("im2col", name, float(N * P * P * ld), lambda impl: ops.im2col(x, k, st, ld, out=out)))


def add_trans(name, N, P, C1, C2, NB, Cn, k=5, st=2, act=True):
    x1 = rnd(N, P, P, C1)
    x2 = rnd(NB, P, P, C2) if C2 else None
    w, b = rnd(k, k, Cn, C1 + C2), rnd(Cn)
    y = torch.empty((N, P * st, P * st, Cn), device=dev)
    bf = ops.conv2d_transpose_planes_eligible(N, (P * st, P * st), C1, C2, NB, Cn, k, st)
    merged = os.environ.get("CTX_BF3_MERGE", "1") != "0" and \
        ops.conv2d_transpose_merged_eligible(N, (P * st, P * st), C1, C2, NB, Cn, k, st)
    cache = {}
    fewt = ops.conv2d_transpose_few_eligible(N, (P * st, P * st), C1, C2, NB, Cn, k, st)

    def run(impl):
        if impl == "p3" and fewt:
            if impl not in cache:
                cache[impl] = (planes_of(x1, impl), planes_of(x2, impl) if C2 else None, ops.conv2d_transpose_few_pack(w))
            return ops.conv2d_transpose_few(cache[impl][0], cache[impl][2], (P * st, P * st), k, Cn, bias=b, stride=st,
                                            x2p=cache[impl][1], act=act, out=y)
        if impl in PL:
            if not (bf or merged):
                return None
            dt, nprod, npl = PL[impl]
            if impl not in cache:
                cache[impl] = (planes_of(x1, impl), planes_of(x2, impl) if C2 else None,
                               ops.conv2d_transpose_merged_pack(w, (P * st, P * st), st, nplanes=npl, dtype=dt) if merged
                               else planes_of(w, impl).view(npl, k * k * Cn, C1 + C2),
                               torch.empty((npl, N, P * st, P * st, Cn), device=dev, dtype=dt))
            x1p, x2p, wp, yp = cache[impl]
            if merged:
                return ops.conv2d_transpose_merged_planes(x1p, wp, (P * st, P * st), k, Cn, bias=b, stride=st, x2p=x2p,
                                                          act=act, out=y, out_planes=yp, nprod=nprod)
            return ops.conv2d_transpose_planes(x1p, wp, (P * st, P * st), k, bias=b, stride=st, x2p=x2p, act=act, out=y,
                                               out_planes=yp, nprod=nprod)
        return ops.conv2d_transpose(x1, w, (P * st, P * st), bias=b, stride=st, x2=x2, act=act, out=y, impl=impl)
    cases.append(("transpose", name, 2.0 * N * P * P * Cn * k * k * (C1 + C2), run))


def add_wgrad(name, N, H, Ca, Cb1, Cb2, NB, k=5, st=2):
    P = -(-H // st)
    big, s1 = rnd(N, H, H, Ca), rnd(N, P, P, Cb1)
    s2 = rnd(NB, P, P, Cb2) if Cb2 else None
    dw = torch.zeros((k, k, Ca, Cb1 + Cb2), device=dev)
    ok = ops.conv2d_wgrad_planes_eligible(N, (H, H), Ca, Cb1, Cb2, NB, k, st)
    few = ops.conv2d_few_eligible(N, (H, H), Ca, Cb1 + Cb2, k, st)
    cache = {}

    def run(impl):
        if impl == "p1" and few:
            if impl not in cache:
                cache[impl] = (planes_of(s1, impl), planes_of(s2, impl) if Cb2 else None)
            return ops.conv2d_few_wgrad(big, cache[impl][0], k, st, small2p=cache[impl][1], dw=dw)
        if impl in PL:
            if not ok:
                return None
            if impl not in cache:
                cache[impl] = (planes_of(big, impl), planes_of(s1, impl), planes_of(s2, impl) if Cb2 else None)
            bp, s1p, s2p = cache[impl]
            return ops.conv2d_wgrad_planes(bp, s1p, k, st, small2p=s2p, dw=dw, nprod=PL[impl][1])
        return ops.conv2d_wgrad(big, s1, k, st, small2=s2, dw=dw, impl=impl)
    cases.append(("wgrad", name, 2.0 * N * P * P * (Cb1 + Cb2) * k * k * Ca, run))


def add_gemm(name, M, N, K, tA=False, tB=False, bias=True, accumulate=False, pad=0):
    if pad:            # patch-matrix operand: K (or M) columns of a wider row
        A = rnd(*((K, M + pad) if tA else (M, K + pad)))[:, :(M if tA else K)]
    else:
        A = rnd(*((K, M) if tA else (M, K)))
    Bm = rnd(*((N, K) if tB else (K, N)))
    bv = rnd(N) if bias else None
    Cm = torch.zeros((M, N), device=dev)
    cases.append(("gemm", name, 2.0 * M * N * K,
                  lambda impl: ops.gemm(A, Bm, out=Cm, transA=tA, transB=tB, bias=bv, act=bias, accumulate=accumulate,
                                        impl=impl)))


E = 2 * B   # images through scope "conv" (src, tgt); ctx encoder has B
D = 2 * B   # decoder passes
if a.model == "inception2":
    # encoder (scope "conv", 2B images): 3x3, strides 1/2/1/2 on 8x8x2048
    add_fprop("h0", E, 8, 2048, 1024, k=3, st=1)
    add_fprop("h1", E, 8, 1024, 1024, k=3, st=2)
    add_fprop("h2", E, 4, 1024, 512, k=3, st=1)
    add_fprop("h3", E, 4, 512, 512, k=3, st=2)
    # decoder (2B passes, skips from the B context images)
    add_trans("d_h1", D, 2, 512, 512, B, 512, k=3, st=2)
    add_trans("d_h2", D, 4, 512, 512, B, 1024, k=3, st=1)
    add_trans("d_h3", D, 4, 1024, 1024, B, 1024, k=3, st=2)
    add_trans("d_h4", D, 8, 1024, 1024, B, 2048, k=3, st=1, act=False)
    # backward of the two big layers
    add_wgrad("d_h4.wgrad", D, 8, 2048, 1024, 1024, B, k=3, st=1)
    add_fprop("d_h4.dgrad", D, 8, 2048, 2048, k=3, st=1)
    add_wgrad("h0.wgrad", E, 8, 2048, 1024, 0, 0, k=3, st=1)
    add_trans("h0.dgrad", E, 8, 1024, 0, 0, 2048, k=3, st=1)
    cases_done = True
else:
    cases_done = False
# forward
if not cases_done:
  add_fprop("h0 (3->64 @64)", E, 64, 3, 64)
  add_fprop("d_h4.dgrad (3->128 @64)", D, 64, 3, 128)
  add_wgrad("h0.wgrad (few)", E, 64, 3, 64, 0, 0)
  add_wgrad("d_h4.wgrad (few)", D, 64, 3, 64, 64, B)
  add_im2col("h0.im2col", E, 64, 3)
  add_gemm("h0.gemm (K=75)", E * 1024, 64, 75, pad=1)
  add_fprop("h1", E, 32, 64, 128)
  add_fprop("h2", E, 16, 128, 256)
  add_fprop("h3", E, 8, 256, 512)
  add_gemm("h4_lin", E, 1024, 8192)
  add_gemm("hz_lin", B, 1024, 1024)
  add_gemm("d_h0_lin", D, 8192, 1024)
  add_trans("d_h1", D, 4, 512, 512, B, 256)
  add_trans("d_h2", D, 8, 256, 256, B, 128)
  add_trans("d_h3", D, 16, 128, 128, B, 64)
  add_trans("d_h4 (128->3 @64)", D, 32, 64, 64, B, 3, act=False)
  # backward: decoder
  add_gemm("d_h4.wgrad (M=75)", 75, 64, D * 1024, tA=True, bias=False, accumulate=True, pad=1)
  add_gemm("d_h4.dgrad (K=75)", D * 1024, 128, 75, bias=False, pad=1)
  add_wgrad("d_h3.wgrad", D, 32, 64, 128, 128, B)
  add_fprop("d_h3.dgrad", D, 32, 64, 256)
  add_wgrad("d_h2.wgrad", D, 16, 128, 256, 256, B)
  add_fprop("d_h2.dgrad", D, 16, 128, 512)
  add_wgrad("d_h1.wgrad", D, 8, 256, 512, 512, B)
  add_fprop("d_h1.dgrad", D, 8, 256, 1024)
  add_gemm("d_h0_lin.wgrad", 1024, 8192, D, tA=True, bias=False, accumulate=True)
  add_gemm("d_h0_lin.dgrad", D, 1024, 8192, tB=True, bias=False)
  # backward: encoder
  add_gemm("h4_lin.wgrad", 8192, 1024, E, tA=True, bias=False, accumulate=True)
  add_gemm("h4_lin.dgrad", E, 8192, 1024, tB=True, bias=False)
  add_gemm("hz_lin.wgrad", 1024, 1024, B, tA=True, bias=False, accumulate=True)
  add_gemm("hz_lin.dgrad", B, 1024, 1024, tB=True, bias=False)
  add_wgrad("h3.wgrad", E, 8, 256, 512, 0, 0)
  add_trans("h3.dgrad", E, 4, 512, 0, 0, 256)
  add_wgrad("h2.wgrad", E, 16, 128, 256, 0, 0)
  add_trans("h2.dgrad", E, 8, 256, 0, 0, 128)
  add_wgrad("h1.wgrad", E, 32, 64, 128, 0, 0)
  add_trans("h1.dgrad", E, 16, 128, 0, 0, 64)
  add_gemm("h0.wgrad (M=75)", 75, 64, E * 1024, tA=True, bias=False, accumulate=True, pad=1)

only = set(x for x in a.only.split(",") if x)
tot = {}
for kind, name, flops, fn in cases:
    if only and kind not in only:
        continue
    if a.one and a.one != "%s:%s" % (kind, name.split(" ")[0]):
        continue
    line = "%-10s %-22s %8.2f GF " % (kind, name, flops / 1e9)
    for tag in a.impl.split(","):
        impl = IMPLS[tag]
        if impl in PL and (kind not in ("fprop", "transpose", "wgrad") or fn(impl) is None):
            line += "| %-3s      n/a                    " % tag
            continue
        for _ in range(2):
            fn(impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn(impl)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        used = {1: "simt", 2: "tc1", 3: "tc3", 5: "p3", 6: "p1"}.get(int(ops._lib.load().ctx_last_impl()), "?")
        tot[tag] = tot.get(tag, 0.0) + ms
        line += "| %s(%s) %8.1f us %7.1f TF/s " % (tag, used, ms * 1e3, flops / ms / 1e9)
    print(line, flush=True)
print("total ms: " + ", ".join("%s %.2f" % kv for kv in tot.items()))
