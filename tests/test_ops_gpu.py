"""GPU parity tests, op by op: every C-ABI entry point against the CPU oracle on seeded inputs.
Tolerances: exact-fp32 SIMT path 2e-5 (summation order only); tcgen05 TF32 path 1e-3 (BASELINE.json north_star);
integer results bit-exact.  Error metric: ||a-b||_inf / ||b||_inf."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from imitation_from_observation_b200 import _lib, ops
from imitation_from_observation_b200._lib import IMPL_AUTO, IMPL_FAST, IMPL_SIMT
from oracle import tf_semantics as O

pytestmark = pytest.mark.gpu
# exact-fp32 SIMT and the 3-pass TF32 split (IMPL_AUTO) are both held to fp32 grade; the single-pass TF32 mode
# (IMPL_FAST) is fed unrounded operands here, so it sees the hardware's truncation (2^-10 per operand)
FP32_TOL, TF32_TOL = 2e-5, 3e-3


def _impls():
    return [(IMPL_SIMT, FP32_TOL), (IMPL_AUTO, FP32_TOL), (IMPL_FAST, TF32_TOL)]


def g(seed):
    return torch.Generator().manual_seed(seed)


def rnd(shape, seed, scale=1.0):
    return torch.randn(shape, generator=g(seed)) * scale


# (N, H, W, C, Co, k, stride): reference layer shapes (scaled down in batch), odd sizes, 3x3/stride-2 asymmetric pad
FPROP_CASES = [
    (2, 64, 64, 3, 64, 5, 2), (2, 32, 32, 64, 128, 5, 2), (3, 16, 16, 128, 256, 5, 2), (4, 8, 8, 256, 512, 5, 2),
    (2, 48, 48, 3, 64, 5, 2), (2, 6, 6, 256, 512, 5, 2), (5, 3, 3, 64, 32, 5, 2),
    (2, 36, 64, 3, 32, 5, 1), (2, 18, 32, 16, 8, 5, 2), (2, 8, 8, 64, 32, 3, 1), (2, 8, 8, 32, 32, 3, 2),
    (1, 7, 9, 5, 7, 3, 2), (1, 1, 1, 4, 4, 5, 2), (2, 5, 5, 8, 130, 1, 1),
    # odd extents with tensor-core channel counts (ContextAEInception on Mixed_5d: 35 -> 18 -> 9 -> 5 -> 3)
    (2, 35, 35, 288, 64, 5, 2), (3, 9, 9, 128, 256, 5, 2), (2, 5, 5, 256, 512, 5, 2), (2, 17, 13, 64, 96, 3, 2),
]


@pytest.mark.parametrize("N,H,W,C,Co,k,st", FPROP_CASES)
@pytest.mark.parametrize("impl,tol", _impls())
def test_conv2d_fprop(N, H, W, C, Co, k, st, impl, tol):
    x, w, b = rnd((N, H, W, C), 1), rnd((k, k, C, Co), 2, (k * k * C) ** -0.5), rnd((Co,), 3, 0.1)
    ref = O.lrelu(O.conv2d(x.double(), w.double(), b.double(), st))
    y = ops.conv2d(x.cuda(), w.cuda(), b.cuda(), stride=st, act=True, impl=impl)
    assert tuple(y.shape) == tuple(ref.shape)
    assert rel_err(y.cpu(), ref) < tol
    ref2 = O.conv2d(x.double(), w.double(), None, st)
    y2 = ops.conv2d(x.cuda(), w.cuda(), None, stride=st, act=False, impl=impl)
    assert rel_err(y2.cpu(), ref2) < tol


# (N, P, Q -> H, W, C1, C2, skip batch, Cn, k, stride)
TRANS_CASES = [
    (4, 4, 4, 8, 8, 512, 512, 2, 256, 5, 2), (4, 8, 8, 16, 16, 256, 256, 2, 128, 5, 2),
    (2, 16, 16, 32, 32, 128, 128, 1, 64, 5, 2), (2, 32, 32, 64, 64, 64, 64, 2, 3, 5, 2),
    (4, 3, 3, 6, 6, 64, 64, 2, 32, 5, 2), (2, 9, 16, 18, 32, 8, 8, 2, 16, 5, 2),
    (2, 18, 32, 18, 32, 16, 16, 2, 16, 5, 1), (2, 36, 64, 36, 64, 32, 32, 1, 3, 5, 1),
    (2, 4, 4, 8, 8, 32, 0, 0, 64, 3, 2), (2, 8, 8, 8, 8, 64, 64, 2, 32, 3, 1), (2, 2, 2, 3, 3, 8, 8, 2, 4, 3, 2),
    (3, 4, 5, 7, 9, 6, 0, 0, 5, 3, 2), (2, 4, 4, 8, 8, 4, 4, 2, 8, 1, 2),
    # odd output extents on the tensor-core path (stride phases of unequal size)
    (2, 3, 3, 5, 5, 512, 512, 2, 256, 5, 2), (4, 5, 5, 9, 9, 256, 256, 2, 128, 5, 2),
    (2, 18, 18, 35, 35, 64, 64, 1, 288, 5, 2), (2, 9, 7, 17, 13, 96, 32, 2, 64, 3, 2),
]


@pytest.mark.parametrize("N,P,Q,H,W,C1,C2,NB,Cn,k,st", TRANS_CASES)
@pytest.mark.parametrize("impl,tol", _impls())
def test_conv2d_transpose_dual_source(N, P, Q, H, W, C1, C2, NB, Cn, k, st, impl, tol):
    x1 = rnd((N, P, Q, C1), 4)
    x2 = rnd((NB, P, Q, C2), 5) if C2 else None
    w = rnd((k, k, Cn, C1 + C2), 6, (k * k * (C1 + C2) / st ** 2) ** -0.5)
    b = rnd((Cn,), 7, 0.1)
    cat = x1 if x2 is None else torch.cat([x1, x2.repeat((N // NB, 1, 1, 1))], 3)      # image n reads skip n % NB
    ref = O.lrelu(O.conv2d_transpose(cat.double(), w.double(), b.double(), (H, W), st))
    y = ops.conv2d_transpose(x1.cuda(), w.cuda(), (H, W), bias=b.cuda(), stride=st,
                             x2=None if x2 is None else x2.cuda(), act=True, impl=impl)
    assert rel_err(y.cpu(), ref) < tol


# few output channels (the image side): one GEMM over (tap, channel) columns + scatter-add epilogue on the tcgen05 path
@pytest.mark.parametrize("N,P,Q,H,W,C1,C2,NB,Cn,k,st,expect_tc", [
    (2, 32, 32, 64, 64, 64, 64, 2, 3, 5, 2, True), (4, 16, 16, 32, 32, 32, 0, 0, 3, 5, 2, True),
    (2, 8, 8, 8, 8, 64, 64, 2, 8, 3, 1, True), (3, 4, 4, 8, 8, 32, 32, 3, 5, 5, 2, False),
    (2, 24, 24, 48, 48, 64, 64, 2, 3, 5, 2, False)])
@pytest.mark.parametrize("impl,tol", [(IMPL_AUTO, FP32_TOL), (IMPL_FAST, TF32_TOL)])
def test_conv2d_transpose_few_channels(N, P, Q, H, W, C1, C2, NB, Cn, k, st, expect_tc, impl, tol):
    x1 = rnd((N, P, Q, C1), 40)
    x2 = rnd((NB, P, Q, C2), 41) if C2 else None
    w = rnd((k, k, Cn, C1 + C2), 42, (k * k * (C1 + C2) / st ** 2) ** -0.5)
    b = rnd((Cn,), 43, 0.1)
    cat = x1 if x2 is None else torch.cat([x1, x2.repeat((N // NB, 1, 1, 1))], 3)
    ref = O.conv2d_transpose(cat.double(), w.double(), b.double(), (H, W), st)
    y = ops.conv2d_transpose(x1.cuda(), w.cuda(), (H, W), bias=b.cuda(), stride=st,
                             x2=None if x2 is None else x2.cuda(), act=False, impl=impl)
    used = int(_lib.load().ctx_last_impl())
    assert rel_err(y.cpu(), ref) < tol
    if expect_tc and ops.tc_available():
        assert used != IMPL_SIMT, "the few-channel transposed conv should take the tcgen05 scatter path"
    prev = rnd((N, H, W, Cn), 44)
    out = prev.clone().cuda()
    ops.conv2d_transpose(x1.cuda(), w.cuda(), (H, W), bias=b.cuda(), stride=st,
                         x2=None if x2 is None else x2.cuda(), act=False, out=out, accumulate=True, impl=impl)
    assert rel_err(out.cpu(), ref + prev.double()) < tol


@pytest.mark.parametrize("N,H,W,C,k,st,ld", [(2, 64, 64, 3, 5, 2, 76), (3, 7, 9, 5, 3, 2, 48), (2, 36, 64, 3, 5, 1, 80),
                                             (1, 6, 6, 8, 3, 1, 72)])
def test_im2col_bit_exact_and_gemm_equals_conv(N, H, W, C, k, st, ld):
    x = rnd((N, H, W, C), 45)
    pt = ops.im2col(x.cuda(), k, st, ld).cpu()
    Pn, Qn = -(-H // st), -(-W // st)
    lo_h, lo_w = O._same(H, k, st)[1], O._same(W, k, st)[1]
    ref = torch.zeros((N, Pn, Qn, ld))
    xp = torch.zeros((N, H + 2 * k, W + 2 * k, C))
    xp[:, k:k + H, k:k + W] = x
    for r in range(k):
        for s_ in range(k):
            h0, w0 = k + r - lo_h, k + s_ - lo_w
            ref[..., (r * k + s_) * C:(r * k + s_ + 1) * C] = xp[:, h0:h0 + Pn * st:st, w0:w0 + Qn * st:st][:, :Pn, :Qn]
    assert torch.equal(pt, ref.view(-1, ld))                   # a gather: bit-exact
    # the few-channel conv as a dense GEMM over the patch matrix (K = k*k*C need not be a multiple of 4 or 32)
    Co, K = 64, k * k * C
    w, b = rnd((k, k, C, Co), 46, K ** -0.5), rnd((Co,), 47, 0.1)
    refc = O.lrelu(O.conv2d(x.double(), w.double(), b.double(), st))
    for impl, tol in _impls():
        y = ops.gemm(pt.cuda()[:, :K], w.cuda().view(K, Co), bias=b.cuda(), act=True, impl=impl)
        assert rel_err(y.cpu().view(refc.shape), refc) < tol
        # ... and its filter gradient: patches^T @ dy (M = K rows, not a multiple of 32)
        dy = rnd((N * Pn * Qn, Co), 48)
        dw = torch.zeros((K, Co), device="cuda")
        ops.gemm(pt.cuda()[:, :K], dy.cuda(), out=dw, transA=True, accumulate=True, impl=impl)
        assert rel_err(dw.cpu(), pt[:, :K].double().t() @ dy.double()) < tol


@pytest.mark.parametrize("impl,tol", _impls())
def test_conv2d_transpose_backward_epilogue(impl, tol):
    """dgrad with fused lrelu' mask and accumulation (what the encoder backward uses)."""
    N, P, Q, H, W, Ck, Cn, k, st = 3, 8, 8, 16, 16, 64, 32, 5, 2
    dy, w = rnd((N, P, Q, Ck), 8), rnd((k, k, Cn, Ck), 9, 0.05)
    yref, prev = rnd((N, H, W, Cn), 10), rnd((N, H, W, Cn), 11)
    ref = O.conv2d_transpose(dy.double(), w.double(), None, (H, W), st) * torch.where(yref >= 0, 1.0, 0.2).double() + prev.double()
    out = prev.clone().cuda()
    ops.conv2d_transpose(dy.cuda(), w.cuda(), (H, W), stride=st, out=out, act_ref=yref.cuda(), accumulate=True, impl=impl)
    assert rel_err(out.cpu(), ref) < tol


# (N, H, W, Ca, Cb1, Cb2, NB, k, stride)
WGRAD_CASES = [
    (4, 64, 64, 3, 64, 0, 0, 5, 2), (4, 16, 16, 128, 256, 0, 0, 5, 2), (8, 8, 8, 256, 512, 0, 0, 5, 2),
    (4, 8, 8, 256, 512, 512, 2, 5, 2), (4, 64, 64, 3, 64, 64, 2, 5, 2), (2, 36, 64, 3, 32, 32, 1, 5, 1),
    (2, 8, 8, 32, 16, 16, 2, 3, 2), (3, 7, 9, 5, 4, 0, 0, 3, 2), (2, 6, 6, 64, 32, 32, 2, 5, 2),
    (2, 35, 35, 288, 64, 64, 2, 5, 2), (4, 9, 9, 128, 256, 256, 2, 5, 2), (2, 17, 13, 64, 96, 32, 2, 3, 2),
]


@pytest.mark.parametrize("N,H,W,Ca,Cb1,Cb2,NB,k,st", WGRAD_CASES)
@pytest.mark.parametrize("impl,tol", _impls())
def test_conv2d_wgrad(N, H, W, Ca, Cb1, Cb2, NB, k, st, impl, tol):
    P, Q = -(-H // st), -(-W // st)
    big = rnd((N, H, W, Ca), 12)
    s1 = rnd((N, P, Q, Cb1), 13)
    s2 = rnd((NB, P, Q, Cb2), 14) if Cb2 else None
    small = s1 if s2 is None else torch.cat([s1, s2.repeat((N // NB, 1, 1, 1))], 3)
    wr = torch.zeros((k, k, Ca, Cb1 + Cb2), dtype=torch.float64, requires_grad=True)
    (O.conv2d(big.double(), wr, None, st) * small.double()).sum().backward()
    base = rnd((k, k, Ca, Cb1 + Cb2), 15)                    # "+=" semantics: accumulates into existing contents
    dw = base.clone().cuda()
    ops.conv2d_wgrad(big.cuda(), s1.cuda(), k, st, small2=None if s2 is None else s2.cuda(), dw=dw, impl=impl)
    assert rel_err(dw.cpu().double() - base.double(), wr.grad) < tol


GEMM_CASES = [(64, 1024, 8192), (50, 100, 1152), (512, 8192, 1024), (3, 16, 16), (1, 1, 1), (130, 70, 33),
              (256, 1024, 2048)]


@pytest.mark.parametrize("M,N,K", GEMM_CASES)
@pytest.mark.parametrize("tA,tB", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("impl,tol", _impls())
def test_gemm_all_transposes(M, N, K, tA, tB, impl, tol):
    A = rnd((K, M) if tA else (M, K), 16)
    B = rnd((N, K) if tB else (K, N), 17, K ** -0.5)
    bias = rnd((N,), 18, 0.1)
    ref = O.lrelu((A.double().t() if tA else A.double()) @ (B.double().t() if tB else B.double()) + bias.double())
    C = ops.gemm(A.cuda(), B.cuda(), transA=tA, transB=tB, bias=bias.cuda(), act=True, impl=impl)
    assert rel_err(C.cpu(), ref) < tol


@pytest.mark.parametrize("impl,tol", _impls())
def test_gemm_leading_dimensions_and_backward_epilogue(impl, tol):
    """Column-sliced operands/outputs (the [src_z ‖ ctx_z] buffer) + lrelu' mask + accumulate."""
    M, N, K = 32, 64, 64
    Abig, Bm = rnd((M, 2 * K), 19), rnd((K, N), 20, 0.1)
    Cbig = rnd((M, 2 * N), 21)
    ref_full = Cbig.double().clone()
    yref = rnd((M, 2 * N), 22)
    acc = Abig[:, K:].double() @ Bm.double()
    ref_full[:, N:] = acc * torch.where(yref[:, N:] >= 0, 1.0, 0.2).double() + Cbig[:, N:].double()
    A_d, C_d, y_d = Abig.cuda(), Cbig.clone().cuda(), yref.cuda()
    epi = _lib.Epilogue(None, y_d.data_ptr() + N * 4, 0.2, 0, 1, 0)
    ops.gemm_raw(False, False, M, N, K, A_d.data_ptr() + K * 4, 2 * K, Bm.cuda().data_ptr(), N,
                 C_d.data_ptr() + N * 4, 2 * N, epi, impl)
    assert rel_err(C_d.cpu(), ref_full) < tol
    assert torch.equal(C_d.cpu()[:, :N], Cbig[:, :N])         # untouched columns stay bit-identical


_CLUSTER_K_SCRIPT = r"""
import sys, torch
sys.path.insert(0, %r)
from imitation_from_observation_b200 import ops, _lib
from imitation_from_observation_b200._lib import IMPL_AUTO, IMPL_FAST
g = torch.Generator().manual_seed(27)
M, N, K = 256, 1024, 1024
A = torch.randn((M, K), generator=g); Bm = torch.randn((N, K), generator=g) * K ** -0.5; bias = torch.randn((N,), generator=g) * 0.1
yref = torch.randn((M, 2 * N), generator=g); base = torch.randn((M, 2 * N), generator=g)
for impl, tol in ((IMPL_AUTO, 2e-5), (IMPL_FAST, 3e-3)):
    C_d, y_d = base.clone().cuda(), yref.cuda()
    ref = base.double().clone()
    ref[:, N:] = (A.double() @ Bm.double().t()) * torch.where(yref[:, N:] >= 0, 1.0, 0.2).double()
    epi = _lib.Epilogue(None, y_d.data_ptr() + N * 4, 0.2, 0, 0, 0)
    ops.gemm_raw(False, True, M, N, K, A.cuda().data_ptr(), K, Bm.cuda().data_ptr(), K, C_d.data_ptr() + N * 4, 2 * N, epi, impl)
    err = float((C_d.cpu().double() - ref).abs().max() / ref.abs().max())
    assert err < tol, ("mask epilogue", impl, err)
    assert torch.equal(C_d.cpu()[:, :N], base[:, :N])
    outs = [ops.gemm(A.cuda(), Bm.cuda(), transB=True, bias=bias.cuda(), act=True, impl=impl).cpu() for _ in range(3)]
    y = A.double() @ Bm.double().t() + bias.double()
    y = torch.where(y >= 0, y, 0.2 * y)
    err = float((outs[0].double() - y).abs().max() / y.abs().max())
    assert err < tol, ("bias + lrelu", impl, err)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    for (m, n, k) in ((64, 1024, 8192), (512, 1024, 8192), (256, 1024, 2048)):
        a = torch.randn((m, k), generator=g); b = torch.randn((k, n), generator=g) * k ** -0.5
        c = ops.gemm(a.cuda(), b.cuda(), impl=impl).cpu().double()
        r = a.double() @ b.double()
        err = float((c - r).abs().max() / r.abs().max())
        assert err < tol, (m, n, k, impl, err)
print("cluster-k ok")
"""


def test_gemm_cluster_split_k_epilogue():
    """CTX_TC_CLUSTER_K=1 (off by default: measured slower than the wide-tile grid split): the skinny dense layers split K
    over a thread-block cluster; rank 0 adds the partial sums that arrive through distributed shared memory and runs the fused
    epilogue — bias + lrelu, lrelu' mask from a column-sliced activation into a column-sliced output — and the sum is
    bit-reproducible.  The knob is read once per process, so the check runs in a child process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CTX_TC_CLUSTER_K="1")
    r = subprocess.run([sys.executable, "-c", _CLUSTER_K_SCRIPT % root], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "cluster-k ok" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


def test_u8_normalize_bit_exact():
    x = torch.randint(0, 256, (3, 5, 7, 9, 3), generator=g(23), dtype=torch.uint8)
    for xs in (x, x.reshape(-1)[1:-2]):                        # aligned and unaligned / odd-length paths
        y = ops.u8_normalize(xs.contiguous().cuda())
        assert torch.equal(y.cpu(), O.normalize_u8(xs))
    assert ops.u8_normalize(torch.zeros(0, dtype=torch.uint8).cuda()).numel() == 0


def test_lrelu_bwd_and_bias_grad():
    dy, y = rnd((1000, 37), 24), rnd((1000, 37), 25)
    out = ops.lrelu_bwd(dy.cuda(), y.cuda())
    assert torch.equal(out.cpu(), dy * torch.where(y >= 0, 1.0, 0.2))
    db = torch.ones(37).cuda()
    ops.bias_grad(dy.cuda(), db)
    assert rel_err(db.cpu(), 1 + dy.double().sum(0)) < 1e-5
    big = rnd((70000, 3), 26)
    db = torch.zeros(3).cuda()
    ops.bias_grad(big.cuda(), db)
    assert rel_err(db.cpu(), big.double().sum(0)) < 1e-4


@pytest.mark.parametrize("resid", [False, True])
def test_l2_loss_fwd_bwd(resid):
    B, per = 5, 3 * 16 * 16
    pred, tgt = rnd((2 * B, 16, 16, 3), 27), rnd((B, 16, 16, 3), 28)
    r = rnd((B, 16, 16, 3), 29) if resid else None
    loss = torch.zeros(2).cuda()
    dpred = torch.empty_like(pred).cuda()
    w = torch.tensor([1.0, 0.0]).cuda()
    ops.l2_loss_fwd_bwd(pred.cuda(), tgt.cuda(), loss, dpred=dpred, resid=None if r is None else r.cuda(), weight=w)
    eff = pred.double().view(2, B, -1) + (r.double().view(1, B, -1) if resid else 0)
    d = eff - tgt.double().view(1, B, -1)
    assert rel_err(loss.cpu(), (d ** 2).sum((1, 2)) / 2) < 1e-5
    assert rel_err(dpred.cpu().view(2, B, -1)[0], d[0]) < 1e-6
    assert float(dpred.cpu().view(2, B, -1)[1].abs().max()) == 0.0


def test_mse_scaled_fwd_bwd():
    B, F = 6, 48
    z2 = rnd((2 * B, F), 30)
    gz = rnd((2 * B, F), 31)
    loss = torch.zeros(1).cuda()
    z2d, gzd = z2.cuda(), gz.clone().cuda()
    inv = 1.0 / (B * F)
    ops.mse_scaled_fwd_bwd(z2d[:B], z2d[B:], 1e3, inv, loss=loss, da=gzd[:B], db=gzd[B:], accumulate=True)
    d = (z2[:B] - z2[B:]).double()
    assert rel_err(loss.cpu(), [(d ** 2).mean() * 1e3]) < 1e-5
    gref = gz.double().clone()
    gref[:B] += 2e3 * inv * d
    gref[B:] -= 2e3 * inv * d
    assert rel_err(gzd.cpu(), gref) < 1e-5


def test_split_skip_grad():
    G, B, pix, C1, C2 = 2, 3, 20, 8, 5
    dcat = rnd((G * B, pix, C1 + C2), 32)
    sref, kref = rnd((G * B, pix, C1), 33), rnd((B, pix, C2), 34)
    prev = rnd((B, pix, C2), 35)
    ds = torch.empty((G * B, pix, C1)).cuda()
    dk = prev.clone().cuda()
    ops.split_skip_grad(dcat.cuda(), G, B, C1, C2, ds, dk, state_ref=sref.cuda(), skip_ref=kref.cuda(),
                        accumulate_skip=True)
    assert rel_err(ds.cpu(), dcat[..., :C1] * torch.where(sref >= 0, 1.0, 0.2)) < 1e-6
    ksum = dcat[..., C1:].view(G, B, pix, C2).sum(0)
    assert rel_err(dk.cpu(), prev + ksum * torch.where(kref >= 0, 1.0, 0.2)) < 1e-6


def test_split_skip_grad_vectorised_and_rounded():
    G, B, pix, C1, C2 = 2, 4, 64, 64, 32
    dcat = rnd((G * B, pix, C1 + C2), 52)
    sref, kref = rnd((G * B, pix, C1), 53), rnd((B, pix, C2), 54)
    for rnd_flag in (False, True):
        ds = torch.empty((G * B, pix, C1)).cuda()
        dk = torch.empty((B, pix, C2)).cuda()
        ops.split_skip_grad(dcat.cuda(), G, B, C1, C2, ds, dk, state_ref=sref.cuda(), skip_ref=kref.cuda(),
                            round_tf32=rnd_flag)
        tol = 6e-4 if rnd_flag else 1e-6
        assert rel_err(ds.cpu(), dcat[..., :C1] * torch.where(sref >= 0, 1.0, 0.2)) < tol
        ksum = dcat[..., C1:].view(G, B, pix, C2).sum(0)
        assert rel_err(dk.cpu(), ksum * torch.where(kref >= 0, 1.0, 0.2)) < tol
        if rnd_flag:                                             # low 13 mantissa bits are zero after rounding
            assert int((ds.view(torch.int32) & 0x1FFF).abs().max()) == 0


def test_adam_step_matches_tf_formula():
    n = 10007
    p, gr = rnd((n,), 36), rnd((n,), 37)
    P, G, M, V = {"w": p.clone()}, {"w": gr}, {"w": torch.zeros(n)}, {"w": torch.zeros(n)}
    pd, md, vd, gd = p.clone().cuda(), torch.zeros(n).cuda(), torch.zeros(n).cuda(), gr.cuda()
    import math
    for t in (1, 2, 3):
        O.tf_adam_step(P, G, M, V, t, 1e-3)
        lr_t = 1e-3 * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        ops.adam_step(pd, gd, md, vd, lr_t)
    assert rel_err(pd.cpu(), P["w"]) < 1e-6 and rel_err(md.cpu(), M["w"]) < 1e-6 and rel_err(vd.cpu(), V["w"]) < 1e-6


def test_round_tf32_and_transpose():
    x = rnd((33, 70), 38)
    y = ops.round_tf32(x.cuda()).cpu()
    bits = y.view(torch.int32)
    assert int((bits & 0x1FFF).abs().max()) == 0                  # low 13 mantissa bits cleared
    assert float((y - x).abs().max()) <= float(x.abs().max()) * 2 ** -11
    t = ops.transpose_round(x.cuda(), 1, 33, 70, round_tf32_=False).cpu()
    assert torch.equal(t.view(70, 33), x.t().contiguous())


def test_nn_err_bit_exact():
    B, nlen = 37, 5
    tgt = rnd((B, 8, 8, 3), 39)
    out = tgt[torch.randperm(B, generator=g(40))] + 0.05 * rnd((B, 8, 8, 3), 41)
    err, idx = ops.nn_err(tgt.cuda(), out.contiguous().cuda(), nlen)
    ref_idx = torch.stack([((tgt - out[b:b + 1]) ** 2).reshape(B, -1).mean(1).argmin() for b in range(B)])
    assert torch.equal(idx.cpu(), ref_idx)
    assert int(err.item()) == O.nn_err(tgt, out, nlen)
    tie = torch.zeros((4, 2, 2, 1))
    err, idx = ops.nn_err(tie.cuda(), tie.cuda(), 2)              # all distances equal -> first index, like tf.argmin
    assert idx.cpu().tolist() == [0, 0, 0, 0] and int(err.item()) == 2


def test_reward_kernels():
    paths, T, F = 7, 25, 64
    z, x = rnd((paths * T, F), 42), rnd((paths * T, 6, 6, 3), 43)
    means, mimgs = rnd((T, F), 44), rnd((T, 6, 6, 3), 45)
    base = np.random.RandomState(0).randn(paths, 2 * T + 1)
    rew = torch.from_numpy(base.copy()).cuda()
    cost = torch.empty((paths, T)).cuda()
    ops.reward_ours(z.cuda(), x.cuda(), means.cuda(), mimgs.cuda(), paths, T, 0.3, rewards=rew, cost=cost)
    for p in range(paths):
        ref = O.reward_ours(z[p * T:(p + 1) * T].numpy(), x[p * T:(p + 1) * T].numpy(), means.numpy(), mimgs.numpy(),
                            0.3, base[p])
        assert np.allclose(rew[p].cpu().numpy(), ref, rtol=1e-5, atol=1e-6 * np.abs(ref).max())
        assert np.array_equal(rew[p].cpu().numpy()[0::2], base[p][0::2])          # even entries untouched
    feat, mu = rnd((paths * T, 2, 2, 16), 46), rnd((T, 2, 2, 16), 47)
    sd = rnd((T, 2, 2, 16), 48).abs()
    sd[:, 0, 0, :4] = 0
    rew2 = torch.from_numpy(base.copy()).cuda()
    ops.reward_inception(feat.cuda(), mu.cuda(), sd.cuda(), paths, T, rewards=rew2)
    for p in range(paths):
        ref = O.reward_inception(feat[p * T:(p + 1) * T].numpy(), mu.numpy(), sd.numpy(), base[p])
        assert np.allclose(rew2[p].cpu().numpy(), ref, rtol=1e-5, atol=1e-6 * np.abs(ref).max())


def test_errors_are_reported_not_ub():
    x = torch.zeros((1, 8, 8, 4)).cuda()
    with pytest.raises(_lib.CtxError):
        ops.conv2d(x, torch.zeros((5, 5, 3, 8)).cuda())            # channel mismatch
    with pytest.raises(_lib.CtxError):
        ops.conv2d(x.cpu(), torch.zeros((5, 5, 4, 8)).cuda())      # CPU tensor: no fallback
    with pytest.raises(_lib.CtxError):
        ops.conv2d_transpose(x, torch.zeros((5, 5, 8, 4)).cuda(), (17, 16))   # inconsistent output_shape
    with pytest.raises(_lib.CtxError):
        ops.gemm(torch.zeros((4, 5)).cuda(), torch.zeros((6, 7)).cuda())
    empty = ops.conv2d(torch.zeros((0, 8, 8, 4)).cuda(), torch.zeros((5, 5, 4, 8)).cuda())
    assert tuple(empty.shape) == (0, 4, 4, 8)                      # empty batch is a no-op, not an error


# ---- demo bank in HBM (train_script.py:150-173) ------------------------------------------------------------------------
@pytest.mark.parametrize("frame_shape", [(16, 16, 3), (5, 7, 3), (3,)])
def test_gather_frames_bit_exact(frame_shape):
    rs = np.random.RandomState(0)
    frames = 40
    bank_f = rs.randn(frames, *frame_shape).astype(np.float32)
    bank_u = rs.randint(0, 256, (frames,) + frame_shape).astype(np.uint8)
    lut = rs.randn(256).astype(np.float32)
    idx = rs.randint(0, frames, 33).astype(np.int64)
    di = torch.from_numpy(idx).cuda()
    assert np.array_equal(ops.gather_frames(torch.from_numpy(bank_f).cuda(), di).cpu().numpy(), bank_f[idx])
    got = ops.gather_frames(torch.from_numpy(bank_u).cuda(), di, lut=torch.from_numpy(lut).cuda()).cpu().numpy()
    assert np.array_equal(got, lut[bank_u[idx]])
    # unaligned bank base (a view one element in) takes the scalar path
    flat = torch.from_numpy(bank_f).cuda().reshape(-1)
    elems = int(np.prod(frame_shape))
    shifted = flat[1:1 + (frames - 1) * elems].view((frames - 1,) + frame_shape)
    i2 = torch.from_numpy(idx % (frames - 1)).cuda()
    ref = bank_f.reshape(-1)[1:1 + (frames - 1) * elems].reshape((frames - 1,) + frame_shape)[idx % (frames - 1)]
    assert np.array_equal(ops.gather_frames(shifted.contiguous() if not shifted.is_contiguous() else shifted, i2).cpu().numpy(), ref)
    # an index outside the bank is loud (NaN frame), the others are still right; empty gathers are fine
    bad = idx.copy()
    bad[3] = frames
    out = ops.gather_frames(torch.from_numpy(bank_f).cuda(), torch.from_numpy(bad).cuda()).cpu().numpy()
    assert np.isnan(out[3]).all() and np.array_equal(out[4], bank_f[idx[4]])
    assert ops.gather_frames(torch.from_numpy(bank_f).cuda(), di[:0]).shape == (0,) + frame_shape
    with pytest.raises(_lib.CtxError):
        ops.gather_frames(torch.from_numpy(bank_u).cuda(), di)                    # uint8 bank without its table


def test_device_bank_answers_the_trainer_expressions():
    from imitation_from_observation_b200 import synthetic
    from imitation_from_observation_b200.data import DeviceBank
    nlen, N, B, ntrain = 5, 12, 10, 8
    vdata = synthetic.to_float(synthetic.video_bank(nlen, N, 16, 16, seed=3))      # float64, as train_script.py:97 saves
    bank = DeviceBank(vdata)
    assert bank.quantised and bank.shape == vdata.shape and bank.store.dtype == torch.uint8
    f32 = vdata.astype(np.float32)                                                # what feeding the float bank gives
    rng = np.random.RandomState(1)
    for host, dev in ((f32[:, :ntrain], bank[:, :ntrain]), (f32[:, ntrain:], bank[:, ntrain:])):
        assert dev.shape == host.shape
        choice = rng.choice(host.shape[1], B)
        t_idx = np.arange(0, B) % nlen
        assert np.array_equal(dev[t_idx, choice].cpu().numpy(), host[t_idx, choice])      # train_script.py:156-157
        assert np.array_equal(dev[0, choice].cpu().numpy(), host[0, choice])              # :158 (context frames)
        with pytest.raises(IndexError):
            dev[t_idx, choice + host.shape[1]]
    # a bank that is not made of transform()'s 256 values is kept in float32, still bit-identical
    noisy = vdata + 1e-3 * rng.randn(*vdata.shape)
    b2 = DeviceBank(noisy)
    assert not b2.quantised and b2.store.dtype == torch.float32
    choice = rng.choice(N, B)
    assert np.array_equal(b2[np.arange(B) % nlen, choice].cpu().numpy(), noisy.astype(np.float32)[np.arange(B) % nlen, choice])
    # raw uint8 pixels: gathered through the same table (rescale=True semantics)
    raw = synthetic.video_bank(nlen, N, 16, 16, seed=3)
    assert np.array_equal(DeviceBank(raw)[0, choice].cpu().numpy(), f32[0, choice])


def test_gae_tail_bit_exact():
    """base.py:258-266 (deltas, special.discount_cumsum = scipy.signal.lfilter) on ragged paths, float64, bit for bit."""
    import scipy.signal
    rs = np.random.RandomState(7)
    lens = [50, 1, 25, 50, 7, 50]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    r, b = rs.randn(offs[-1]) * 3.0, rs.randn(offs[-1])
    gamma, lam = 0.99, 0.97
    adv, ret = ops.gae_tail(torch.from_numpy(r).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(offs).cuda(), gamma, lam)

    def discount_cumsum(x, d):
        return scipy.signal.lfilter([1], [1, float(-d)], x[::-1], axis=0)[::-1]
    for i in range(len(lens)):
        s, e = offs[i], offs[i + 1]
        pb = np.append(b[s:e], 0)
        deltas = r[s:e] + gamma * pb[1:] - pb[:-1]
        assert np.array_equal(adv[s:e].cpu().numpy(), discount_cumsum(deltas, gamma * lam))
        assert np.array_equal(ret[s:e].cpu().numpy(), discount_cumsum(r[s:e], gamma))
