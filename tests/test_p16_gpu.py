"""GPU parity of the 16-bit operand planes path (csrc/tc_h16.cu, csrc/planes.cu) through the C-ABI.

Two bars.  Three-product mode (the forward pass) is held to the fp32 grade of the kernels it replaces: 2e-5 max-norm
against the fp64 oracle.  Single-product mode (fp16 hi planes, the backward pass) multiplies EXACTLY the values its planes
hold, so it is compared with the fp64 oracle evaluated on those rounded operands at the same 2e-5 bar, and against the
unrounded oracle at the TF32 bar (fp16 keeps TF32's 11-bit significand, rounded to nearest: 2^-11 per operand).
The plane producers are integer/bit work and are compared bit for bit with a torch restatement."""
import pytest
import torch

from conftest import rel_err
from imitation_from_observation_b200 import _lib, ops
from oracle import tf_semantics as O

pytestmark = pytest.mark.gpu
TOL, TOL_RN11 = 2e-5, 1.5e-3


def rnd(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def planes_ref(x, dtype, nplanes=2):
    hi = x.clamp(-65504, 65504).to(dtype) if dtype == torch.float16 else x.to(dtype)
    if nplanes == 1:
        return hi[None]
    r = x - hi.float()
    lo = r.clamp(-65504, 65504).to(dtype) if dtype == torch.float16 else r.to(dtype)
    return torch.stack([hi, lo])


def same_bits(a, b):
    return torch.equal(a.cpu().view(torch.int16), b.cpu().view(torch.int16))


def hi(x):                                                      # the value a single-product GEMM multiplies
    return x.half().double()


# ---- producers ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("nplanes", [1, 2])
@pytest.mark.parametrize("shape", [(4, 8, 8, 64), (3, 5, 7), (1001,)])
def test_split_planes_bit_exact(shape, nplanes, dtype):
    x = rnd(shape, 1) * torch.logspace(-6, 3, steps=shape[-1])
    x.view(-1)[0] = 1e6                                         # beyond fp16: saturates to the largest finite value
    p = ops.split_planes(x.cuda(), nplanes=nplanes, dtype=dtype)
    assert p.dtype == dtype and tuple(p.shape) == (nplanes,) + tuple(shape)
    assert same_bits(p, planes_ref(x, dtype, nplanes))


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_split_planes_transpose_layout(dtype):
    w = rnd((3, 3, 40, 96), 2, 0.05)                            # [kh,kw,R,Cc] -> [kh*kw*Cc][R]
    p = ops.split_planes_transpose(w.cuda(), dtype=dtype)
    assert same_bits(p, planes_ref(w.permute(0, 1, 3, 2).reshape(9 * 96, 40).contiguous(), dtype))
    p1 = ops.split_planes_transpose(w.cuda(), nplanes=1, dtype=dtype)
    assert same_bits(p1, planes_ref(w.permute(0, 1, 3, 2).reshape(9 * 96, 40).contiguous(), dtype, 1))


def scale_ref(amax):
    """2^k with amax * 2^k in [2^14, 2^15) (csrc/planes.cu p16_scale_for)."""
    import math
    if amax == 0:
        return 1.0
    return 2.0 ** (14 - math.floor(math.log2(amax)))


@pytest.mark.parametrize("rows,C,ldx,with_ref", [(512, 64, 64, True), (300, 128, 192, True), (37, 256, 256, False),
                                                 (2048, 512, 1024, True)])
def test_grad_planes_bit_exact(rows, C, ldx, with_ref):
    x = rnd((rows, ldx), 3, 3e-4)
    ref = rnd((rows, C), 4) if with_ref else None
    v = x[:, :C] * (torch.where(ref >= 0, 1.0, 0.2) if with_ref else 1.0)
    slot = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.amax(x.cuda(), slot)
    assert slot.cpu().view(torch.float32).item() == x.abs().max().item()
    s = scale_ref(x.abs().max().item())
    scale = torch.zeros(1, device="cuda")
    db = torch.full((C,), 1e-3, device="cuda")               # "+=" semantics
    out32 = torch.empty((rows, C), device="cuda")
    p = torch.empty((2, rows, C), device="cuda", dtype=torch.float16)
    ops.grad_planes(x.cuda(), C, planes=p, ref=None if ref is None else ref.cuda(), amax=slot, scale=scale, db=db,
                    out_f32=out32, ldx=ldx)
    assert scale.item() == s
    assert torch.equal(out32.cpu(), v)
    assert same_bits(p, planes_ref(v * s, torch.float16))
    # (fp32 partial sums + atomics over cancelling terms: the bar is relative to sum |v|)
    assert (db.cpu().double() - 1e-3 - v.double().sum(0)).abs().max() < 1e-5 * v.double().abs().sum(0).max()
    # hi + lo of the scaled gradient carries 22 bits; the hi plane alone 11
    assert ((p[0].float().cpu() / s - v).abs() <= v.abs() * 2.0 ** -11 + 1e-30).all()
    # bf16 planes are never scaled
    pb = torch.empty((1, rows, C), device="cuda", dtype=torch.bfloat16)
    ops.grad_planes(x.cuda(), C, planes=pb, ref=None if ref is None else ref.cuda(), amax=slot, scale=scale, ldx=ldx)
    assert scale.item() == 1.0 and same_bits(pb, planes_ref(v, torch.bfloat16, 1))


def test_skip_grad_sum():
    G, rows, C1, C2 = 2, 96, 64, 128
    dcat, ref, base = rnd((G * rows, C1 + C2), 5), rnd((rows, C2), 6), rnd((rows, C2), 7)
    want = dcat.view(G, rows, C1 + C2)[:, :, C1:].sum(0) * torch.where(ref >= 0, 1.0, 0.2)
    got = ops.skip_grad_sum(dcat.cuda(), G, C1, C2, torch.empty((rows, C2), device="cuda"), skip_ref=ref.cuda())
    assert torch.allclose(got.cpu(), want, rtol=0, atol=1e-6)
    got = ops.skip_grad_sum(dcat.cuda(), G, C1, C2, base.clone().cuda(), skip_ref=None, accumulate=True)
    assert torch.allclose(got.cpu(), dcat.view(G, rows, C1 + C2)[:, :, C1:].sum(0) + base, rtol=0, atol=1e-6)


# ---- conv fprop ----------------------------------------------------------------------------------------------------------
FPROP = [(2, 32, 32, 64, 128, 5, 2), (3, 16, 16, 128, 256, 5, 2), (4, 8, 8, 256, 512, 5, 2), (8, 8, 8, 64, 64, 3, 1),
         (5, 32, 32, 64, 256, 3, 2), (16, 4, 4, 128, 64, 3, 2), (2, 16, 8, 64, 64, 5, 1), (6, 8, 8, 192, 320, 3, 1)]


@pytest.mark.parametrize("N,H,W,C,Co,k,st", FPROP)
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_conv2d_planes_three_products(N, H, W, C, Co, k, st, dtype):
    x, w, b = rnd((N, H, W, C), 3), rnd((k, k, C, Co), 4, (k * k * C) ** -0.5), rnd((Co,), 5, 0.1)
    assert ops.conv2d_planes_eligible(N, (H, W), C, Co, k, st)
    xp, wp = ops.split_planes(x.cuda(), dtype=dtype), ops.split_planes_transpose(w.cuda(), dtype=dtype)
    ref = O.lrelu(O.conv2d(x.double(), w.double(), b.double(), st))
    for npl in (2, 1):
        yp = torch.zeros((npl,) + tuple(ref.shape), device="cuda", dtype=dtype)
        y = ops.conv2d_planes(xp, wp, k, bias=b.cuda(), stride=st, act=True, out_planes=yp, nprod=3)
        assert int(_lib.load().ctx_last_impl()) == _lib.IMPL_P16X3
        assert rel_err(y.cpu(), ref) < TOL
        assert same_bits(yp, planes_ref(y.cpu(), dtype, npl))


@pytest.mark.parametrize("N,H,W,C,Co,k,st", FPROP)
def test_conv2d_planes_single_product(N, H, W, C, Co, k, st):
    """Also the data gradient of a transposed conv: scaled gradient planes, lrelu' mask, accumulate, amax of the result."""
    x, w = rnd((N, H, W, C), 3, 2e-5), rnd((k, k, C, Co), 4, (k * k * C) ** -0.5)
    s = scale_ref(x.abs().max().item())
    scale = torch.tensor([s], device="cuda")
    xp = ops.split_planes((x * s).cuda(), nplanes=1)
    wp = ops.split_planes_transpose(w.cuda(), nplanes=1)
    exact = O.conv2d(hi(x * s) / s, hi(w), None, st)
    y = ops.conv2d_planes(xp, wp, k, stride=st, nprod=1, a_scale=scale)
    assert int(_lib.load().ctx_last_impl()) == _lib.IMPL_P16X1
    assert rel_err(y.cpu(), exact) < TOL
    assert rel_err(y.cpu(), O.conv2d(x.double(), w.double(), None, st)) < TOL_RN11
    aref, base = rnd(tuple(y.shape), 6), rnd(tuple(y.shape), 7, float(exact.abs().max()))
    slot = torch.zeros(1, dtype=torch.int32, device="cuda")
    y2 = ops.conv2d_planes(xp, wp, k, stride=st, nprod=1, a_scale=scale, act_ref=aref.cuda(), accumulate=True,
                           out=base.clone().cuda(), amax=slot)
    want = exact * torch.where(aref >= 0, 1.0, 0.2).double() + base.double()
    assert rel_err(y2.cpu(), want) < TOL
    assert slot.cpu().view(torch.float32).item() == y2.abs().max().item()


def test_planes_reject_bad_modes():
    xp = torch.zeros((1, 2, 8, 8, 64), device="cuda", dtype=torch.bfloat16)
    wp = torch.zeros((1, 9 * 64, 64), device="cuda", dtype=torch.bfloat16)
    with pytest.raises(_lib.CtxError):                          # one product needs fp16 (bf16 keeps 8 bits)
        ops.conv2d_planes(xp, wp, 3, stride=1, nprod=1)
    with pytest.raises(_lib.CtxError):                          # three products need (hi, lo)
        ops.conv2d_planes(xp, wp, 3, stride=1, nprod=3)
    with pytest.raises(_lib.CtxError):                          # mixed element types
        ops.conv2d_planes(xp.half(), wp, 3, stride=1, nprod=1)
    assert not ops.conv2d_planes_eligible(2, (48, 48), 64, 128, 5, 2)
    assert not ops.conv2d_planes_eligible(2, (64, 64), 3, 64, 5, 2)


# ---- transposed conv ------------------------------------------------------------------------------------------------------
TRANS = [(16, 4, 4, 8, 8, 512, 512, 8, 256, 5, 2), (4, 8, 8, 16, 16, 256, 256, 2, 128, 5, 2),
         (2, 16, 16, 32, 32, 128, 128, 1, 64, 5, 2), (16, 2, 2, 4, 4, 64, 0, 0, 64, 3, 2),
         (2, 8, 8, 8, 8, 64, 64, 2, 64, 3, 1), (24, 4, 4, 8, 8, 128, 64, 8, 512, 5, 2)]


def _cat(x1, x2, N, NB):
    return x1 if x2 is None else torch.cat([x1, x2.repeat((N // NB, 1, 1, 1))], 3)


@pytest.mark.parametrize("N,P,Q,H,W,C1,C2,NB,Cn,k,st", TRANS)
def test_conv2d_transpose_planes_fp16(N, P, Q, H, W, C1, C2, NB, Cn, k, st):
    x1 = rnd((N, P, Q, C1), 6)
    x2 = rnd((NB, P, Q, C2), 7) if C2 else None
    w = rnd((k, k, Cn, C1 + C2), 8, (k * k * (C1 + C2) / st ** 2) ** -0.5)
    b = rnd((Cn,), 9, 0.1)
    assert ops.conv2d_transpose_planes_eligible(N, (H, W), C1, C2, NB, Cn, k, st)
    cat = _cat(x1, x2, N, NB)
    ref = O.lrelu(O.conv2d_transpose(cat.double(), w.double(), b.double(), (H, W), st))
    x1p, wp = ops.split_planes(x1.cuda()), ops.split_planes(w.cuda())
    x2p = None if x2 is None else ops.split_planes(x2.cuda())
    yp = torch.zeros((2,) + tuple(ref.shape), device="cuda", dtype=torch.float16)
    y = ops.conv2d_transpose_planes(x1p, wp.view(2, k * k * Cn, C1 + C2), (H, W), k, bias=b.cuda(), stride=st, x2p=x2p,
                                    act=True, out_planes=yp, nprod=3)
    assert rel_err(y.cpu(), ref) < TOL
    assert same_bits(yp, planes_ref(y.cpu(), torch.float16))
    # one product: exactly the product of the hi planes (the data gradient of a conv reads dY and the HWIO filter this way)
    exact = O.conv2d_transpose(hi(cat), hi(w), None, (H, W), st)
    aref = rnd(tuple(ref.shape), 10)
    y1 = ops.conv2d_transpose_planes(x1p[:1], wp[:1].view(1, k * k * Cn, C1 + C2), (H, W), k, stride=st,
                                     x2p=None if x2p is None else x2p[:1], nprod=1, act_ref=aref.cuda())
    assert int(_lib.load().ctx_last_impl()) == _lib.IMPL_P16X1
    assert rel_err(y1.cpu(), exact * torch.where(aref >= 0, 1.0, 0.2).double()) < TOL


@pytest.mark.parametrize("N,P,Q,C1,C2,NB,k", [(2, 16, 16, 128, 128, 1, 5), (4, 8, 8, 64, 64, 2, 3), (3, 16, 8, 64, 0, 0, 5),
                                               (16, 4, 4, 128, 64, 8, 5), (2, 32, 32, 64, 64, 2, 5)])
@pytest.mark.parametrize("Cn", [64, 128])
def test_conv2d_transpose_merged_planes_fp16(N, P, Q, C1, C2, NB, k, Cn):
    st, H, W = 2, 2 * P, 2 * Q
    x1 = rnd((N, P, Q, C1), 10)
    x2 = rnd((NB, P, Q, C2), 11) if C2 else None
    w = rnd((k, k, Cn, C1 + C2), 12, (k * k * (C1 + C2) / st ** 2) ** -0.5)
    b = rnd((Cn,), 13, 0.1)
    assert ops.conv2d_transpose_merged_eligible(N, (H, W), C1, C2, NB, Cn, k, st)
    cat = _cat(x1, x2, N, NB)
    ref = O.lrelu(O.conv2d_transpose(cat.double(), w.double(), b.double(), (H, W), st))
    wm = ops.conv2d_transpose_merged_pack(w.cuda(), (H, W), st)
    x1p = ops.split_planes(x1.cuda())
    x2p = None if x2 is None else ops.split_planes(x2.cuda())
    yp = torch.zeros((2,) + tuple(ref.shape), device="cuda", dtype=torch.float16)
    y = ops.conv2d_transpose_merged_planes(x1p, wm, (H, W), k, Cn, bias=b.cuda(), stride=st, x2p=x2p, act=True, out_planes=yp)
    assert rel_err(y.cpu(), ref) < TOL
    assert same_bits(yp, planes_ref(y.cpu(), torch.float16))
    wm1 = ops.conv2d_transpose_merged_pack(w.cuda(), (H, W), st, nplanes=1)
    y1 = ops.conv2d_transpose_merged_planes(x1p[:1], wm1, (H, W), k, Cn, stride=st, x2p=None if x2p is None else x2p[:1],
                                            nprod=1)
    assert rel_err(y1.cpu(), O.conv2d_transpose(hi(cat), hi(w), None, (H, W), st)) < TOL


# ---- filter gradient (MN-major operands, split-K) ----------------------------------------------------------------------------
# (N, H, W, Ca, Cb1, Cb2, NB, k, stride): 64-pixel K blocks inside an image and across whole images, both sources, pairs and
# single CTAs, odd tile counts (25 taps x 64 channels = 12.5 tiles), stride 1 and 2
WGRAD = [(4, 16, 16, 128, 256, 0, 0, 5, 2), (8, 8, 8, 256, 512, 0, 0, 5, 2), (4, 8, 8, 256, 512, 512, 4, 5, 2),
         (2, 32, 32, 64, 128, 128, 1, 5, 2), (32, 4, 4, 64, 64, 64, 16, 3, 2), (4, 8, 8, 128, 64, 192, 4, 3, 1),
         (16, 8, 8, 64, 128, 0, 0, 3, 1), (64, 2, 2, 128, 128, 0, 0, 3, 2)]


@pytest.mark.parametrize("N,H,W,Ca,Cb1,Cb2,NB,k,st", WGRAD)
@pytest.mark.parametrize("nprod", [1, 3])
def test_conv2d_wgrad_planes(N, H, W, Ca, Cb1, Cb2, NB, k, st, nprod):
    P, Q = -(-H // st), -(-W // st)
    assert ops.conv2d_wgrad_planes_eligible(N, (H, W), Ca, Cb1, Cb2, NB, k, st)
    big = rnd((N, H, W, Ca), 12, 3e-5)                          # the gradient operand: scaled into fp16's range
    s1 = rnd((N, P, Q, Cb1), 13)
    s2 = rnd((NB, P, Q, Cb2), 14) if Cb2 else None
    small = _cat(s1, s2, N, NB)
    s = scale_ref(big.abs().max().item())
    scale = torch.tensor([s], device="cuda")
    npl = 2 if nprod == 3 else 1
    bigp = ops.split_planes((big * s).cuda(), nplanes=npl)
    s1p = ops.split_planes(s1.cuda(), nplanes=npl)
    s2p = None if s2 is None else ops.split_planes(s2.cuda(), nplanes=npl)
    wr = torch.zeros((k, k, Ca, Cb1 + Cb2), dtype=torch.float64, requires_grad=True)
    if nprod == 1:
        (O.conv2d(hi(big * s) / s, wr, None, st) * hi(small)).sum().backward()
    else:
        (O.conv2d(big.double(), wr, None, st) * small.double()).sum().backward()
    base = rnd((k, k, Ca, Cb1 + Cb2), 15, float(wr.grad.abs().max()))
    dw = base.clone().cuda()
    ops.conv2d_wgrad_planes(bigp, s1p, k, st, small2p=s2p, dw=dw, nprod=nprod, a_scale=scale)
    assert rel_err(dw.cpu().double() - base.double(), wr.grad) < TOL


# ---- image-side layers: few-channel input, patch matrix built in shared memory (csrc/tc_few.cu) -----------------------------
# (N, H, W, C, Co, k, stride): the reference's h0 (64x64x3, 5x5 s2 -> 64), whole-image tiles, stride 1, 3x3, 128 outputs,
# odd tile counts, the 128x128 frames of the deeper net
FEW = [(3, 64, 64, 3, 64, 5, 2), (5, 16, 16, 3, 64, 5, 2), (2, 32, 32, 3, 128, 5, 2), (7, 8, 8, 3, 64, 5, 2),
       (2, 16, 16, 4, 64, 3, 1), (1, 128, 128, 3, 64, 5, 2), (4, 16, 8, 2, 128, 3, 2), (2, 16, 32, 1, 64, 5, 1),
       (100, 64, 64, 3, 64, 5, 2)]           # 800 tiles: several per persistent CTA


@pytest.mark.parametrize("N,H,W,C,Co,k,st", FEW)
def test_conv2d_few_three_products(N, H, W, C, Co, k, st):
    x, w, b = rnd((N, H, W, C), 21), rnd((k, k, C, Co), 22, (k * k * C) ** -0.5), rnd((Co,), 23, 0.1)
    assert ops.conv2d_few_eligible(N, (H, W), C, Co, k, st)
    ref = O.lrelu(O.conv2d(x.double(), w.double(), b.double(), st))
    wimg = ops.conv2d_few_pack(w.cuda())
    yp = torch.zeros((2,) + tuple(ref.shape), device="cuda", dtype=torch.float16)
    slot = torch.zeros(1, dtype=torch.int32, device="cuda")
    y = ops.conv2d_few(x.cuda(), wimg, k, Co, bias=b.cuda(), stride=st, act=True, out_planes=yp, amax=slot)
    assert tuple(y.shape) == tuple(ref.shape)
    assert rel_err(y.cpu(), ref) < TOL
    assert same_bits(yp, planes_ref(y.cpu(), torch.float16))
    assert slot.cpu().view(torch.float32).item() == y.abs().max().item()
    y2 = ops.conv2d_few(x.cuda(), wimg, k, Co, stride=st)
    assert rel_err(y2.cpu(), O.conv2d(x.double(), w.double(), None, st)) < TOL


def test_conv2d_few_uint8_frames():
    """uint8 frames are normalised inside the kernel exactly as base.py:117-119 does (three fp32 roundings)."""
    N, H, W, C, Co, k, st = 4, 64, 64, 3, 64, 5, 2
    u = torch.randint(0, 256, (N, H, W, C), generator=torch.Generator().manual_seed(24), dtype=torch.uint8)
    w, b = rnd((k, k, C, Co), 25, 0.1), rnd((Co,), 26, 0.1)
    xf = ops.u8_normalize(u.cuda())
    wimg = ops.conv2d_few_pack(w.cuda())
    ya = ops.conv2d_few(u.cuda(), wimg, k, Co, bias=b.cuda(), stride=st, act=True)
    yb = ops.conv2d_few(xf, wimg, k, Co, bias=b.cuda(), stride=st, act=True)
    assert torch.equal(ya, yb)
    assert rel_err(ya.cpu(), O.lrelu(O.conv2d(xf.cpu().double(), w.double(), b.double(), st))) < TOL


@pytest.mark.parametrize("N,H,W,C,Co,k,st", FEW)
def test_conv2d_few_single_product_gradient_input(N, H, W, C, Co, k, st):
    """The data gradient of the last decoder layer: x = d loss / d pred (tiny values, scaled inside from its max)."""
    x, w = rnd((N, H, W, C), 27, 3e-6), rnd((k, k, C, Co), 28, (k * k * C) ** -0.5)
    slot = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.amax(x.cuda(), slot)
    s = scale_ref(x.abs().max().item())
    exact = O.conv2d(hi(x * s) / s, hi(w), None, st)
    y = ops.conv2d_few(x.cuda(), ops.conv2d_few_pack(w.cuda(), nplanes=1), k, Co, stride=st, nprod=1, x_amax=slot)
    assert rel_err(y.cpu(), exact) < TOL
    assert rel_err(y.cpu(), O.conv2d(x.double(), w.double(), None, st)) < TOL_RN11


# (N, H, W, C, Cb1, Cb2, NB, k, stride)
FEW_WGRAD = [(6, 64, 64, 3, 64, 0, 0, 5, 2), (4, 64, 64, 3, 64, 64, 2, 5, 2), (5, 16, 16, 3, 128, 0, 0, 5, 2),
             (16, 8, 8, 3, 64, 64, 8, 5, 2), (2, 16, 16, 4, 64, 0, 0, 3, 1), (600, 8, 8, 3, 64, 0, 0, 5, 2),
             (120, 64, 64, 3, 64, 0, 0, 5, 2)]   # 960 tiles: the double-buffered tile loop of a CTA


@pytest.mark.parametrize("N,H,W,C,Cb1,Cb2,NB,k,st", FEW_WGRAD)
@pytest.mark.parametrize("grad_is_x", [False, True])
def test_conv2d_few_wgrad(N, H, W, C, Cb1, Cb2, NB, k, st, grad_is_x):
    """grad_is_x: the transposed conv's filter gradient (x = dY, scaled inside; small = activations);
    else the conv's (x = image; small1 = a scaled gradient plane)."""
    if Cb2 and not grad_is_x:
        pytest.skip("two small sources only occur in the transposed conv's filter gradient (x = dY)")
    P, Q = -(-H // st), -(-W // st)
    x = rnd((N, H, W, C), 29, 3e-6 if grad_is_x else 1.0)
    s1 = rnd((N, P, Q, Cb1), 30, 1.0 if grad_is_x else 2e-5)
    s2 = rnd((NB, P, Q, Cb2), 31) if Cb2 else None
    small = _cat(s1, s2, N, NB)
    slot = scale = None
    if grad_is_x:
        slot = torch.zeros(1, dtype=torch.int32, device="cuda")
        ops.amax(x.cuda(), slot)
        sx, ss = scale_ref(x.abs().max().item()), 1.0
    else:
        sx, ss = 1.0, scale_ref(s1.abs().max().item())
        scale = torch.tensor([ss], device="cuda")
    s1p = ops.split_planes((s1 * ss).cuda(), nplanes=1)
    s2p = None if s2 is None else ops.split_planes(s2.cuda(), nplanes=1)
    wr = torch.zeros((k, k, C, Cb1 + Cb2), dtype=torch.float64, requires_grad=True)
    small_h = hi(small) if s2 is None else torch.cat([hi(s1 * ss) / ss, hi(s2).repeat((N // NB, 1, 1, 1))], 3)
    if s2 is None:
        small_h = hi(s1 * ss) / ss
    (O.conv2d(hi(x * sx) / sx, wr, None, st) * small_h).sum().backward()
    base = rnd((k, k, C, Cb1 + Cb2), 32, float(wr.grad.abs().max()))
    dw = base.clone().cuda()
    ops.conv2d_few_wgrad(x.cuda(), s1p, k, st, small2p=s2p, dw=dw, x_amax=slot, r_scale=scale)
    assert rel_err(dw.cpu().double() - base.double(), wr.grad) < TOL


# (N, P, Q, C1, C2, skip batch, Cn, k, stride): the reference's d_h4 (32x32x[64+64] -> 64x64x3), partial last tile, one source,
# stride 1, 3x3, whole-image tiles, the deeper net's 64x64 -> 128x128
FEW_T = [(4, 32, 32, 64, 64, 2, 3, 5, 2), (3, 16, 16, 64, 64, 3, 3, 5, 2), (2, 32, 32, 128, 0, 0, 3, 5, 2),
         (2, 16, 16, 64, 64, 1, 4, 3, 1), (5, 8, 8, 64, 0, 0, 2, 3, 2), (1, 64, 64, 64, 64, 1, 3, 5, 2),
         (300, 32, 32, 64, 64, 150, 3, 5, 2)]


@pytest.mark.parametrize("N,P,Q,C1,C2,NB,Cn,k,st", FEW_T)
def test_conv2d_transpose_few(N, P, Q, C1, C2, NB, Cn, k, st):
    H, W = P * st, Q * st
    x1 = rnd((N, P, Q, C1), 33)
    x2 = rnd((NB, P, Q, C2), 34) if C2 else None
    w = rnd((k, k, Cn, C1 + C2), 35, (k * k * (C1 + C2) / st ** 2) ** -0.5)
    b = rnd((Cn,), 36, 0.1)
    assert ops.conv2d_transpose_few_eligible(N, (H, W), C1, C2, NB, Cn, k, st)
    small = N <= 8
    wimg = ops.conv2d_transpose_few_pack(w.cuda())
    x1p = ops.split_planes(x1.cuda())
    x2p = None if x2 is None else ops.split_planes(x2.cuda())
    y = ops.conv2d_transpose_few(x1p, wimg, (H, W), k, Cn, bias=b.cuda(), stride=st, x2p=x2p)
    if small:
        ref = O.conv2d_transpose(_cat(x1, x2, N, NB).double(), w.double(), b.double(), (H, W), st)
    else:       # many tiles per persistent CTA: against the exact-fp32 kernel (the fp64 oracle would take minutes)
        ref = ops.conv2d_transpose(x1.cuda(), w.cuda(), (H, W), bias=b.cuda(), stride=st, x2=None if x2 is None else x2.cuda(),
                                   impl=_lib.IMPL_SIMT).cpu().double()
    assert tuple(y.shape) == tuple(ref.shape)
    assert rel_err(y.cpu(), ref) < TOL
    # the reconstruction loss evaluated in the same pass == ctx_l2_loss_fwd_bwd on the stored result, bit for bit in dpred
    if N % 2 == 0:
        Bp = N // 2
        tgt, res, wts = rnd((Bp, H, W, Cn), 37).cuda(), rnd((Bp, H, W, Cn), 38, 0.1).cuda(), torch.tensor([1.0, 0.5]).cuda()
        loss, dpred = torch.zeros(2, device="cuda"), torch.empty_like(y)
        slot = torch.zeros(1, dtype=torch.int32, device="cuda")
        y2 = ops.conv2d_transpose_few(x1p, wimg, (H, W), k, Cn, bias=b.cuda(), stride=st, x2p=x2p,
                                      l2=dict(tgt=tgt, resid=res, weight=wts, loss=loss, dpred=dpred, amax=slot, pass_images=Bp))
        assert torch.equal(y2, y)
        loss_ref, dpred_ref = torch.zeros(2, device="cuda"), torch.empty_like(y)
        ops.l2_loss_fwd_bwd(y, tgt, loss_ref, dpred=dpred_ref, resid=res, weight=wts)
        assert torch.equal(dpred, dpred_ref)
        assert rel_err(loss.cpu(), loss_ref.cpu()) < 1e-5
        assert slot.cpu().view(torch.float32).item() == dpred.abs().max().item()


def test_conv2d_transpose_few_rejects():
    assert not ops.conv2d_transpose_few_eligible(2, (64, 64), 64, 64, 1, 8, 5, 2)      # kw * Cn > 16
    assert not ops.conv2d_transpose_few_eligible(2, (64, 64), 48, 0, 0, 3, 5, 2)       # channels % 64
