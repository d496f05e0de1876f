"""GPU parity of the forward path on pre-split bf16 planes (csrc/tc_bf16x3.cu): the split kernels are bit-exact against
torch's round-to-nearest bf16 cast; the convolutions are held to the same fp32 bar (2e-5, max-norm against the fp64
oracle) as the 3-pass TF32 kernels they replace in the forward pass."""
import pytest
import torch

from conftest import rel_err
from imitation_from_observation_b200 import _lib, ops
from oracle import tf_semantics as O

pytestmark = pytest.mark.gpu
TOL = 2e-5


def rnd(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def planes_ref(x):
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    return torch.stack([hi, lo])


@pytest.mark.parametrize("shape", [(4, 8, 8, 64), (3, 5, 7), (1001,)])
def test_split_bf16_bit_exact(shape):
    x = rnd(shape, 1) * torch.logspace(-6, 3, steps=shape[-1])          # wide dynamic range
    p = ops.split_bf16(x.cuda()).cpu()
    ref = planes_ref(x)
    assert p.dtype == torch.bfloat16 and tuple(p.shape) == (2,) + tuple(shape)
    assert torch.equal(p.view(torch.int16), ref.view(torch.int16))
    # hi + lo carries 16 significant bits
    assert ((p[0].float() + p[1].float() - x).abs() <= x.abs() * 2.0 ** -16).all()


def test_split_bf16_conv_filter_layout():
    w = rnd((5, 5, 64, 96), 2, 0.05)                                     # [kh,kw,Cin,Cout]
    p = ops.split_bf16_conv_filter(w.cuda()).cpu()
    ref = planes_ref(w.permute(0, 1, 3, 2).reshape(25 * 96, 64).contiguous())     # per-tap [Cout][Cin]
    assert torch.equal(p.view(torch.int16), ref.view(torch.int16))


# (N, H, W, C, Co, k, stride): exact 128-pixel tilings; single CTAs (Co = 64, 128) and CTA pairs (Co % 256 == 0), odd tile
# counts for the pair's phantom partner, whole-image groups (NB > 1), stride 1 and the stride-2 parity views
FPROP = [(2, 32, 32, 64, 128, 5, 2), (3, 16, 16, 128, 256, 5, 2), (4, 8, 8, 256, 512, 5, 2), (8, 8, 8, 64, 64, 3, 1),
         (2, 16, 16, 64, 128, 3, 1), (5, 32, 32, 64, 256, 3, 2), (16, 4, 4, 128, 64, 3, 2), (2, 16, 8, 64, 64, 5, 1)]


@pytest.mark.parametrize("N,H,W,C,Co,k,st", FPROP)
def test_conv2d_bf16x3(N, H, W, C, Co, k, st):
    x, w, b = rnd((N, H, W, C), 3), rnd((k, k, C, Co), 4, (k * k * C) ** -0.5), rnd((Co,), 5, 0.1)
    assert ops.conv2d_bf16x3_eligible(N, (H, W), C, Co, k, st)
    xp, wp = ops.split_bf16(x.cuda()), ops.split_bf16_conv_filter(w.cuda())
    ref = O.lrelu(O.conv2d(x.double(), w.double(), b.double(), st))
    yp = torch.zeros((2,) + tuple(ref.shape), device="cuda", dtype=torch.bfloat16)
    y = ops.conv2d_bf16x3(xp, wp, k, bias=b.cuda(), stride=st, act=True, out_planes=yp)
    assert int(_lib.load().ctx_last_impl()) == _lib.IMPL_BF16X3
    assert tuple(y.shape) == tuple(ref.shape)
    assert rel_err(y.cpu(), ref) < TOL
    assert torch.equal(yp.cpu().view(torch.int16), planes_ref(y.cpu()).view(torch.int16))    # planes of the stored result
    y2 = ops.conv2d_bf16x3(xp, wp, k, stride=st)                                               # no bias, no act, no planes
    assert rel_err(y2.cpu(), O.conv2d(x.double(), w.double(), None, st)) < TOL


def test_conv2d_bf16x3_rejects_other_shapes():
    assert not ops.conv2d_bf16x3_eligible(2, (64, 64), 3, 64, 5, 2)           # 3 channels
    assert not ops.conv2d_bf16x3_eligible(2, (48, 48), 64, 128, 5, 2)         # 24 x 24 does not tile into 128 pixels
    assert not ops.conv2d_bf16x3_eligible(2, (32, 32), 64, 96, 5, 2)          # Co % 64
    xp = torch.zeros((2, 2, 48, 48, 64), device="cuda", dtype=torch.bfloat16)
    wp = torch.zeros((2, 25 * 128, 64), device="cuda", dtype=torch.bfloat16)
    with pytest.raises(_lib.CtxError):
        ops.conv2d_bf16x3(xp, wp, 5, stride=2)


# (N, P, Q -> H, W, C1, C2, skip batch, Cn, k, stride); a tile of several whole images needs skip batch % images == 0
TRANS = [(16, 4, 4, 8, 8, 512, 512, 8, 256, 5, 2), (4, 8, 8, 16, 16, 256, 256, 2, 128, 5, 2),
         (2, 16, 16, 32, 32, 128, 128, 1, 64, 5, 2), (16, 2, 2, 4, 4, 64, 0, 0, 64, 3, 2),
         (2, 8, 8, 8, 8, 64, 64, 2, 64, 3, 1), (24, 4, 4, 8, 8, 128, 64, 8, 512, 5, 2)]


@pytest.mark.parametrize("N,P,Q,H,W,C1,C2,NB,Cn,k,st", TRANS)
def test_conv2d_transpose_bf16x3(N, P, Q, H, W, C1, C2, NB, Cn, k, st):
    x1 = rnd((N, P, Q, C1), 6)
    x2 = rnd((NB, P, Q, C2), 7) if C2 else None
    w = rnd((k, k, Cn, C1 + C2), 8, (k * k * (C1 + C2) / st ** 2) ** -0.5)
    b = rnd((Cn,), 9, 0.1)
    assert ops.conv2d_transpose_bf16x3_eligible(N, (H, W), C1, C2, NB, Cn, k, st)
    cat = x1 if x2 is None else torch.cat([x1, x2.repeat((N // NB, 1, 1, 1))], 3)             # image n reads skip n % NB
    ref = O.lrelu(O.conv2d_transpose(cat.double(), w.double(), b.double(), (H, W), st))
    yp = torch.zeros((2,) + tuple(ref.shape), device="cuda", dtype=torch.bfloat16)
    y = ops.conv2d_transpose_bf16x3(ops.split_bf16(x1.cuda()), ops.split_bf16(w.cuda()), (H, W), k, bias=b.cuda(), stride=st,
                                    x2p=None if x2 is None else ops.split_bf16(x2.cuda()), act=True, out_planes=yp)
    assert rel_err(y.cpu(), ref) < TOL
    assert torch.equal(yp.cpu().view(torch.int16), planes_ref(y.cpu()).view(torch.int16))


# stride-2 transposed conv with 64 output channels: the four stride phases as ONE 256-column GEMM over packed filters
@pytest.mark.parametrize("N,P,Q,C1,C2,NB,k", [(2, 16, 16, 128, 128, 1, 5), (4, 8, 8, 64, 64, 2, 3), (3, 16, 8, 64, 0, 0, 5),
                                               (16, 4, 4, 128, 64, 8, 5), (2, 32, 32, 64, 64, 2, 5)])
def test_conv2d_transpose_merged_bf16x3(N, P, Q, C1, C2, NB, k):
    Cn, st, H, W = 64, 2, 2 * P, 2 * Q
    x1 = rnd((N, P, Q, C1), 10)
    x2 = rnd((NB, P, Q, C2), 11) if C2 else None
    w = rnd((k, k, Cn, C1 + C2), 12, (k * k * (C1 + C2) / st ** 2) ** -0.5)
    b = rnd((Cn,), 13, 0.1)
    assert ops.conv2d_transpose_merged_eligible(N, (H, W), C1, C2, NB, Cn, k, st)
    cat = x1 if x2 is None else torch.cat([x1, x2.repeat((N // NB, 1, 1, 1))], 3)
    ref = O.lrelu(O.conv2d_transpose(cat.double(), w.double(), b.double(), (H, W), st))
    wm = ops.conv2d_transpose_merged_pack(w.cuda(), (H, W), st, dtype=torch.bfloat16)
    yp = torch.zeros((2,) + tuple(ref.shape), device="cuda", dtype=torch.bfloat16)
    y = ops.conv2d_transpose_merged_bf16x3(ops.split_bf16(x1.cuda()), wm, (H, W), k, Cn, bias=b.cuda(), stride=st,
                                           x2p=None if x2 is None else ops.split_bf16(x2.cuda()), act=True, out_planes=yp)
    assert rel_err(y.cpu(), ref) < TOL
    assert torch.equal(yp.cpu().view(torch.int16), planes_ref(y.cpu()).view(torch.int16))
    assert not ops.conv2d_transpose_merged_eligible(N, (H, W), C1, C2, NB, 192, k, st)      # 4 * Cn must be 256 or 512
