"""Tensor-level operator API over the C-ABI: one Python function per reference TF op call site.

Every function takes/returns CUDA ``torch.Tensor``s (fp32, contiguous NHWC) — torch is only the container and
the stream provider; all arithmetic happens in libctx_b200.so.  There is no CPU path: a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ConvGeom, Epilogue, IMPL_AUTO, IMPL_SIMT, IMPL_TC, check
from .spec import same_out, same_pad

LRELU_SLOPE = 0.2

# ---- optional per-call profiling (bench.py's roofline leg) -------------------------------------------------------
# When PROFILE is a list, every GEMM-shaped call appends (kind, impl, flops, start_event, end_event); the events
# are recorded on the stream the kernels are launched on (torch's current stream).
PROFILE = None


class _Prof:
    __slots__ = ("kind", "flops", "e0", "dims")

    def __init__(self, kind, flops, dims=None):
        self.kind, self.flops, self.dims = kind, flops, dims
        self.e0 = None
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def done(self, impl=None):
        if self.e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append((self.kind, int(_lib.load().ctx_last_impl()) if impl is None else impl, self.flops, self.e0, e1,
                            self.dims))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, name: str, dtype=torch.float32) -> int:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.CtxError("%s must be a CUDA tensor (the B200 path has no CPU fallback)" % name)
    if t.dtype != dtype:
        raise _lib.CtxError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise _lib.CtxError("%s must be contiguous" % name)
    return t.data_ptr()


def _opt(t: Optional[torch.Tensor], name: str) -> Optional[int]:
    return None if t is None else _req(t, name)


def make_epilogue(bias=None, act: bool = False, act_ref=None, accumulate: bool = False, round_tf32: bool = False,
                  slope: float = LRELU_SLOPE) -> Epilogue:
    return Epilogue(_opt(bias, "bias"), _opt(act_ref, "act_ref"), float(slope), int(act), int(accumulate),
                    int(round_tf32))


def conv_geom(n: int, big_hw: Tuple[int, int], k: int, stride: int) -> ConvGeom:
    """SAME geometry for a (transposed) conv whose big grid is ``big_hw``."""
    H, W = big_hw
    return ConvGeom(n, H, W, same_out(H, stride), same_out(W, stride), k, k, stride,
                    same_pad(H, k, stride)[0], same_pad(W, k, stride)[0])


_REPLAYED = 0        # kernels of this library run by CUDA-graph replays (the library counts a captured launch once)


def count_replayed(n: int):
    global _REPLAYED
    _REPLAYED += int(n)


def launch_count() -> int:
    """Kernels of libctx_b200.so run in this process so far: direct launches + those inside replayed CUDA graphs."""
    return int(_lib.load().ctx_launch_count()) + _REPLAYED


def device_info():
    lib = _lib.load()
    a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
    check(lib.ctx_device_info(C.byref(a), C.byref(b), C.byref(c)), "ctx_device_info")
    return a.value, b.value, c.value


def tc_available() -> bool:
    return bool(_lib.load().ctx_tc_available())


# ---- convolutions ----------------------------------------------------------------------------------------------
def conv2d(x, w, bias=None, stride: int = 2, act: bool = False, out=None, act_ref=None, accumulate=False,
           round_tf32=False, impl=IMPL_AUTO):
    """tf.nn.conv2d(SAME) + bias_add (+ lrelu) — arm_shaping.py:21-32.  x [N,H,W,C], w [kh,kw,C,Co] (HWIO)."""
    N, H, W_, Cin = x.shape
    kh, kw, Cw, Co = w.shape
    if Cw != Cin or kh != kw:
        raise _lib.CtxError("conv2d: filter %s does not match input %s" % (tuple(w.shape), tuple(x.shape)))
    g = conv_geom(N, (H, W_), kh, stride)
    if out is None:
        out = torch.empty((N, g.P, g.Q, Co), device=x.device, dtype=torch.float32)
    epi = make_epilogue(bias, act, act_ref, accumulate, round_tf32)
    pr = _Prof("fprop", 2.0 * N * g.P * g.Q * Co * kh * kw * Cin, (N, g.P, g.Q, Co, kh, kw, Cin,))
    check(_lib.load().ctx_conv2d_fprop(_req(x, "x"), Cin, _req(w, "w"), _req(out, "out"), Co, C.byref(g),
                                       C.byref(epi), impl, _stream()), "ctx_conv2d_fprop")
    pr.done()
    return out


def conv2d_transpose(x1, w, out_hw, bias=None, stride: int = 2, x2=None, act: bool = False, out=None, act_ref=None,
                     accumulate=False, round_tf32=False, impl=IMPL_AUTO):
    """tf.nn.conv2d_transpose(SAME) + bias_add (+ lrelu) over concat([x1, x2], 3) — arm_shaping.py:62-85,1323-1329.
    x1 [N,P,Q,C1]; x2 [N2,P,Q,C2] is read at image n % N2; w [kh,kw,Cn,C1+C2]."""
    N, P, Q, C1 = x1.shape
    kh, kw, Cn, Ck = w.shape
    C2 = 0 if x2 is None else x2.shape[3]
    if Ck != C1 + C2 or kh != kw:
        raise _lib.CtxError("conv2d_transpose: filter %s does not match inputs" % (tuple(w.shape),))
    g = conv_geom(N, tuple(out_hw), kh, stride)
    if (g.P, g.Q) != (P, Q):
        raise _lib.CtxError("conv2d_transpose: output %s inconsistent with input grid %s at stride %d"
                            % (tuple(out_hw), (P, Q), stride))
    if x2 is not None and tuple(x2.shape[1:3]) != (P, Q):
        raise _lib.CtxError("conv2d_transpose: skip grid mismatch")
    if out is None:
        out = torch.empty((N, g.H, g.W, Cn), device=x1.device, dtype=torch.float32)
    epi = make_epilogue(bias, act, act_ref, accumulate, round_tf32)
    pr = _Prof("transpose", 2.0 * N * P * Q * Cn * kh * kw * Ck, (N, P, Q, Cn, kh, kw, Ck,))
    check(_lib.load().ctx_conv2d_transpose(_req(x1, "x1"), C1, _opt(x2, "x2"), C2, 0 if x2 is None else x2.shape[0],
                                           _req(w, "w"), _req(out, "out"), Cn, C.byref(g), C.byref(epi), impl,
                                           _stream()), "ctx_conv2d_transpose")
    pr.done()
    return out


# ---- convolutions on 16-bit operand planes (csrc/tc_h16.cu) ---------------------------------------------------------------
# A planes tensor is torch.float16 or torch.bfloat16 of shape (nplanes,) + shape: [0] = hi = rn16(s*x), [1] = lo = rn16(s*x - hi).
PLANES_DTYPE = torch.float16          # default element type of the planes (fp16: 22 significant bits in hi + lo, and the
#                                       hi plane alone is a TF32-grade operand for the single-product backward GEMMs)


def _p16_dtype(t: torch.Tensor) -> int:
    if t.dtype == torch.float16:
        return _lib.P16_F16
    if t.dtype == torch.bfloat16:
        return _lib.P16_BF16
    raise _lib.CtxError("a planes tensor must be float16 or bfloat16, got %s" % t.dtype)


def _reqp(t: torch.Tensor, name: str, like: Optional[torch.Tensor] = None) -> int:
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype not in (torch.float16, torch.bfloat16):
        raise _lib.CtxError("%s must be a CUDA float16 / bfloat16 planes tensor" % name)
    if like is not None and t.dtype != like.dtype:
        raise _lib.CtxError("%s: every planes operand of a call must have the same element type" % name)
    if not t.is_contiguous():
        raise _lib.CtxError("%s must be contiguous" % name)
    return t.data_ptr()


def _p16(planes: torch.Tensor, nprod: int, a_scale=None, amax=None) -> "_lib.P16":
    if nprod == 3 and planes.shape[0] < 2:
        raise _lib.CtxError("three-product mode needs (hi, lo) planes")
    return _lib.P16(_p16_dtype(planes), int(nprod), None if a_scale is None else _req(a_scale, "a_scale"),
                    None if amax is None else _req(amax, "amax", torch.int32))


def split_planes(x, out=None, nplanes: int = 2, dtype=None, scale: float = 1.0):
    """fp32 tensor -> planes (nplanes, *x.shape) of scale * x (scale: a power of two)."""
    p = _req(x, "x")
    if out is None:
        out = torch.empty((nplanes,) + tuple(x.shape), device=x.device, dtype=dtype or PLANES_DTYPE)
    elif out.numel() != nplanes * x.numel():
        raise _lib.CtxError("split_planes: out has the wrong size")
    check(_lib.load().ctx_split_planes(p, x.numel(), float(scale), _reqp(out, "out"), nplanes, _p16_dtype(out), _stream()),
          "ctx_split_planes")
    return out


def split_planes_transpose(w, out=None, nplanes: int = 2, dtype=None, scale: float = 1.0):
    """w [kh,kw,R,Cc] -> planes (nplanes, kh*kw*Cc, R): a conv filter [kh,kw,Cin,Cout] becomes the per-tap K-major rows
    [Cout][Cin] conv2d_planes reads; a transposed-conv filter [kh,kw,Cout,Cin] becomes the rows [Cin][Cout] its data
    gradient reads."""
    kh, kw, R, Cc = w.shape
    if out is None:
        out = torch.empty((nplanes, kh * kw * Cc, R), device=w.device, dtype=dtype or PLANES_DTYPE)
    elif out.numel() != nplanes * w.numel():
        raise _lib.CtxError("split_planes_transpose: out has the wrong size")
    check(_lib.load().ctx_split_planes_transpose(_req(w, "w"), kh * kw, R, Cc, float(scale), _reqp(out, "out"), nplanes,
                                                 _p16_dtype(out), _stream()), "ctx_split_planes_transpose")
    return out


def split_bf16(x, out=None):
    return split_planes(x, out=out, nplanes=2, dtype=torch.bfloat16)


def split_bf16_conv_filter(w, out=None):
    return split_planes_transpose(w, out=out, nplanes=2, dtype=torch.bfloat16)


def conv2d_planes_eligible(n, in_hw, Cin, Co, k, stride) -> bool:
    g = conv_geom(n, tuple(in_hw), k, stride)
    return bool(_lib.load().ctx_conv2d_fprop_p16_eligible(Cin, Co, C.byref(g)))


def conv2d_planes(xp, wp, k: int, bias=None, stride: int = 2, act: bool = False, out=None, out_planes=None, nprod: int = 3,
                  act_ref=None, accumulate: bool = False, a_scale=None, amax=None):
    """conv2d(SAME) + bias (+ lrelu) from planes: xp (np,N,H,W,C), wp (np,k*k*Co,C) -> out fp32 [N,P,Q,Co] (+ out_planes).
    nprod = 1 with a_scale / amax is the data gradient of a transposed conv (xp = planes of dY)."""
    _, N, H, W_, Cin = xp.shape
    Co = wp.shape[1] // (k * k)
    if wp.shape[2] != Cin or wp.shape[1] != k * k * Co:
        raise _lib.CtxError("conv2d_planes: filter planes %s do not match input planes %s" % (tuple(wp.shape), tuple(xp.shape)))
    g = conv_geom(N, (H, W_), k, stride)
    if out is None:
        out = torch.empty((N, g.P, g.Q, Co), device=xp.device, dtype=torch.float32)
    if out_planes is not None and out_planes.numel() != out_planes.shape[0] * out.numel():
        raise _lib.CtxError("conv2d_planes: out_planes has the wrong size")
    epi = make_epilogue(bias, act, act_ref, accumulate)
    cfg = _p16(xp, nprod, a_scale, amax)
    pr = _Prof("fprop", 2.0 * N * g.P * g.Q * Co * k * k * Cin, (N, g.P, g.Q, Co, k, k, Cin,))
    check(_lib.load().ctx_conv2d_fprop_p16(_reqp(xp, "xp"), Cin, _reqp(wp, "wp", xp), _req(out, "out"),
                                           None if out_planes is None else _reqp(out_planes, "out_planes", xp),
                                           0 if out_planes is None else out_planes.shape[0], Co, C.byref(g), C.byref(epi),
                                           C.byref(cfg), _stream()), "ctx_conv2d_fprop_p16")
    pr.done()
    return out


def conv2d_transpose_planes_eligible(n, out_hw, C1, C2, x2_batch, Cn, k, stride) -> bool:
    g = conv_geom(n, tuple(out_hw), k, stride)
    return bool(_lib.load().ctx_conv2d_transpose_p16_eligible(C1, C2, x2_batch, Cn, C.byref(g)))


def conv2d_transpose_planes(x1p, wp, out_hw, k: int, bias=None, stride: int = 2, x2p=None, act: bool = False, out=None,
                            out_planes=None, nprod: int = 3, act_ref=None, accumulate: bool = False, a_scale=None,
                            amax=None):
    """conv2d_transpose(SAME) + bias (+ lrelu) over concat([x1, x2], 3) from planes: x1p (np,N,P,Q,C1), x2p (np,N2,P,Q,C2),
    wp (np, k*k*Cn, C1+C2) = planes of the filter [k,k,Cn,Ck] -> out fp32 [N,H,W,Cn] (+ out_planes).  nprod = 1 with
    a_scale / amax / act_ref / accumulate is the data gradient of a conv (x1p = planes of dY, wp = planes of the HWIO filter)."""
    _, N, P, Q, C1 = x1p.shape
    C2 = 0 if x2p is None else x2p.shape[4]
    Ck = wp.shape[-1]
    wp = wp.view(wp.shape[0], -1, Ck)                     # (np, k, k, Cn, Ck) planes of the filter as stored are fine too
    Cn = wp.shape[1] // (k * k)
    if Ck != C1 + C2 or wp.shape[1] != k * k * Cn:
        raise _lib.CtxError("conv2d_transpose_planes: filter planes do not match the inputs")
    g = conv_geom(N, tuple(out_hw), k, stride)
    if (g.P, g.Q) != (P, Q):
        raise _lib.CtxError("conv2d_transpose_planes: output %s inconsistent with input grid %s" % (tuple(out_hw), (P, Q)))
    if out is None:
        out = torch.empty((N, g.H, g.W, Cn), device=x1p.device, dtype=torch.float32)
    if out_planes is not None and out_planes.numel() != out_planes.shape[0] * out.numel():
        raise _lib.CtxError("conv2d_transpose_planes: out_planes has the wrong size")
    epi = make_epilogue(bias, act, act_ref, accumulate)
    cfg = _p16(x1p, nprod, a_scale, amax)
    pr = _Prof("transpose", 2.0 * N * P * Q * Cn * k * k * Ck, (N, P, Q, Cn, k, k, Ck,))
    check(_lib.load().ctx_conv2d_transpose_p16(
        _reqp(x1p, "x1p"), C1, None if x2p is None else _reqp(x2p, "x2p", x1p), C2, 0 if x2p is None else x2p.shape[1],
        _reqp(wp, "wp", x1p), _req(out, "out"), None if out_planes is None else _reqp(out_planes, "out_planes", x1p),
        0 if out_planes is None else out_planes.shape[0], Cn, C.byref(g), C.byref(epi), C.byref(cfg), _stream()),
        "ctx_conv2d_transpose_p16")
    pr.done()
    return out


def conv2d_transpose_merged_eligible(n, out_hw, C1, C2, x2_batch, Cn, k, stride) -> bool:
    g = conv_geom(n, tuple(out_hw), k, stride)
    return bool(_lib.load().ctx_conv2d_transpose_merged_p16_eligible(C1, C2, x2_batch, Cn, C.byref(g)))


def conv2d_transpose_merged_pack(w, out_hw, stride: int = 2, out=None, nplanes: int = 2, dtype=None, scale: float = 1.0):
    """Transposed-conv filter [k,k,Cn,Ck] -> packed planes (nplanes, rows, Ck) for conv2d_transpose_merged_planes."""
    k, _, Cn, Ck = w.shape
    g = conv_geom(1, tuple(out_hw), k, stride)
    rows = int(_lib.load().ctx_conv2d_transpose_merged_rows(Cn, C.byref(g)))
    if rows <= 0:
        raise _lib.CtxError("conv2d_transpose_merged_pack: geometry not eligible")
    if out is None:
        out = torch.empty((nplanes, rows, Ck), device=w.device, dtype=dtype or PLANES_DTYPE)
    elif tuple(out.shape) != (nplanes, rows, Ck):
        raise _lib.CtxError("conv2d_transpose_merged_pack: out must be (%d, %d, %d)" % (nplanes, rows, Ck))
    check(_lib.load().ctx_conv2d_transpose_merged_pack_p16(_req(w, "w"), Ck, Cn, C.byref(g), float(scale), _reqp(out, "out"),
                                                           nplanes, _p16_dtype(out), _stream()),
          "ctx_conv2d_transpose_merged_pack_p16")
    return out


def conv2d_transpose_merged_planes(x1p, wmp, out_hw, k: int, Cn: int, bias=None, stride: int = 2, x2p=None,
                                   act: bool = False, out=None, out_planes=None, nprod: int = 3, act_ref=None,
                                   accumulate: bool = False, a_scale=None, amax=None):
    """conv2d_transpose(SAME, stride 2) + bias (+ lrelu) with the four stride phases merged into one GEMM; wmp from
    conv2d_transpose_merged_pack."""
    _, N, P, Q, C1 = x1p.shape
    C2 = 0 if x2p is None else x2p.shape[4]
    if wmp.shape[-1] != C1 + C2:
        raise _lib.CtxError("conv2d_transpose_merged_planes: packed filter does not match the inputs")
    g = conv_geom(N, tuple(out_hw), k, stride)
    if (g.P, g.Q) != (P, Q):
        raise _lib.CtxError("conv2d_transpose_merged_planes: output %s inconsistent with input grid %s" % (tuple(out_hw), (P, Q)))
    if out is None:
        out = torch.empty((N, g.H, g.W, Cn), device=x1p.device, dtype=torch.float32)
    if out_planes is not None and out_planes.numel() != out_planes.shape[0] * out.numel():
        raise _lib.CtxError("conv2d_transpose_merged_planes: out_planes has the wrong size")
    epi = make_epilogue(bias, act, act_ref, accumulate)
    cfg = _p16(x1p, nprod, a_scale, amax)
    pr = _Prof("transpose", 2.0 * N * P * Q * Cn * k * k * (C1 + C2), (N, P, Q, Cn, k, k, (C1 + C2),))
    check(_lib.load().ctx_conv2d_transpose_merged_p16(
        _reqp(x1p, "x1p"), C1, None if x2p is None else _reqp(x2p, "x2p", x1p), C2, 0 if x2p is None else x2p.shape[1],
        _reqp(wmp, "wmp", x1p), _req(out, "out"), None if out_planes is None else _reqp(out_planes, "out_planes", x1p),
        0 if out_planes is None else out_planes.shape[0], Cn, C.byref(g), C.byref(epi), C.byref(cfg), _stream()),
        "ctx_conv2d_transpose_merged_p16")
    pr.done()
    return out


def conv2d_wgrad_planes_eligible(n, big_hw, Ca, Cb1, Cb2, small2_batch, k, stride) -> bool:
    g = conv_geom(n, tuple(big_hw), k, stride)
    return bool(_lib.load().ctx_conv2d_wgrad_p16_eligible(Ca, Cb1, Cb2, small2_batch, C.byref(g)))


def conv2d_wgrad_planes(bigp, small1p, k: int, stride: int, small2p=None, dw=None, nprod: int = 1, a_scale=None):
    """dw[k,k,Ca,Cb1+Cb2] += sum_pixels big[shifted] (x) [small1 ‖ small2] from planes: bigp (np,N,H,W,Ca), small*p
    (np,.,P,Q,Cb*).  a_scale: device scalar the gradient operand's planes were scaled by."""
    _, N, H, W_, Ca = bigp.shape
    Cb1 = small1p.shape[4]
    Cb2 = 0 if small2p is None else small2p.shape[4]
    g = conv_geom(N, (H, W_), k, stride)
    if tuple(small1p.shape[1:4]) != (N, g.P, g.Q):
        raise _lib.CtxError("conv2d_wgrad_planes: small grid %s does not match geometry (%d,%d,%d)"
                            % (tuple(small1p.shape), N, g.P, g.Q))
    if dw is None:
        dw = torch.zeros((k, k, Ca, Cb1 + Cb2), device=bigp.device, dtype=torch.float32)
    cfg = _p16(bigp, nprod, a_scale, None)
    if nprod == 3 and (small1p.shape[0] < 2 or (small2p is not None and small2p.shape[0] < 2)):
        raise _lib.CtxError("conv2d_wgrad_planes: three-product mode needs (hi, lo) planes of every operand")
    pr = _Prof("wgrad", 2.0 * N * g.P * g.Q * (Cb1 + Cb2) * k * k * Ca, (N, g.P, g.Q, (Cb1 + Cb2), k, k, Ca,))
    check(_lib.load().ctx_conv2d_wgrad_p16(_reqp(bigp, "bigp"), Ca, _reqp(small1p, "small1p", bigp), Cb1,
                                           None if small2p is None else _reqp(small2p, "small2p", bigp), Cb2,
                                           0 if small2p is None else small2p.shape[1], _req(dw, "dw"), C.byref(g),
                                           C.byref(cfg), _stream()), "ctx_conv2d_wgrad_p16")
    pr.done()
    return dw


def grad_planes(x, C_: int, planes=None, ref=None, amax=None, scale=None, db=None, out_f32=None, ldx: Optional[int] = None,
                slope: float = LRELU_SLOPE):
    """Raw fp32 gradient x (any shape, rows of `ldx` floats of which the first C_ are used) -> lrelu' mask from `ref`
    -> planes scaled by the power of two written to `scale` (from `amax`, the bits of max |x|), + bias gradient db, +
    optional fp32 copy.  planes: (np, rows, C_)-sized tensor."""
    ldx = C_ if ldx is None else ldx
    rows = x.numel() // ldx
    if not x.is_contiguous() or x.numel() != rows * ldx:
        raise _lib.CtxError("grad_planes: x must be dense rows of ldx floats")
    if planes is not None and planes.numel() != planes.shape[0] * rows * C_:
        raise _lib.CtxError("grad_planes: planes has the wrong size")
    check(_lib.load().ctx_grad_planes(_req(x, "x"), rows, C_, ldx, _opt(ref, "ref"), float(slope),
                                      None if amax is None else _req(amax, "amax", torch.int32), _opt(scale, "scale"),
                                      None if planes is None else _reqp(planes, "planes"),
                                      0 if planes is None else planes.shape[0],
                                      _lib.P16_F16 if planes is None else _p16_dtype(planes), _opt(db, "db"),
                                      _opt(out_f32, "out_f32"), _stream()), "ctx_grad_planes")
    return planes


def amax(x, slot):
    """slot (int32[1], device) = max(slot, bits(max |x|))."""
    check(_lib.load().ctx_amax(_req(x, "x"), x.numel(), _req(slot, "slot", torch.int32), _stream()), "ctx_amax")
    return slot


def skip_grad_sum(dcat, G: int, C1: int, C2: int, dskip, skip_ref=None, accumulate: bool = False, slope: float = LRELU_SLOPE):
    """dskip[rows, C2] (+)= (sum over the G stacked passes of dcat[..., C1:]) * lrelu'(skip_ref)."""
    rows = dcat.numel() // (G * (C1 + C2))
    check(_lib.load().ctx_skip_grad_sum(_req(dcat, "dcat"), G, rows, C1, C2, _opt(skip_ref, "skip_ref"), float(slope),
                                        _req(dskip, "dskip"), int(accumulate), _stream()), "ctx_skip_grad_sum")
    return dskip


# ---- image-side layers: few-channel input, patch matrix built in shared memory (csrc/tc_few.cu) -----------------------------
def conv2d_few_eligible(n, in_hw, Cin, Co, k, stride) -> bool:
    g = conv_geom(n, tuple(in_hw), k, stride)
    return bool(_lib.load().ctx_conv2d_few_eligible(Cin, Co, C.byref(g)))


def conv2d_few_pack(w, nplanes: int = 2, out=None):
    """HWIO filter [k,k,C,Co] -> the packed operand image conv2d_few reads (uint8 bytes; once per parameter update)."""
    k, _, Cin, Co = w.shape
    g = conv_geom(1, (k, k), k, 1)
    nbytes = int(_lib.load().ctx_conv2d_few_pack_bytes(Co, C.byref(g), nplanes))
    if out is None:
        out = torch.empty(nbytes, device=w.device, dtype=torch.uint8)
    elif out.numel() != nbytes or out.dtype != torch.uint8:
        raise _lib.CtxError("conv2d_few_pack: out must be %d uint8 bytes" % nbytes)
    check(_lib.load().ctx_conv2d_few_pack(_req(w, "w"), Cin, Co, C.byref(g), _req(out, "out", torch.uint8), nplanes, _stream()),
          "ctx_conv2d_few_pack")
    return out


def _req_img(x, name):
    if not isinstance(x, torch.Tensor) or not x.is_cuda or x.dtype not in (torch.float32, torch.uint8) or not x.is_contiguous():
        raise _lib.CtxError("%s must be a contiguous CUDA float32 or uint8 tensor" % name)
    return x.data_ptr()


def conv2d_few(x, wimg, k: int, Co: int, bias=None, stride: int = 2, act: bool = False, out=None, out_planes=None,
               nprod: int = 3, x_amax=None, amax=None):
    """conv2d(SAME) + bias (+ lrelu) of a few-channel x [N,H,W,C] (fp32, or uint8 frames normalised on the fly) with the
    packed filter `wimg` (conv2d_few_pack).  x_amax: x is a gradient, scaled internally by the power of two its max asks for."""
    N, H, W_, Cin = x.shape
    g = conv_geom(N, (H, W_), k, stride)
    if out is None:
        out = torch.empty((N, g.P, g.Q, Co), device=x.device, dtype=torch.float32)
    if out_planes is not None and (out_planes.dtype != torch.float16 or out_planes.numel() != out_planes.shape[0] * out.numel()):
        raise _lib.CtxError("conv2d_few: out_planes must be float16 planes of the output")
    epi = make_epilogue(bias, act)
    pr = _Prof("fprop", 2.0 * N * g.P * g.Q * Co * k * k * Cin, (N, g.P, g.Q, Co, k, k, Cin,))
    check(_lib.load().ctx_conv2d_few_fprop(
        _req_img(x, "x"), int(x.dtype == torch.uint8), Cin, _req(wimg, "wimg", torch.uint8), _req(out, "out"),
        None if out_planes is None else _reqp(out_planes, "out_planes"), 0 if out_planes is None else out_planes.shape[0],
        Co, C.byref(g), C.byref(epi), int(nprod), None if x_amax is None else _req(x_amax, "x_amax", torch.int32),
        None if amax is None else _req(amax, "amax", torch.int32), _stream()), "ctx_conv2d_few_fprop")
    pr.done(_lib.IMPL_P16X3 if nprod == 3 else _lib.IMPL_P16X1)
    return out


def conv2d_few_wgrad(x, small1p, k: int, stride: int, small2p=None, dw=None, x_amax=None, r_scale=None):
    """dw[k,k,C,Cb1+Cb2] += patches(x)^T [small1 ‖ small2]; small*p: one fp16 plane (1,.,P,Q,Cb*) each."""
    N, H, W_, Cin = x.shape
    Cb1 = small1p.shape[-1]
    Cb2 = 0 if small2p is None else small2p.shape[-1]
    g = conv_geom(N, (H, W_), k, stride)
    if small1p.dtype != torch.float16 or small1p.numel() != N * g.P * g.Q * Cb1:
        raise _lib.CtxError("conv2d_few_wgrad: small1p must be one float16 plane of [N,P,Q,Cb1]")
    if small2p is not None and (small2p.dtype != torch.float16 or small2p.numel() % (g.P * g.Q * Cb2) != 0):
        raise _lib.CtxError("conv2d_few_wgrad: small2p must be one float16 plane of [N2,P,Q,Cb2]")
    if dw is None:
        dw = torch.zeros((k, k, Cin, Cb1 + Cb2), device=x.device, dtype=torch.float32)
    pr = _Prof("wgrad", 2.0 * N * g.P * g.Q * (Cb1 + Cb2) * k * k * Cin, (N, g.P, g.Q, (Cb1 + Cb2), k, k, Cin,))
    check(_lib.load().ctx_conv2d_few_wgrad(
        _req_img(x, "x"), int(x.dtype == torch.uint8), Cin, _reqp(small1p, "small1p"), Cb1,
        None if small2p is None else _reqp(small2p, "small2p"), Cb2,
        0 if small2p is None else small2p.numel() // (g.P * g.Q * Cb2), _req(dw, "dw"), C.byref(g),
        None if x_amax is None else _req(x_amax, "x_amax", torch.int32), _opt(r_scale, "r_scale"), _stream()),
        "ctx_conv2d_few_wgrad")
    pr.done(_lib.IMPL_P16X1)
    return dw


def conv2d_transpose_few_eligible(n, out_hw, C1, C2, x2_batch, Cn, k, stride) -> bool:
    g = conv_geom(n, tuple(out_hw), k, stride)
    return bool(_lib.load().ctx_conv2d_transpose_few_eligible(C1, C2, x2_batch, Cn, C.byref(g)))


def conv2d_transpose_few_pack(w, out=None):
    """Transposed-conv filter [k,k,Cn,Ck] (few output channels Cn) -> the packed operand image conv2d_transpose_few reads."""
    k, _, Cn, Ck = w.shape
    g = conv_geom(1, (k, k), k, 1)
    nbytes = int(_lib.load().ctx_conv2d_transpose_few_pack_bytes(Ck, C.byref(g)))
    if nbytes <= 0:
        raise _lib.CtxError("conv2d_transpose_few_pack: filter %s not eligible" % (tuple(w.shape),))
    if out is None:
        out = torch.empty(nbytes, device=w.device, dtype=torch.uint8)
    elif out.numel() != nbytes or out.dtype != torch.uint8:
        raise _lib.CtxError("conv2d_transpose_few_pack: out must be %d uint8 bytes" % nbytes)
    check(_lib.load().ctx_conv2d_transpose_few_pack(_req(w, "w"), Ck, Cn, C.byref(g), _req(out, "out", torch.uint8), _stream()),
          "ctx_conv2d_transpose_few_pack")
    return out


def conv2d_transpose_few(x1p, wimg, out_hw, k: int, Cn: int, bias=None, stride: int = 2, x2p=None, act: bool = False, out=None,
                         l2: Optional[dict] = None):
    """conv2d_transpose(SAME) + bias (+ lrelu) to a FEW output channels over concat([x1, x2], 3) from fp16 (hi, lo) planes
    x1p (2,N,P,Q,C1), x2p (2,N2,P,Q,C2); wimg from conv2d_transpose_few_pack.  l2 = dict(tgt, loss, pass_images[, resid, weight,
    dpred, amax]) also evaluates the reconstruction loss of the result in the same pass (ops.l2_loss_fwd_bwd semantics)."""
    _, N, P, Q, C1 = x1p.shape
    C2 = 0 if x2p is None else x2p.shape[4]
    if x1p.dtype != torch.float16 or x1p.shape[0] != 2 or (x2p is not None and (x2p.dtype != torch.float16 or x2p.shape[0] != 2)):
        raise _lib.CtxError("conv2d_transpose_few: operands must be float16 (hi, lo) planes")
    g = conv_geom(N, tuple(out_hw), k, stride)
    if (g.P, g.Q) != (P, Q):
        raise _lib.CtxError("conv2d_transpose_few: output %s inconsistent with input grid %s" % (tuple(out_hw), (P, Q)))
    if out is None:
        out = torch.empty((N, g.H, g.W, Cn), device=x1p.device, dtype=torch.float32)
    epi = make_epilogue(bias, act)
    lf = None
    if l2 is not None:
        lf = _lib.L2Fused(_req(l2["tgt"], "tgt"), _opt(l2.get("resid"), "resid"), _opt(l2.get("weight"), "weight"),
                          _req(l2["loss"], "loss"), _opt(l2.get("dpred"), "dpred"),
                          None if l2.get("amax") is None else _req(l2["amax"], "amax", torch.int32), int(l2["pass_images"]))
    pr = _Prof("transpose", 2.0 * N * P * Q * Cn * k * k * (C1 + C2), (N, P, Q, Cn, k, k, (C1 + C2),))
    check(_lib.load().ctx_conv2d_transpose_few(
        _reqp(x1p, "x1p"), C1, None if x2p is None else _reqp(x2p, "x2p", x1p), C2, 0 if x2p is None else x2p.shape[1],
        _req(wimg, "wimg", torch.uint8), _req(out, "out"), Cn, C.byref(g), C.byref(epi),
        None if lf is None else C.byref(lf), _stream()), "ctx_conv2d_transpose_few")
    pr.done(_lib.IMPL_P16X3)
    return out


# round-1 names (bf16 planes, three products)
conv2d_bf16x3_eligible = conv2d_planes_eligible
conv2d_bf16x3 = conv2d_planes
conv2d_transpose_bf16x3_eligible = conv2d_transpose_planes_eligible
conv2d_transpose_bf16x3 = conv2d_transpose_planes
conv2d_transpose_merged_bf16x3 = conv2d_transpose_merged_planes


def conv2d_wgrad(big, small1, k: int, stride: int, small2=None, dw=None, impl=IMPL_AUTO):
    """dw[k,k,Ca,Cb1+Cb2] += sum_pixels big[shifted] (x) [small1 ‖ small2].  big [N,H,W,Ca], small* [.,P,Q,Cb*]."""
    N, H, W_, Ca = big.shape
    Cb1 = small1.shape[3]
    Cb2 = 0 if small2 is None else small2.shape[3]
    g = conv_geom(N, (H, W_), k, stride)
    if tuple(small1.shape[:3]) != (N, g.P, g.Q):
        raise _lib.CtxError("conv2d_wgrad: small grid %s does not match geometry (%d,%d,%d)"
                            % (tuple(small1.shape), N, g.P, g.Q))
    if dw is None:
        dw = torch.zeros((k, k, Ca, Cb1 + Cb2), device=big.device, dtype=torch.float32)
    pr = _Prof("wgrad", 2.0 * N * g.P * g.Q * (Cb1 + Cb2) * k * k * Ca, (N, g.P, g.Q, (Cb1 + Cb2), k, k, Ca,))
    check(_lib.load().ctx_conv2d_wgrad(_req(big, "big"), Ca, _req(small1, "small1"), Cb1, _opt(small2, "small2"), Cb2,
                                       0 if small2 is None else small2.shape[0], _req(dw, "dw"), C.byref(g), impl,
                                       _stream()), "ctx_conv2d_wgrad")
    pr.done()
    return dw


def _geom_general(n, H, W_, kh, kw, stride, padding):
    """TF geometry for a rectangular window: padding 'SAME' (asymmetric, low half first) or 'VALID'."""
    if padding == "SAME":
        P, Q = same_out(H, stride), same_out(W_, stride)
        ph, pw = same_pad(H, kh, stride)[0], same_pad(W_, kw, stride)[0]
    elif padding == "VALID":
        P, Q = (H - kh) // stride + 1, (W_ - kw) // stride + 1
        ph = pw = 0
    else:
        raise _lib.CtxError("padding must be 'SAME' or 'VALID'")
    if P <= 0 or Q <= 0:
        raise _lib.CtxError("window %dx%d does not fit a %dx%d input" % (kh, kw, H, W_))
    return ConvGeom(n, H, W_, P, Q, kh, kw, stride, ph, pw)


def conv2d_general(x, w, bias=None, stride: int = 1, padding: str = "SAME", act: bool = False, slope: float = 0.0,
                   out=None, impl=IMPL_AUTO, into=None, c_off: int = 0):
    """slim.conv2d as InceptionV3 uses it (nets/inception_v3.py:104-413): rectangular kernels (1x7, 7x1, 1x3, 3x1),
    SAME or VALID, folded-BN bias, activation max(v, slope*v) (slope 0 = ReLU).  x [N,H,W,C], w [kh,kw,C,Co]."""
    N, H, W_, Cin = x.shape
    kh, kw, Cw, Co = w.shape
    if Cw != Cin:
        raise _lib.CtxError("conv2d_general: filter %s does not match input %s" % (tuple(w.shape), tuple(x.shape)))
    g = _geom_general(N, H, W_, kh, kw, stride, padding)
    epi = make_epilogue(bias, act, None, False, False, slope)
    pr = _Prof("fprop", 2.0 * N * g.P * g.Q * Co * kh * kw * Cin, (N, g.P, g.Q, Co, kh, kw, Cin,))
    if into is not None:
        # write channels [c_off, c_off + Co) of the wider NHWC tensor `into` (the block's concat result)
        if tuple(into.shape[:3]) != (N, g.P, g.Q) or c_off < 0 or c_off + Co > into.shape[3]:
            raise _lib.CtxError("conv2d_general: destination %s does not take [%d,%d,%d,%d] at channel %d"
                                % (tuple(into.shape), N, g.P, g.Q, Co, c_off))
        check(_lib.load().ctx_conv2d_fprop_ld(_req(x, "x"), Cin, _req(w, "w"), _req(into, "into") + 4 * c_off,
                                              into.shape[3], Co, C.byref(g), C.byref(epi), impl, _stream()),
              "ctx_conv2d_fprop_ld")
        pr.done()
        return into
    if out is None:
        out = torch.empty((N, g.P, g.Q, Co), device=x.device, dtype=torch.float32)
    check(_lib.load().ctx_conv2d_fprop(_req(x, "x"), Cin, _req(w, "w"), _req(out, "out"), Co, C.byref(g),
                                       C.byref(epi), impl, _stream()), "ctx_conv2d_fprop")
    pr.done()
    return out


def pool2d(x, k: int, stride: int, padding: str, mode: str, out=None):
    """slim.max_pool2d / avg_pool2d on NHWC (avg: TF's count-valid divisor)."""
    N, H, W_, Cc = x.shape
    g = _geom_general(N, H, W_, k, k, stride, padding)
    if out is None:
        out = torch.empty((N, g.P, g.Q, Cc), device=x.device, dtype=torch.float32)
    check(_lib.load().ctx_pool2d(_req(x, "x"), _req(out, "out"), Cc, C.byref(g), {"max": 0, "avg": 1}[mode], _stream()),
          "ctx_pool2d")
    return out


def copy_channels(src, dst, c_off: int):
    """dst[..., c_off:c_off+Cs] = src  (tf.concat(axis=3) into a pre-allocated NHWC tensor)."""
    Cs, Cd = src.shape[-1], dst.shape[-1]
    rows = src.numel() // Cs
    check(_lib.load().ctx_copy_channels(_req(src, "src"), Cs, _req(dst, "dst"), Cd, c_off, rows, _stream()),
          "ctx_copy_channels")
    return dst


# ---- dense -------------------------------------------------------------------------------------------------------
def gemm_raw(tA: bool, tB: bool, M: int, N: int, K: int, A_ptr: int, lda: int, B_ptr: int, ldb: int, C_ptr: int,
             ldc: int, epi: Optional[Epilogue] = None, impl=IMPL_AUTO):
    pr = _Prof("gemm", 2.0 * M * N * K, (M, N, K,))
    check(_lib.load().ctx_gemm(int(tA), int(tB), M, N, K, A_ptr, lda, B_ptr, ldb, C_ptr, ldc,
                               C.byref(epi) if epi is not None else None, impl, _stream()), "ctx_gemm")
    pr.done()


def _mat(t: torch.Tensor, name: str):
    """(ptr, rows, cols, ld) of a 2-D row-major view with unit column stride (column slices allowed)."""
    if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise _lib.CtxError("%s must be a 2-D fp32 CUDA matrix with unit column stride" % name)
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))
    return t.data_ptr(), t.shape[0], t.shape[1], ld


def gemm(A, B, out=None, transA=False, transB=False, bias=None, act=False, act_ref=None, accumulate=False,
         round_tf32=False, impl=IMPL_AUTO, slope: float = LRELU_SLOPE):
    """C = opA(A) @ opB(B) with the fused epilogue (tf.matmul + bias, arm_shaping.py:48-59, and its gradients).
    A, B, out may be row-slices / column-slices of larger row-major matrices."""
    pa, ar, ac, lda = _mat(A, "A")
    pb, br, bc, ldb = _mat(B, "B")
    M, K = (ac, ar) if transA else (ar, ac)
    K2, N = (bc, br) if transB else (br, bc)
    if K != K2:
        raise _lib.CtxError("gemm: inner dimensions differ (%d vs %d)" % (K, K2))
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=torch.float32)
    pc, cr, cc, ldc = _mat(out, "out")
    if (cr, cc) != (M, N):
        raise _lib.CtxError("gemm: out has shape %s, expected %s" % ((cr, cc), (M, N)))
    if act_ref is not None and (not act_ref.is_contiguous() or ldc != N):
        raise _lib.CtxError("gemm: act_ref requires a dense output")
    epi = make_epilogue(bias, act, act_ref, accumulate, round_tf32, slope)
    gemm_raw(transA, transB, M, N, K, pa, lda, pb, ldb, pc, ldc, epi, impl)
    return out


# ---- pointwise / reductions -----------------------------------------------------------------------------------------
def u8_normalize(x_u8, out=None):
    """convert_image_dtype, -0.5, *2 — base.py:117-119."""
    p = _req(x_u8, "x", torch.uint8)
    if out is None:
        out = torch.empty(x_u8.shape, device=x_u8.device, dtype=torch.float32)
    check(_lib.load().ctx_u8_normalize(p, _req(out, "out"), x_u8.numel(), _stream()), "ctx_u8_normalize")
    return out


def gather_frames(bank, frame_idx, out=None, lut=None):
    """out[b] = float(bank[frame_idx[b]]) — the trainer's batch assembly (train_script.py:154-159) on a device bank.
    bank [frames, ...] fp32 or uint8 (uint8 needs `lut`, 256 floats on the device); frame_idx int64 on the device."""
    if bank.dtype not in (torch.float32, torch.uint8):
        raise _lib.CtxError("gather_frames: the bank must be float32 or uint8")
    is_u8 = bank.dtype == torch.uint8
    pb = _req(bank, "bank", bank.dtype)
    pi = _req(frame_idx, "frame_idx", torch.int64)
    n = frame_idx.numel()
    frame_shape = tuple(bank.shape[1:])
    elems = 1
    for d in frame_shape:
        elems *= int(d)
    if out is None:
        out = torch.empty((n,) + frame_shape, device=bank.device, dtype=torch.float32)
    elif out.numel() != n * elems:
        raise _lib.CtxError("gather_frames: out has %d elements, expected %d" % (out.numel(), n * elems))
    pl = _req(lut, "lut") if is_u8 else None
    if is_u8 and lut.numel() != 256:
        raise _lib.CtxError("gather_frames: lut must hold 256 floats")
    check(_lib.load().ctx_gather_frames(pb, int(is_u8), pl, bank.shape[0], elems, pi, n, _req(out, "out"), _stream()),
          "ctx_gather_frames")
    return out


def lrelu_bwd(dy, y_ref, out=None, accumulate=False, slope=LRELU_SLOPE):
    if out is None:
        out = torch.empty_like(dy)
    check(_lib.load().ctx_lrelu_bwd(_req(dy, "dy"), _req(y_ref, "y_ref"), _req(out, "out"), dy.numel(), slope,
                                    int(accumulate), _stream()), "ctx_lrelu_bwd")
    return out


def bias_grad(dy2d, db):
    """db[C] += column sums of dy (any leading shape, last dim C, dense)."""
    Cc = dy2d.shape[-1]
    M = dy2d.numel() // Cc
    check(_lib.load().ctx_bias_grad(_req(dy2d, "dy"), M, Cc, Cc, _req(db, "db"), _stream()), "ctx_bias_grad")
    return db


def l2_loss_fwd_bwd(pred, tgt, loss, dpred=None, resid=None, weight=None):
    """loss[g] += 0.5*sum (pred_g + resid - tgt)^2 ; dpred = weight[g]*(pred_g + resid - tgt)."""
    B = tgt.shape[0]
    G = pred.shape[0] // B
    per_img = tgt.numel() // B
    check(_lib.load().ctx_l2_loss_fwd_bwd(_req(pred, "pred"), _req(tgt, "tgt"), _opt(resid, "resid"), G, B, per_img,
                                          _opt(weight, "weight"), _req(loss, "loss"), _opt(dpred, "dpred"), _stream()),
          "ctx_l2_loss_fwd_bwd")


def mse_scaled_fwd_bwd(a, b, scale, inv_count, loss=None, da=None, db=None, grad_weight=1.0, accumulate=False):
    pa, rows, F, lda = _mat(a, "a")
    pb, _, _, ldb = _mat(b, "b")
    pda, ldda = (None, 0) if da is None else (_mat(da, "da")[0], _mat(da, "da")[3])
    pdb, lddb = (None, 0) if db is None else (_mat(db, "db")[0], _mat(db, "db")[3])
    check(_lib.load().ctx_mse_scaled_fwd_bwd(pa, lda, pb, ldb, rows, F, scale, inv_count, grad_weight,
                                             _opt(loss, "loss"), pda, ldda, pdb, lddb, int(accumulate), _stream()),
          "ctx_mse_scaled_fwd_bwd")


def split_skip_grad(dcat, G, B, C1, C2, dstate, dskip, state_ref=None, skip_ref=None, accumulate_skip=False,
                    slope=LRELU_SLOPE, round_tf32=False):
    pix = dcat.numel() // (G * B * (C1 + C2))
    check(_lib.load().ctx_split_skip_grad(_req(dcat, "dcat"), G, B, pix, C1, C2, _opt(state_ref, "state_ref"),
                                          _opt(skip_ref, "skip_ref"), slope, _req(dstate, "dstate"),
                                          _req(dskip, "dskip"), int(accumulate_skip), int(round_tf32), _stream()),
          "ctx_split_skip_grad")


def add_bcast(a, b, out=None):
    if out is None:
        out = torch.empty_like(a)
    Cc = a.numel() // a.shape[0]
    check(_lib.load().ctx_add_bcast(_req(a, "a"), _req(b, "b"), _req(out, "out"), a.shape[0], b.shape[0], Cc,
                                    _stream()), "ctx_add_bcast")
    return out


def im2col(x, k: int, stride: int, ld: int, out=None, round_tf32=False, padding: str = "SAME"):
    """Patch matrix [N*P*Q, ld] of a SAME / VALID conv over a few-channel NHWC tensor (columns (r,s,c), zero padded)."""
    N, H, W_, Cc = x.shape
    g = _geom_general(N, H, W_, k, k, stride, padding)
    if out is None:
        out = torch.empty((N * g.P * g.Q, ld), device=x.device, dtype=torch.float32)
    check(_lib.load().ctx_im2col(_req(x, "x"), Cc, _req(out, "out"), ld, C.byref(g), int(round_tf32), _stream()),
          "ctx_im2col")
    return out


def adam_schedule(step, lr: float, b1: float, b2: float, lr_t):
    """step (int64[1], device) += 1; lr_t (float32[1], device) = lr * sqrt(1 - b2^step) / (1 - b1^step)."""
    check(_lib.load().ctx_adam_schedule(_req(step, "step", torch.int64), float(lr), float(b1), float(b2), _req(lr_t, "lr_t"),
                                        _stream()), "ctx_adam_schedule")


def adam_step(p, g, m, v, lr_t, b1=0.9, b2=0.999, eps=1e-8, grad_scale=1.0, p_tf32=None):
    if torch.is_tensor(lr_t):                              # the rate lives on the device (CUDA-graph replays)
        check(_lib.load().ctx_adam_step_dev(_req(p, "p"), _req(g, "g"), _req(m, "m"), _req(v, "v"), p.numel(),
                                            _req(lr_t, "lr_t"), b1, b2, eps, grad_scale, _opt(p_tf32, "p_tf32"), _stream()),
              "ctx_adam_step_dev")
        return
    check(_lib.load().ctx_adam_step(_req(p, "p"), _req(g, "g"), _req(m, "m"), _req(v, "v"), p.numel(), lr_t, b1, b2,
                                    eps, grad_scale, _opt(p_tf32, "p_tf32"), _stream()), "ctx_adam_step")


def round_tf32(x, out=None):
    if out is None:
        out = torch.empty_like(x)
    check(_lib.load().ctx_round_tf32(_req(x, "x"), _req(out, "out"), x.numel(), _stream()), "ctx_round_tf32")
    return out


def transpose_round(x, batch, rows, cols, out=None, round_tf32_=True):
    if out is None:
        out = torch.empty_like(x)
    check(_lib.load().ctx_transpose_round(_req(x, "x"), _req(out, "out"), batch, rows, cols, int(round_tf32_),
                                          _stream()), "ctx_transpose_round")
    return out


def nn_err(tgt, out, nlen: int):
    """scripts/train_script.py:148.  Returns (err int64[1], idx int64[B]) device tensors."""
    B = tgt.shape[0]
    per_img = tgt.numel() // B
    dist = torch.empty((B, B), device=tgt.device, dtype=torch.float32)
    idx = torch.empty((B,), device=tgt.device, dtype=torch.int64)
    err = torch.zeros((1,), device=tgt.device, dtype=torch.int64)
    check(_lib.load().ctx_nn_err(_req(tgt, "tgt"), _req(out, "out"), B, per_img, nlen, dist.data_ptr(),
                                 idx.data_ptr(), err.data_ptr(), _stream()), "ctx_nn_err")
    return err, idx


def reward_ours(z, x, means, mean_imgs, paths: int, T: int, scale: float, rewards=None, cost=None,
                feat_weight=1.0, img_weight=1.0):
    """rllab/sampler/base.py:243-257 batched over paths.  rewards: float64 [paths, >=2T] updated in place."""
    F = means.shape[-1]
    per_img = mean_imgs.numel() // T
    rs = 0 if rewards is None else rewards.stride(0)
    check(_lib.load().ctx_reward_ours(_opt(z, "z"), _opt(x, "x"), _req(means, "means"), _req(mean_imgs, "mean_imgs"),
                                      paths, T, F, per_img, scale, feat_weight, img_weight,
                                      None if rewards is None else _req(rewards, "rewards", torch.float64), rs,
                                      _opt(cost, "cost"), _stream()), "ctx_reward_ours")


def reward_inception(feat, means, std, paths: int, T: int, rewards=None, cost=None):
    per = means.numel() // T
    rs = 0 if rewards is None else rewards.stride(0)
    check(_lib.load().ctx_reward_inception(_req(feat, "feat"), _req(means, "means"), _req(std, "std"), paths, T, per,
                                           None if rewards is None else _req(rewards, "rewards", torch.float64), rs,
                                           _opt(cost, "cost"), _stream()), "ctx_reward_inception")


def gae_tail(rewards, baselines, offsets, discount: float, gae_lambda: float):
    """rllab/sampler/base.py:258-266 on the device: float64 rewards / baselines of all paths laid end to end, int64 offsets
    [paths + 1] -> (advantages, returns), bit-identical to the reference's numpy / scipy evaluation."""
    paths = offsets.numel() - 1
    adv, ret = torch.empty_like(rewards), torch.empty_like(rewards)
    check(_lib.load().ctx_gae_tail(_req(rewards, "rewards", torch.float64), _req(baselines, "baselines", torch.float64),
                                   _req(offsets, "offsets", torch.int64), paths, float(discount), float(gae_lambda),
                                   _req(adv, "advantages", torch.float64), _req(ret, "returns", torch.float64), _stream()),
          "ctx_gae_tail")
    return adv, ret
