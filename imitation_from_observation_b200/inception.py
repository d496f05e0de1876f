"""InceptionV3 feature extractor (inference only) for the Inception reward modes.

Reference: nets/inception_v3.py:29-416 (`inception_v3_base`) + the `PreLogits` pooling of `inception_v3` (:484-496)
under `inception_arg_scope` (nets/inception_utils.py:31-71): every conv is bias-free, followed by inference batch norm
*without gamma* — y = (x - moving_mean) / sqrt(moving_variance + 1e-3) + beta — and ReLU.  The reference uses it with
`is_training=False` as (a) the baseline reward featurizer (rllab/sampler/base.py:69-111,179-189: endpoints `Mixed_7c`
or `PreLogits`) and (b) the frozen front-end of `ContextAEInception2` (base.py:121-132).

B200 design: the batch norm is folded into the conv once per weight load (w' = w / sqrt(var + eps) per output channel,
b' = beta - mean / sqrt(var + eps)), so every layer is one launch of the same implicit-GEMM kernel as the context model
with a fused bias + ReLU epilogue (`act` with slope 0); pools and channel concatenations are HBM-bound kernels.
The architecture below is a table: each entry is (scope, out_channels, kh, kw, stride, padding).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops

BN_EPS = 1e-3                                    # inception_arg_scope(batch_norm_epsilon=0.001)
ENDPOINTS = ("Conv2d_1a_3x3", "Conv2d_2a_3x3", "Conv2d_2b_3x3", "MaxPool_3a_3x3", "Conv2d_3b_1x1", "Conv2d_4a_3x3",
             "MaxPool_5a_3x3", "Mixed_5b", "Mixed_5c", "Mixed_5d", "Mixed_6a", "Mixed_6b", "Mixed_6c", "Mixed_6d",
             "Mixed_6e", "Mixed_7a", "Mixed_7b", "Mixed_7c", "PreLogits")

S, V = "SAME", "VALID"


def _c(scope, co, kh, kw, stride=1, pad=S):
    return ("conv", scope, co, kh, kw, stride, pad)


def _stem():
    """inception_v3.py:100-141 (arg_scope default: stride 1, VALID)."""
    return [_c("Conv2d_1a_3x3", 32, 3, 3, 2, V), _c("Conv2d_2a_3x3", 32, 3, 3, 1, V), _c("Conv2d_2b_3x3", 64, 3, 3, 1, S),
            ("maxpool", "MaxPool_3a_3x3", 3, 2, V), _c("Conv2d_3b_1x1", 80, 1, 1, 1, V), _c("Conv2d_4a_3x3", 192, 3, 3, 1, V),
            ("maxpool", "MaxPool_5a_3x3", 3, 2, V)]


def _mixed_5(pool_ch, b1_names=("Conv2d_0a_1x1", "Conv2d_0b_5x5")):
    """35x35 blocks, inception_v3.py:147-216."""
    return [[_c("Branch_0/Conv2d_0a_1x1", 64, 1, 1)],
            [_c("Branch_1/" + b1_names[0], 48, 1, 1), _c("Branch_1/" + b1_names[1], 64, 5, 5)],
            [_c("Branch_2/Conv2d_0a_1x1", 64, 1, 1), _c("Branch_2/Conv2d_0b_3x3", 96, 3, 3), _c("Branch_2/Conv2d_0c_3x3", 96, 3, 3)],
            [("avgpool", "Branch_3/AvgPool_0a_3x3", 3, 1, S), _c("Branch_3/Conv2d_0b_1x1", pool_ch, 1, 1)]]


def _mixed_6(c7):
    """17x17 blocks with factorised 7x7, inception_v3.py:241-356."""
    return [[_c("Branch_0/Conv2d_0a_1x1", 192, 1, 1)],
            [_c("Branch_1/Conv2d_0a_1x1", c7, 1, 1), _c("Branch_1/Conv2d_0b_1x7", c7, 1, 7), _c("Branch_1/Conv2d_0c_7x1", 192, 7, 1)],
            [_c("Branch_2/Conv2d_0a_1x1", c7, 1, 1), _c("Branch_2/Conv2d_0b_7x1", c7, 7, 1), _c("Branch_2/Conv2d_0c_1x7", c7, 1, 7),
             _c("Branch_2/Conv2d_0d_7x1", c7, 7, 1), _c("Branch_2/Conv2d_0e_1x7", 192, 1, 7)],
            [("avgpool", "Branch_3/AvgPool_0a_3x3", 3, 1, S), _c("Branch_3/Conv2d_0b_1x1", 192, 1, 1)]]


def _mixed_7(second_3x1):
    """8x8 blocks with split 1x3 / 3x1 heads, inception_v3.py:381-413.  ("fork", [a, b]) concatenates both heads."""
    return [[_c("Branch_0/Conv2d_0a_1x1", 320, 1, 1)],
            [_c("Branch_1/Conv2d_0a_1x1", 384, 1, 1),
             ("fork", [_c("Branch_1/Conv2d_0b_1x3", 384, 1, 3), _c("Branch_1/" + second_3x1, 384, 3, 1)])],
            [_c("Branch_2/Conv2d_0a_1x1", 448, 1, 1), _c("Branch_2/Conv2d_0b_3x3", 384, 3, 3),
             ("fork", [_c("Branch_2/Conv2d_0c_1x3", 384, 1, 3), _c("Branch_2/Conv2d_0d_3x1", 384, 3, 1)])],
            [("avgpool", "Branch_3/AvgPool_0a_3x3", 3, 1, S), _c("Branch_3/Conv2d_0b_1x1", 192, 1, 1)]]


BLOCKS = [
    ("Mixed_5b", _mixed_5(32)),
    ("Mixed_5c", _mixed_5(64, ("Conv2d_0b_1x1", "Conv_1_0c_5x5"))),       # (sic) scope names of the reference, :187-189
    ("Mixed_5d", _mixed_5(64)),
    ("Mixed_6a", [[_c("Branch_0/Conv2d_1a_1x1", 384, 3, 3, 2, V)],                       # :222-236
                  [_c("Branch_1/Conv2d_0a_1x1", 64, 1, 1), _c("Branch_1/Conv2d_0b_3x3", 96, 3, 3),
                   _c("Branch_1/Conv2d_1a_1x1", 96, 3, 3, 2, V)],
                  [("maxpool", "Branch_2/MaxPool_1a_3x3", 3, 2, V)]]),
    ("Mixed_6b", _mixed_6(128)), ("Mixed_6c", _mixed_6(160)), ("Mixed_6d", _mixed_6(160)), ("Mixed_6e", _mixed_6(192)),
    ("Mixed_7a", [[_c("Branch_0/Conv2d_0a_1x1", 192, 1, 1), _c("Branch_0/Conv2d_1a_3x3", 320, 3, 3, 2, V)],   # :359-378
                  [_c("Branch_1/Conv2d_0a_1x1", 192, 1, 1), _c("Branch_1/Conv2d_0b_1x7", 192, 1, 7),
                   _c("Branch_1/Conv2d_0c_7x1", 192, 7, 1), _c("Branch_1/Conv2d_1a_3x3", 192, 3, 3, 2, V)],
                  [("maxpool", "Branch_2/MaxPool_1a_3x3", 3, 2, V)]]),
    ("Mixed_7b", _mixed_7("Conv2d_0b_3x1")), ("Mixed_7c", _mixed_7("Conv2d_0c_3x1")),
]


def conv_table(c_in: int = 3) -> List[Tuple[str, int, int, int, int]]:
    """Every conv of the base network in execution order: (scope, kh, kw, Cin, Cout) — shapes only, no tensors."""
    out: List[Tuple[str, int, int, int, int]] = []

    def run(ops_list, c, prefix):
        for op in ops_list:
            if op[0] == "conv":
                out.append((prefix + op[1], op[3], op[4], c, op[2]))
                c = op[2]
            elif op[0] == "fork":
                tot = 0
                for h in op[1]:
                    out.append((prefix + h[1], h[3], h[4], c, h[2]))
                    tot += h[2]
                c = tot
        return c

    c = run(_stem(), c_in, "InceptionV3/")
    for name, branches in BLOCKS:
        c = sum(run(br, c, "InceptionV3/%s/" % name) for br in branches)
    return out


def param_shapes(c_in: int = 3) -> Dict[str, Tuple[int, ...]]:
    """slim variable names -> shapes (weights HWIO; BatchNorm beta / moving_mean / moving_variance, no gamma)."""
    shapes: Dict[str, Tuple[int, ...]] = {}
    for scope, kh, kw, ci, co in conv_table(c_in):
        shapes[scope + "/weights"] = (kh, kw, ci, co)
        for v in ("beta", "moving_mean", "moving_variance"):
            shapes[scope + "/BatchNorm/" + v] = (co,)
    return shapes


class InceptionV3:
    """Inference-only InceptionV3 bound to a device.  `forward(images)` returns the requested end points as NHWC
    device tensors; images are uint8 (normalised like base.py:72-76: x/255, -0.5, *2) or already-normalised fp32."""

    def __init__(self, device="cuda", precision: str = "mixed", seed: Optional[int] = None):
        _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CtxError("InceptionV3 needs a CUDA device: the B200 path has no CPU fallback")
        if precision not in ("mixed", "tf32x3", "tf32", "fp32"):
            raise ValueError("precision must be 'mixed', 'tf32x3', 'tf32' or 'fp32'")
        # inference only: "mixed" and "tf32x3" both mean the fp32-faithful 3-pass forward
        self.impl = {"mixed": _lib.IMPL_AUTO, "tf32x3": _lib.IMPL_AUTO, "tf32": _lib.IMPL_FAST, "fp32": _lib.IMPL_SIMT}[precision]
        self.shapes = param_shapes()
        self.params: Dict[str, torch.Tensor] = {}
        self.folded: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        self._bufs: Dict[str, torch.Tensor] = {}
        if seed is not None:
            self.init_random(seed)

    # ---- parameters ------------------------------------------------------------------------------------------------
    def init_random(self, seed: int = 0):
        """Variance-scaling weights and plausible BN statistics (no checkpoint ships with the reference)."""
        g = torch.Generator().manual_seed(seed)
        sd = {}
        for name, shp in self.shapes.items():
            if name.endswith("/weights"):
                fan_in = shp[0] * shp[1] * shp[2]
                sd[name] = torch.randn(shp, generator=g) * (2.0 / fan_in) ** 0.5
            elif name.endswith("moving_variance"):
                sd[name] = torch.rand(shp, generator=g) * 0.5 + 0.75
            else:
                sd[name] = torch.randn(shp, generator=g) * 0.1
        self.load_state_dict(sd)

    def load_state_dict(self, sd: Dict[str, "np.ndarray | torch.Tensor"]):
        for name, shp in self.shapes.items():
            if name not in sd:
                raise KeyError("missing parameter %s" % name)
            v = sd[name]
            v = torch.as_tensor(np.asarray(v)) if not torch.is_tensor(v) else v
            if tuple(v.shape) != tuple(shp):
                raise ValueError("parameter %s has shape %s, expected %s" % (name, tuple(v.shape), shp))
            self.params[name] = v.to(device=self.device, dtype=torch.float32).contiguous()
        self._fold()

    def state_dict(self) -> Dict[str, np.ndarray]:
        return {k: v.detach().cpu().numpy().copy() for k, v in self.params.items()}

    def load_tf_checkpoint(self, prefix: str):
        """Load slim's `inception_v3.ckpt` (rllab/sampler/base.py:130-132 restores it with tf.train.Saver).  The published
        file is a single-file V1 checkpoint (SavedTensorSlices in one table); a V2 bundle `<prefix>.index` + `<prefix>.data-*`
        is read as well (tf_checkpoint.read_checkpoint dispatches).  Variables InceptionV3/<scope>/{weights, BatchNorm/beta,
        BatchNorm/moving_mean, BatchNorm/moving_variance}; the logits / aux-logits heads in the file are ignored."""
        from . import tf_checkpoint
        self.load_state_dict(tf_checkpoint.read_checkpoint(prefix, list(self.shapes)))

    def _fold(self):
        """Fold the inference batch norm into each conv (host-side torch arithmetic on parameters, once per load)."""
        self.folded = {}
        for scope, kh, kw, ci, co in conv_table():
            w = self.params[scope + "/weights"]
            inv = torch.rsqrt(self.params[scope + "/BatchNorm/moving_variance"] + BN_EPS)
            b = self.params[scope + "/BatchNorm/beta"] - self.params[scope + "/BatchNorm/moving_mean"] * inv
            self.folded[scope] = ((w * inv.view(1, 1, 1, co)).contiguous(), b.contiguous())

    # ---- execution -------------------------------------------------------------------------------------------------
    def _buf(self, name: str, shape) -> torch.Tensor:
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(tuple(shape), device=self.device, dtype=torch.float32)
            self._bufs[name] = t
        return t

    def _conv(self, x, prefix, op, into=None, c_off=0):
        _, scope, co, kh, kw, stride, pad = op
        w, b = self.folded[prefix + scope]
        N, H, W_, C = x.shape
        g = ops._geom_general(N, H, W_, kh, kw, stride, pad)
        if into is not None:                             # last conv of a branch: straight into the block's concat
            return ops.conv2d_general(x, w, b, stride=stride, padding=pad, act=True, slope=0.0, impl=self.impl,
                                      into=into, c_off=c_off)
        out = self._buf(prefix + scope, (N, g.P, g.Q, co))
        few = C % 32 != 0 and C * kh * kw <= 256 and co % 4 == 0 and kh == kw
        if few and self.impl != _lib.IMPL_SIMT and ops.tc_available():
            # few-channel first layer (3 -> 32): patch matrix + dense tensor-core GEMM, like the context model's h0
            K = kh * kw * C
            ld = -(-K // 4) * 4
            pt = ops.im2col(x, kh, stride, ld, out=self._buf(prefix + scope + ":patches", (N * g.P * g.Q, ld)), padding=pad)
            ops.gemm(pt[:, :K], w.view(K, co), out=out.view(-1, co), bias=b, act=True, slope=0.0, impl=self.impl)
            return out
        return ops.conv2d_general(x, w, b, stride=stride, padding=pad, act=True, slope=0.0, out=out, impl=self.impl)

    def _run(self, x, ops_list, prefix, into=None, c_off=0):
        """Run a chain of ops.  With `into`, the chain's result is written to channels [c_off, ...) of `into` (the
        block's tf.concat result): convs write there directly, a trailing pool is copied."""
        last = len(ops_list) - 1
        for i, op in enumerate(ops_list):
            dst = into if i == last else None
            if op[0] == "conv":
                x = self._conv(x, prefix, op, dst, c_off)
            elif op[0] in ("maxpool", "avgpool"):
                _, scope, k, stride, pad = op
                N, H, W_, C = x.shape
                g = ops._geom_general(N, H, W_, k, k, stride, pad)
                x = ops.pool2d(x, k, stride, pad, "max" if op[0] == "maxpool" else "avg",
                               out=self._buf(prefix + scope, (N, g.P, g.Q, C)))
                if dst is not None:
                    ops.copy_channels(x, dst, c_off)
            elif op[0] == "fork":
                if dst is not None:
                    off = c_off
                    for h in op[1]:
                        self._conv(x, prefix, h, dst, off)
                        off += h[2]
                    x = dst
                else:
                    heads = [self._conv(x, prefix, h) for h in op[1]]
                    x = self._concat(prefix + op[1][0][1] + ":fork", heads)
        return x

    @staticmethod
    def _chain_out(ops_list, c_in):
        """(channels, spatial stride factor is implicit) -> output channel count of a chain."""
        c = c_in
        for op in ops_list:
            if op[0] == "conv":
                c = op[2]
            elif op[0] == "fork":
                c = sum(h[2] for h in op[1])
        return c

    def _concat(self, name, parts: Sequence[torch.Tensor]) -> torch.Tensor:
        N, H, W_ = parts[0].shape[:3]
        ctot = sum(p.shape[3] for p in parts)
        out = self._buf(name, (N, H, W_, ctot))
        off = 0
        for p in parts:
            ops.copy_channels(p, out, off)
            off += p.shape[3]
        return out

    def forward(self, images: torch.Tensor, endpoints: Sequence[str] = ("Mixed_7c",)) -> Dict[str, torch.Tensor]:
        for e in endpoints:
            if e not in ENDPOINTS:
                raise KeyError("unknown end point %r" % e)
        if not self.folded:
            raise _lib.CtxError("InceptionV3 has no parameters: call load_state_dict() or init_random()")
        if images.dtype == torch.uint8:
            x = ops.u8_normalize(images.contiguous(), out=self._buf("input", tuple(images.shape)))
        else:
            x = images.contiguous()
        want, got = set(endpoints), {}
        last = max(ENDPOINTS.index(e) for e in endpoints)
        pos = 0
        for op in _stem():
            x = self._run(x, [op], "InceptionV3/")
            if op[1] in want:
                got[op[1]] = x
            pos += 1
            if pos > last:
                return got
        for name, branches in BLOCKS:
            N, H, W_, C = x.shape
            chans = [self._chain_out(br, C) for br in branches]
            # output grid of the block = that of its first branch's geometry: every branch of a block ends on it
            stride2 = any(op[0] in ("conv",) and op[5] == 2 for op in branches[0]) or \
                any(op[0] == "maxpool" for br in branches for op in br)
            P = (H - 3) // 2 + 1 if stride2 else H
            Q = (W_ - 3) // 2 + 1 if stride2 else W_
            blk = self._buf("InceptionV3/%s" % name, (N, P, Q, sum(chans)))
            off = 0
            for br, c in zip(branches, chans):
                self._run(x, br, "InceptionV3/%s/" % name, into=blk, c_off=off)
                off += c
            x = blk
            if name in want:
                got[name] = x
            pos += 1
            if pos > last:
                return got
        # PreLogits: 8x8 (whole-map) VALID average pool; dropout is the identity at inference (inception_v3.py:484-496)
        N, H, W_, C = x.shape
        got["PreLogits"] = ops.pool2d(x, H, 1, "VALID", "avg", out=self._buf("PreLogits", (N, 1, 1, C)))
        return got
