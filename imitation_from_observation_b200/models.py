"""Reference-facing model classes: same names, constructor arguments and attribute surface as
gym/envs/mujoco/arm_shaping.py, evaluated by :class:`compat.Session` (or used directly with device tensors).

    model = ContextSkipNew()                      # arm_shaping.py:1260
    model.build(image)                            # image: [3,B,H,W,C] placeholder / tensor   (:1272)
    model.input_z, .translated_z, .out, .out2, .recon1, .recon2, .simloss, .loss   # fetch handles (:1296-1354)

plus the functional aliases named by BASELINE.json's north_star:
    enc1 == scope "conv" (source / target encoder), enc2 == scope "conv_context" (context encoder),
    translate == scope "translate", dec == scope "deconv".
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib, ops, spec as _spec
from .network import FWD_ALL, FWD_SRC, FWD_TRANS, ContextNet

_FETCH_LEVEL = {"input_z": FWD_SRC, "translated_z": FWD_TRANS, "out": FWD_TRANS, "ctx_z": FWD_TRANS,
                "image_trans": 0, "out2": FWD_ALL, "target_z": FWD_ALL, "recon1": FWD_ALL, "recon2": FWD_ALL,
                "simloss": FWD_ALL, "loss": FWD_ALL}


class Fetch:
    """Symbolic handle to a graph output (what a TF tensor attribute is in the reference)."""

    def __init__(self, model: "_ContextModel", name: str):
        self.model, self.name, self.level = model, name, _FETCH_LEVEL[name]

    def __repr__(self):
        return "<Fetch %s.%s>" % (type(self.model).__name__, self.name)


class _ContextModel:
    """Common machinery; subclasses only pick the :class:`spec.NetSpec`."""

    precision = "mixed"       # class-level default; override per instance or via build(..., precision=)

    def _make_spec(self, height: int, width: int, c_dim: int) -> _spec.NetSpec:
        raise NotImplementedError

    # -- reference API ---------------------------------------------------------------------------------------------
    def build(self, image, device="cuda", precision: Optional[str] = None, seed: int = 0,
              global_batch: Optional[int] = None, ablation_type: str = "None", scope: str = ""):
        """`image` only has to expose `.shape == [3,B,H,W,C]` (a compat placeholder, a torch tensor or an ndarray)."""
        shape = [int(s) for s in image.shape]
        if len(shape) != 5 or shape[0] != 3:
            raise _lib.CtxError("build() expects a [3,B,H,W,C] image tensor, got %s" % (shape,))
        _, B, H, W, C = shape
        self.batch_size = B
        self.output_height, self.output_width = H, W
        self.spec = self._make_spec(H, W, C)
        self.scope = scope                       # tf.variable_scope prefix, e.g. "contextmodel/" (train_script.py:120)
        self.placeholder = image
        self.net = ContextNet(self.spec, B, device=device, precision=precision or self.precision, seed=seed,
                              global_batch=global_batch, ablation_type=ablation_type)
        for name in _FETCH_LEVEL:
            setattr(self, name, Fetch(self, name))
        if hasattr(image, "_register_consumer"):
            image._register_consumer(self)
        return self

    def trainable_variables(self) -> List[str]:
        return [self.scope + n for n in self.net.shapes]

    # -- functional surface (device tensors in, device tensors out) ------------------------------------------------
    def _prep(self, frames: torch.Tensor) -> torch.Tensor:
        if frames.dtype == torch.uint8:
            return ops.u8_normalize(frames.contiguous())
        return frames.contiguous().float()

    def _encode(self, scope: str, frames: torch.Tensor, hz_act: bool):
        net = self.net
        x = self._prep(frames)
        n = x.shape[0]
        hs = [torch.empty((n,) + net.sizes[j + 1] + (net.chans[j + 1],), device=x.device) for j in range(net.L)]
        h4 = torch.empty((n, net.F), device=x.device)
        z = torch.empty((n, net.F), device=x.device)
        net._fwd_encoder(scope, x, hs, h4)
        net._fwd_hz(scope, h4, z.data_ptr(), net.F, act=hz_act)
        return z, hs

    def enc1(self, frames: torch.Tensor, return_activations: bool = False):
        """Scope "conv": z = lrelu(hz_lin(lrelu(h4_lin(conv stack)))) — arm_shaping.py:1290-1297."""
        z, hs = self._encode("conv", frames, True)
        return (z, hs) if return_activations else z

    def enc2(self, frames: torch.Tensor):
        """Scope "conv_context" (or the shared encoder): returns (ctx_z, [h0,h1,h2,h3]) — arm_shaping.py:1282-1288."""
        return self._encode(self.net.ctx_scope, frames, self.spec.ctx_hz_lrelu)

    def translate(self, z_src: torch.Tensor, z_ctx: torch.Tensor) -> torch.Tensor:
        """arm_shaping.py:1309-1312."""
        net, F = self.net, self.net.F
        cat = torch.cat([z_src, z_ctx], 1).contiguous()
        n = cat.shape[0]
        t0 = torch.empty((n, F), device=cat.device)
        tz = torch.empty((n, F), device=cat.device)
        net._gemm(False, False, n, F, 2 * F, cat.data_ptr(), 2 * F, net.Wf["translate/trans_h0/Matrix"].data_ptr(), F,
                  t0.data_ptr(), F, bias=net.P["translate/trans_h0/bias"], act=True)
        net._gemm(False, False, n, F, F, t0.data_ptr(), F, net.Wf["translate/trans_z/Matrix"].data_ptr(), F,
                  tz.data_ptr(), F, bias=net.P["translate/trans_z/bias"], act=False)
        return tz

    def dec(self, z: torch.Tensor, skips: Sequence[torch.Tensor], ctx_frames: Optional[torch.Tensor] = None):
        """Scope "deconv" — arm_shaping.py:1314-1329.  `skips` = enc2 activations (batch may be smaller: n % batch)."""
        net = self.net
        n = z.shape[0]
        d = [torch.empty((n,) + net.sizes[net.L] + (net.chans[net.L],), device=z.device)]
        for L in net.dec:
            d.append(torch.empty((n,) + tuple(L["out_hw"]) + (L["c_out"],), device=z.device))
        net._fwd_decoder(z.contiguous(), skips, d)
        out = d[net.L]
        if self.spec.residual_out:
            if ctx_frames is None:
                raise _lib.CtxError("dec(): this model adds the context input to the output; pass ctx_frames")
            out = ops.add_bcast(out, self._prep(ctx_frames))
        return out

    def forward(self, src, ctx, tgt) -> dict:
        """Direct (non-session) evaluation of the whole graph on device tensors [B,H,W,C]."""
        self.net.set_input(torch.stack([src, ctx, tgt]))
        self.net.forward(FWD_ALL)
        return {k: self.net.fetch(k) for k in ("input_z", "translated_z", "out", "out2", "recon1", "recon2",
                                               "simloss", "loss")}

    def loss_and_grads(self):
        self.net.zero_grad()
        self.net.forward(FWD_ALL)
        self.net.backward()
        return self.net.fetch("loss"), self.net.G


class ContextSkipNew(_ContextModel):
    """arm_shaping.py:1260-1354."""

    def __init__(self, gf_dim=64, df_dim=64, gfc_dim=1024, dfc_dim=1024, c_dim=3, featsize=1024):
        self.gf_dim, self.df_dim, self.c_dim = gf_dim, df_dim, c_dim
        self.gfc_dim, self.dfc_dim = gfc_dim, dfc_dim
        self.featsize = featsize               # hard-coded 1024 in the reference (:1277); exposed for small tests
        if gf_dim != df_dim:
            raise _lib.CtxError("ContextSkipNew: skip concatenation requires gf_dim == df_dim (as in the reference)")

    def _make_spec(self, H, W, C):
        s = _spec.context_skip_new(H, W, C, self.df_dim, self.featsize)
        # the reference sizes the decoder with int(s/2^k) (:1316-1319); SAME convs give ceil — they must agree
        for k, (h, w) in enumerate(s.enc_sizes()):
            if (h, w) != (int(H / 2 ** (k + 1)), int(W / 2 ** (k + 1))):
                raise _lib.CtxError("ContextSkipNew: %dx%d is not divisible by 16 the way the reference needs" % (H, W))
        return s


class ContextAEReal(_ContextModel):
    """arm_shaping.py:1599-1684 (dropout sites are identity: keep_prob = 1.0, :1476)."""

    def __init__(self, gf_dim=64, df_dim=64, gfc_dim=1024, dfc_dim=1024, c_dim=3, featsize=100,
                 filters=(32, 16, 16, 8)):
        self.gf_dim, self.df_dim, self.c_dim = gf_dim, df_dim, c_dim
        self.featsize, self.filters = featsize, tuple(filters)

    def _make_spec(self, H, W, C):
        return _spec.context_ae_real(H, W, C, self.featsize, self.filters)


class ContextAEInception2(_ContextModel):
    """arm_shaping.py:1786-1894."""

    def __init__(self, strides, kernels, filters, featsize=1024):
        self.strides, self.kernels, self.filters = list(strides), list(kernels), list(filters)
        self.featsize = featsize

    def _make_spec(self, H, W, C):
        return _spec.context_ae_inception2(H, W, C, self.strides, self.kernels, self.filters, self.featsize)


class ContextAEInception(_ContextModel):
    """arm_shaping.py:1687-1783 — the legacy 5x5 stride-2 feature-map model (note: train_script.py:48-49 maps the *name*
    "ContextAEInception" to ContextAEInception2, and so does MODEL_REGISTRY; this class is the one arm_shaping defines)."""

    def __init__(self, gf_dim=64, df_dim=64, gfc_dim=1024, dfc_dim=1024, c_dim=3, featsize=1024):
        self.gf_dim, self.df_dim, self.c_dim, self.featsize = gf_dim, df_dim, c_dim, featsize
        if gf_dim != df_dim:
            raise _lib.CtxError("ContextAEInception: skip concatenation requires gf_dim == df_dim (as in the reference)")

    def _make_spec(self, H, W, C):
        return _spec.context_ae_inception(H, W, C, self.df_dim, self.featsize)


class ContextSkipDeep(_ContextModel):
    """BASELINE.json configs[3]: "Striker 128x128x3, deeper 6-layer enc/dec".  NOT a reference class — the reference's
    builders are 4 layers deep; this is ContextSkipNew's graph with one conv / transposed conv per entry of `filters`
    (spec.context_skip_deep).  Same kernels, same fetch surface, same checkpoint naming scheme
    (`h<L>_lin`, `d_h1..d_h<L>`)."""

    def __init__(self, filters=(64, 128, 256, 512, 512, 512), kernels=None, strides=None, featsize=1024):
        self.filters = tuple(filters)
        self.kernels = tuple(kernels) if kernels else None
        self.strides = tuple(strides) if strides else None
        self.featsize = featsize

    def _make_spec(self, H, W, C):
        return _spec.context_skip_deep(H, W, C, self.filters, self.kernels, self.strides, self.featsize)


MODEL_REGISTRY = {"ContextSkipNew": ContextSkipNew, "ContextAEReal": ContextAEReal,
                  "ContextAEInception": ContextAEInception2,      # train_script.py:48-49 maps this name to ...2
                  "ContextAEInception2": ContextAEInception2,
                  "ContextSkipDeep": ContextSkipDeep}             # not in the reference (BASELINE configs[3])
