"""Static description of the context-translation networks (pure Python, no torch, no CUDA).

The reference builds three wired-up graph variants in gym/envs/mujoco/arm_shaping.py:

* ``ContextSkipNew``       arm_shaping.py:1260-1354  (push / reach / strike / throw)
* ``ContextAEReal``        arm_shaping.py:1599-1684  (real / sweep)
* ``ContextAEInception2``  arm_shaping.py:1786-1894  (feature-map model behind InceptionV3)

All three are the same topology: L = 4 conv encoder layers + 2 dense layers, a 2-layer translator on
``[src_z ‖ ctx_z]``, and a decoder (dense + L transposed convs) whose every level is channel-concatenated
with the *context* encoder's activation of the same resolution.  They differ only in the numbers held by
:class:`NetSpec`.  The depth L is the length of ``filters``: the reference only wires up L = 4; BASELINE.json's
configs[3] ("Striker 128x128, deeper 6-layer enc/dec") is the same graph with L = 6 (:func:`context_skip_deep`), with
the dense layer after the last conv named ``h<L>_lin`` and the transposed convs ``d_h1 .. d_h<L>``, which for L = 4
are exactly the reference's names.  Variable names / layouts follow the TF checkpoint keys (arm_shaping.py:25,29,52,54,67,79):

    {conv,conv_context,translate,deconv}/{h0_conv..h3_conv}/w        [kh,kw,Cin,Cout]   (HWIO)
    .../{d_h1..d_h4}/w                                              [kh,kw,Cout,Cin]
    .../{h4_lin,hz_lin,trans_h0,trans_z,d_h0_lin}/Matrix            [in,out]
    .../biases | .../bias                                           [Cout] | [out]
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple


def same_out(size: int, stride: int) -> int:
    """TF 'SAME' output extent: ceil(in / stride)."""
    return -(-size // stride)


def same_pad(size: int, k: int, stride: int) -> Tuple[int, int]:
    """TF 'SAME' padding (lo, hi): total = max((out-1)*s + k - in, 0); lo = total//2 (extra goes bottom/right)."""
    out = same_out(size, stride)
    total = max((out - 1) * stride + k - size, 0)
    lo = total // 2
    return lo, total - lo


@dataclass(frozen=True)
class NetSpec:
    name: str
    height: int
    width: int
    c_dim: int
    strides: Tuple[int, ...]
    kernels: Tuple[int, ...]
    filters: Tuple[int, ...]
    featsize: int
    separate_ctx_encoder: bool      # ContextSkipNew / Inception2: scope "conv_context" has its own weights
    ctx_hz_lrelu: bool              # ContextSkipNew: ctx hz_lin is linear (arm_shaping.py:1288); others lrelu
    residual_out: bool              # Inception2: out = dec(.) + ctx input (arm_shaping.py:1890-1891)

    def __post_init__(self):
        if not (len(self.strides) == len(self.kernels) == len(self.filters) >= 1):
            raise ValueError("strides, kernels and filters must have the same, non-zero length")

    # ---- derived geometry -------------------------------------------------------------------------------
    @property
    def depth(self) -> int:
        """Number of encoder convs = number of decoder transposed convs (4 in every reference model)."""
        return len(self.filters)

    @property
    def lin_name(self) -> str:
        """Dense layer after the last conv: `h4_lin` for the reference's depth."""
        return "h%d_lin" % self.depth

    def enc_sizes(self) -> List[Tuple[int, int]]:
        """Spatial size after each encoder conv (SAME): [(h0,w0),...,(h_{L-1},w_{L-1})]."""
        h, w = self.height, self.width
        out = []
        for s in self.strides:
            h, w = same_out(h, s), same_out(w, s)
            out.append((h, w))
        return out

    def enc_channels(self) -> List[int]:
        return [self.c_dim] + list(self.filters)

    def flat_dim(self) -> int:
        h3, w3 = self.enc_sizes()[-1]
        return h3 * w3 * self.filters[-1]

    def dec_layers(self) -> List[dict]:
        """Decoder transposed-conv layers d_h1..d_h<L>, in execution order.

        Level i (1..L) consumes [state ‖ ctx skip h_{L-i}] at the resolution of encoder output L-i and
        produces the resolution of encoder output L-1-i (the image for i=L), mirroring encoder layer L-i's
        kernel and stride (arm_shaping.py:1323-1329, 1662-1670, 1834-1848).
        """
        sizes = [(self.height, self.width)] + self.enc_sizes()   # sizes[j] = input of encoder layer j
        chans = self.enc_channels()                               # chans[j] = channels at sizes[j]
        layers = []
        for i in range(1, self.depth + 1):
            e = self.depth - i                                             # mirrored encoder layer
            layers.append(dict(
                name="d_h%d" % i, k=self.kernels[e], stride=self.strides[e],
                in_hw=sizes[e + 1], out_hw=sizes[e],
                c_state=self.filters[e], c_skip=self.filters[e], c_out=chans[e],
                skip="h%d" % e,
            ))
        return layers

    # ---- parameters -------------------------------------------------------------------------------------
    def param_shapes(self) -> Dict[str, Tuple[int, ...]]:
        """Ordered name -> shape.  Order = the order gradients complete in backward (decoder first), so the
        flat gradient buffer can be all-reduced in contiguous buckets while backward is still running."""
        shapes: Dict[str, Tuple[int, ...]] = {}
        F = self.featsize
        # decoder
        for L in reversed(self.dec_layers()):
            shapes["deconv/%s/w" % L["name"]] = (L["k"], L["k"], L["c_out"], L["c_state"] + L["c_skip"])
            shapes["deconv/%s/biases" % L["name"]] = (L["c_out"],)
        shapes["deconv/d_h0_lin/Matrix"] = (F, self.flat_dim())
        shapes["deconv/d_h0_lin/bias"] = (self.flat_dim(),)
        # translator
        shapes["translate/trans_z/Matrix"] = (F, F)
        shapes["translate/trans_z/bias"] = (F,)
        shapes["translate/trans_h0/Matrix"] = (2 * F, F)
        shapes["translate/trans_h0/bias"] = (F,)
        # encoders
        scopes = ["conv"] + (["conv_context"] if self.separate_ctx_encoder else [])
        ch = self.enc_channels()
        for sc in scopes:
            shapes["%s/hz_lin/Matrix" % sc] = (F, F)
            shapes["%s/hz_lin/bias" % sc] = (F,)
            shapes["%s/%s/Matrix" % (sc, self.lin_name)] = (self.flat_dim(), F)
            shapes["%s/%s/bias" % (sc, self.lin_name)] = (F,)
            for j in reversed(range(self.depth)):
                k = self.kernels[j]
                shapes["%s/h%d_conv/w" % (sc, j)] = (k, k, ch[j], ch[j + 1])
                shapes["%s/h%d_conv/biases" % (sc, j)] = (ch[j + 1],)
        return shapes

    def num_params(self) -> int:
        n = 0
        for shp in self.param_shapes().values():
            m = 1
            for d in shp:
                m *= d
            n += m
        return n

    def fwd_macs_per_triple(self) -> int:
        """Algorithmic multiply-accumulates of one forward pass over one (src, ctx, tgt) triple."""
        macs = 0
        sizes = self.enc_sizes()
        ch = self.enc_channels()
        enc = 0
        for j in range(self.depth):
            h, w = sizes[j]
            enc += h * w * ch[j + 1] * self.kernels[j] ** 2 * ch[j]
        enc += self.flat_dim() * self.featsize + self.featsize ** 2
        macs += 3 * enc
        macs += 2 * self.featsize * self.featsize + self.featsize * self.featsize     # translator
        dec = self.featsize * self.flat_dim()
        for L in self.dec_layers():
            # transposed conv == adjoint of a conv whose *output* is in_hw: MACs = in_h*in_w*k*k*Cin*Cout
            dec += L["in_hw"][0] * L["in_hw"][1] * L["k"] ** 2 * (L["c_state"] + L["c_skip"]) * L["c_out"]
        macs += 2 * dec
        return macs


def context_skip_new(height: int = 64, width: int = 64, c_dim: int = 3, df_dim: int = 64,
                     featsize: int = 1024) -> NetSpec:
    """arm_shaping.py:1260-1354.  5x5 stride-2 everywhere, filters df,2df,4df,8df, separate ctx encoder."""
    return NetSpec("ContextSkipNew", height, width, c_dim, (2, 2, 2, 2), (5, 5, 5, 5),
                   (df_dim, 2 * df_dim, 4 * df_dim, 8 * df_dim), featsize,
                   separate_ctx_encoder=True, ctx_hz_lrelu=False, residual_out=False)


def context_ae_real(height: int = 36, width: int = 64, c_dim: int = 3, featsize: int = 100,
                    filters=(32, 16, 16, 8)) -> NetSpec:
    """arm_shaping.py:1599-1684.  One shared encoder for all three frames, strides 1/2/1/2."""
    return NetSpec("ContextAEReal", height, width, c_dim, (1, 2, 1, 2), (5, 5, 5, 5), tuple(filters), featsize,
                   separate_ctx_encoder=False, ctx_hz_lrelu=True, residual_out=False)


def context_ae_inception2(height: int, width: int, c_dim: int, strides, kernels, filters,
                          featsize: int = 1024) -> NetSpec:
    """arm_shaping.py:1786-1894.  Generic (strides, kernels, filters) on a feature map, residual output."""
    return NetSpec("ContextAEInception2", height, width, c_dim, tuple(strides), tuple(kernels), tuple(filters),
                   featsize, separate_ctx_encoder=True, ctx_hz_lrelu=True, residual_out=True)


def context_ae_inception(height: int = 35, width: int = 35, c_dim: int = 288, df_dim: int = 64,
                         featsize: int = 1024) -> NetSpec:
    """arm_shaping.py:1687-1783 (legacy; the strike-inception notebook runs it on Mixed_5d, 35x35x288).  ContextSkipNew's
    graph — 5x5 stride-2, filters df..8df, linear context hz (:1711) — on a feature map, with the residual output of
    :1778-1779 and every decoder size taken from the encoder's actual shapes (:1739-1743), so odd extents work."""
    return NetSpec("ContextAEInception", height, width, c_dim, (2, 2, 2, 2), (5, 5, 5, 5),
                   (df_dim, 2 * df_dim, 4 * df_dim, 8 * df_dim), featsize,
                   separate_ctx_encoder=True, ctx_hz_lrelu=False, residual_out=True)


def context_skip_deep(height: int = 128, width: int = 128, c_dim: int = 3,
                      filters=(64, 128, 256, 512, 512, 512), kernels=None, strides=None,
                      featsize: int = 1024) -> NetSpec:
    """BASELINE.json configs[3] ("Striker 128x128x3, deeper 6-layer enc/dec").  Not in the reference: its builders
    stop at 4 layers (arm_shaping.py:1282-1329).  Defined as ContextSkipNew's graph (separate context encoder with a
    linear hz, skips at every level, no residual) with one conv / transposed conv per entry of `filters`; the default
    is SURVEY.md §8d config 4: 5x5 stride-2 everywhere, 128 -> 2."""
    L = len(filters)
    return NetSpec("ContextSkipDeep", height, width, c_dim, tuple(strides or [2] * L), tuple(kernels or [5] * L),
                   tuple(filters), featsize, separate_ctx_encoder=True, ctx_hz_lrelu=False, residual_out=False)
