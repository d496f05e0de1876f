"""ORACLE — TEST INFRASTRUCTURE ONLY.  **PARITY UNPINNED.**

CPU restatement (PyTorch, fp32 or fp64) of the reference's context-translation hot path.  The reference's
arithmetic lives in TensorFlow (not vendored, not installable here), the reference ships no golden vectors,
weights or tests for this path, so this file is a *restatement* of the algorithm, anchored on
the reference call sites cited per function; it is cross-checked against an independent direct-loop numpy
restatement (oracle/np_direct.py) and by adjointness / finite-difference properties (tests/test_oracle.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.
The shipped product (imitation_from_observation_b200/) never does.

TF semantics restated here that are not visible in the reference source (SURVEY.md §8c, row "exact semantics"):
  1. SAME padding is asymmetric: out=ceil(in/s); total=max((out-1)s+k-in,0); lo=total//2; hi=total-lo.
  2. conv2d_transpose(SAME) is the exact adjoint of the SAME conv with an HWIO filter [kh,kw,Cout_t,Cin_t]:
     full transposed conv of size (in-1)s+k cropped [lo : lo+out].  No kernel flip in either op.
  3. NHWC memory; reshape(h3,[B,-1]) flattens (h,w,c).
  4. tf.nn.l2_loss = sum(t**2)/2.   5. lrelu = max(x, 0.2x); d/dx at x==0 is 1 (tf.maximum passes the tie to x).
  8. TF-Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps).
 10. uint8 frames: x*(1/255) -> -0.5 -> *2 in fp32 (three roundings).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------------------
# primitive ops (gym/envs/mujoco/arm_shaping.py:18-85)
# ------------------------------------------------------------------------------------------------------------
def _same(size: int, k: int, s: int) -> Tuple[int, int, int]:
    out = int(math.ceil(size / s))
    total = max((out - 1) * s + k - size, 0)
    lo = total // 2
    return out, lo, total - lo


class _LRelu(torch.autograd.Function):
    """arm_shaping.py:18-19 ``tf.maximum(x, leak*x)``; gradient 1 for x>=0 (TF routes ties to the first arg)."""

    @staticmethod
    def forward(ctx, x, leak):
        ctx.save_for_backward(x)
        ctx.leak = leak
        return torch.where(x >= 0, x, x * leak)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return torch.where(x >= 0, g, g * ctx.leak), None


def lrelu(x: torch.Tensor, leak: float = 0.2) -> torch.Tensor:
    return _LRelu.apply(x, leak)


def conv2d(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], stride: int) -> torch.Tensor:
    """arm_shaping.py:21-32.  x NHWC, w HWIO [kh,kw,Cin,Cout], SAME, stride (d_h=d_w), + bias."""
    kh, kw = int(w.shape[0]), int(w.shape[1])
    _, lo_h, hi_h = _same(int(x.shape[1]), kh, stride)
    _, lo_w, hi_w = _same(int(x.shape[2]), kw, stride)
    xn = F.pad(x.permute(0, 3, 1, 2), (lo_w, hi_w, lo_h, hi_h))
    y = F.conv2d(xn, w.permute(3, 2, 0, 1), None, stride=stride)
    y = y.permute(0, 2, 3, 1)
    if b is not None:
        y = y + b
    return y.contiguous()


def conv2d_transpose(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor],
                     out_hw: Tuple[int, int], stride: int) -> torch.Tensor:
    """arm_shaping.py:62-85 ``deconv2d``.  x NHWC [N,P,Q,Cin], w [kh,kw,Cout,Cin], SAME, output_shape given."""
    kh, kw = int(w.shape[0]), int(w.shape[1])
    H, W = out_hw
    oh, lo_h, _ = _same(H, kh, stride)
    ow, lo_w, _ = _same(W, kw, stride)
    assert (oh, ow) == (int(x.shape[1]), int(x.shape[2])), "output_shape inconsistent with SAME/stride"
    # F.conv_transpose2d weight layout: [Cin, Cout, kh, kw]
    full = F.conv_transpose2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), None, stride=stride)
    # full extent (in-1)*s+k may be smaller than lo+out when the SAME pad was clamped to zero: pad with zeros
    need_h, need_w = lo_h + H, lo_w + W
    if full.shape[2] < need_h or full.shape[3] < need_w:
        full = F.pad(full, (0, max(need_w - full.shape[3], 0), 0, max(need_h - full.shape[2], 0)))
    y = full[:, :, lo_h:lo_h + H, lo_w:lo_w + W].permute(0, 2, 3, 1)
    if b is not None:
        y = y + b
    return y.contiguous()


def linear(x: torch.Tensor, matrix: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """arm_shaping.py:48-59.  x @ Matrix + bias."""
    y = x @ matrix
    return y + bias if bias is not None else y


def normalize_u8(frames_u8: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """rllab/sampler/base.py:116-119: convert_image_dtype (x * 1/255), subtract 0.5, multiply 2.0 — in fp32."""
    x = frames_u8.to(torch.float32) * torch.tensor(1.0 / 255.0, dtype=torch.float32)
    x = x - 0.5
    x = x * 2.0
    return x.to(dtype)


# ------------------------------------------------------------------------------------------------------------
# parameters
# ------------------------------------------------------------------------------------------------------------
def init_params(spec, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Initialisers of arm_shaping.py:25-29,52-55,67-79: conv ``truncated_normal(0.02)``, deconv / linear
    ``random_normal(0.02)``, biases 0.  RNG streams cannot match TF's; parity tests load identical tensors."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in spec.param_shapes().items():
        if name.endswith("biases") or name.endswith("bias"):
            t = torch.zeros(shape, dtype=torch.float32)
        else:
            t = torch.randn(shape, generator=g, dtype=torch.float32) * 0.02
            if "_conv/w" in name:            # truncated normal: resample beyond 2 sigma
                for _ in range(8):
                    bad = t.abs() > 0.04
                    if not bad.any():
                        break
                    t = torch.where(bad, torch.randn(shape, generator=g, dtype=torch.float32) * 0.02, t)
                t = t.clamp(-0.04, 0.04)
        out[name] = t.to(dtype)
    return out


# ------------------------------------------------------------------------------------------------------------
# the graph (arm_shaping.py:1272-1354 / 1611-1684 / 1792-1894)
# ------------------------------------------------------------------------------------------------------------
def encode(spec, P: Dict[str, torch.Tensor], scope: str, img: torch.Tensor, hz_lrelu: bool):
    """One encoder pass.  Returns (h0,..,h_{L-1}, h_L, z); L = len(spec.filters) = 4 in every reference model, where
    this is the reference's (h0,h1,h2,h3,h4,z).  The dense layer after the last conv is named h<L>_lin."""
    hs = []
    h = img
    nl = len(spec.filters)
    lin = "%s/h%d_lin" % (scope, nl)
    for j in range(nl):
        h = lrelu(conv2d(h, P["%s/h%d_conv/w" % (scope, j)], P["%s/h%d_conv/biases" % (scope, j)],
                         spec.strides[j]))
        hs.append(h)
    flat = hs[-1].reshape(hs[-1].shape[0], -1)                     # (h,w,c) order
    h4 = lrelu(linear(flat, P[lin + "/Matrix"], P[lin + "/bias"]))
    z = linear(h4, P[scope + "/hz_lin/Matrix"], P[scope + "/hz_lin/bias"])
    if hz_lrelu:
        z = lrelu(z)
    return tuple(hs) + (h4, z)


def translate(P, src_z: torch.Tensor, ctx_z: torch.Tensor) -> torch.Tensor:
    """arm_shaping.py:1309-1312."""
    t0 = lrelu(linear(torch.cat([src_z, ctx_z], 1), P["translate/trans_h0/Matrix"], P["translate/trans_h0/bias"]))
    return linear(t0, P["translate/trans_z/Matrix"], P["translate/trans_z/bias"])


def decode(spec, P, z: torch.Tensor, skips: Sequence[torch.Tensor]) -> torch.Tensor:
    """arm_shaping.py:1314-1329.  ``skips`` = ctx encoder (h0,..,h_{L-1})."""
    nl = len(spec.filters)
    h3h, h3w = spec.enc_sizes()[-1]
    s = lrelu(linear(z, P["deconv/d_h0_lin/Matrix"], P["deconv/d_h0_lin/bias"]))
    s = s.reshape(-1, h3h, h3w, spec.filters[-1])
    layers = spec.dec_layers()
    for i, L in enumerate(layers):
        skip = skips[nl - 1 - i]
        s = conv2d_transpose(torch.cat([s, skip], 3), P["deconv/%s/w" % L["name"]],
                             P["deconv/%s/biases" % L["name"]], L["out_hw"], L["stride"])
        if i < nl - 1:
            s = lrelu(s)
    return s


def forward(spec, P: Dict[str, torch.Tensor], image: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Whole graph.  image [3,B,H,W,C]: 0 = source frame, 1 = target *context* frame, 2 = target frame."""
    src, ctx, tgt = image[0], image[1], image[2]
    ctx_scope = "conv_context" if spec.separate_ctx_encoder else "conv"
    c = encode(spec, P, ctx_scope, ctx, spec.ctx_hz_lrelu)
    s = encode(spec, P, "conv", src, True)
    t = encode(spec, P, "conv", tgt, True)
    nl = len(spec.filters)
    trans_z = translate(P, s[-1], c[-1])
    out = decode(spec, P, trans_z, c[:nl])
    out2 = decode(spec, P, t[-1], c[:nl])
    if spec.residual_out:
        out = out + ctx
        out2 = out2 + ctx
    simloss = ((trans_z - t[-1]) ** 2).mean() * 1e3                # arm_shaping.py:1345
    recon1 = ((tgt - out) ** 2).sum() / 2                          # tf.nn.l2_loss, :1352
    recon2 = ((tgt - out2) ** 2).sum() / 2                         # :1353
    return dict(input_z=s[-1], translated_z=trans_z, target_z=t[-1], ctx_z=c[-1], out=out, out2=out2,
                simloss=simloss, recon1=recon1, recon2=recon2, loss=recon1 + recon2 + simloss,
                ctx_h=c[:nl], src_h=s[:nl])


def ablation_loss(r: Dict[str, torch.Tensor], ablation_type: str = "None") -> torch.Tensor:
    """ablations_code/ablations.py:175-182."""
    if ablation_type == "None":
        return r["recon1"] + r["recon2"] + r["simloss"]
    if ablation_type == "L2":
        return r["recon1"] + r["recon2"]
    if ablation_type == "L2L3":
        return r["recon1"]
    if ablation_type == "L1":
        return r["recon2"] + r["simloss"]
    raise ValueError(ablation_type)


def loss_and_grads(spec, P: Dict[str, torch.Tensor], image: torch.Tensor, ablation_type: str = "None"):
    """Forward + autodiff, the CPU equivalent of ``AdamOptimizer.minimize``'s gradient part."""
    Pg = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    r = forward(spec, Pg, image)
    loss = ablation_loss(r, ablation_type)
    grads = torch.autograd.grad(loss, list(Pg.values()), allow_unused=True)
    G = {k: (g if g is not None else torch.zeros_like(Pg[k])) for k, g in zip(Pg.keys(), grads)}
    return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in r.items()}, G


# ------------------------------------------------------------------------------------------------------------
# optimiser, metrics, rewards
# ------------------------------------------------------------------------------------------------------------
def tf_adam_step(P, G, M, V, t: int, lr: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
    """tf.train.AdamOptimizer (scripts/train_script.py:128): epsilon is *outside* the bias correction.
    Restated term for term from TF's published ApplyAdam functor (tensorflow/core/kernels/training_ops.cc, TF 1.x;
    TensorFlow is a third-party dependency absent from /root/reference):
        alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
        m += (g - m) * (1 - beta1);  v += (g*g - v) * (1 - beta2);  var -= (m * alpha) / (sqrt(v) + epsilon)
    with every scalar held in the tensors' dtype (so for fp32, 1 - beta2 is the fp32 value 0.00100004673)."""
    lr_t = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    for k in P:
        dt = P[k].dtype
        one = torch.ones((), dtype=dt)
        ob1 = one - torch.tensor(b1, dtype=dt)
        ob2 = one - torch.tensor(b2, dtype=dt)
        alpha = torch.tensor(lr_t, dtype=dt)
        M[k].add_((G[k] - M[k]) * ob1)
        V[k].add_((G[k] * G[k] - V[k]) * ob2)
        P[k].sub_((M[k] * alpha) / (V[k].sqrt() + torch.tensor(eps, dtype=dt)))


def nn_err(tgt: torch.Tensor, out: torch.Tensor, nlen: int) -> int:
    """scripts/train_script.py:148:  sum_b | argmin_a mean_hwc (tgt[a]-out[b])^2  -  (b mod nlen) |  (integer)."""
    B = tgt.shape[0]
    idx = torch.empty(B, dtype=torch.int64)
    for b in range(B):
        d = ((tgt - out[b:b + 1]) ** 2).reshape(B, -1).mean(1)
        idx[b] = int(torch.argmin(d))                               # first minimum, like tf.argmin
    return int((idx - (torch.arange(B) % nlen)).abs().sum())


def reward_ours(z: np.ndarray, x: np.ndarray, means: np.ndarray, mean_imgs: np.ndarray, scale: float,
                rewards: np.ndarray) -> np.ndarray:
    """rllab/sampler/base.py:243-257 for one path and one viewpoint: features/images are float32,
    ``path['rewards']`` float64; frame j penalises reward index 2j+1 by cost_j * j^2."""
    costs = np.sum((means - z) ** 2, axis=1) + scale * np.sum((mean_imgs - x) ** 2, axis=(1, 2, 3))
    rewards = rewards.copy()
    for j in range(z.shape[0]):
        rewards[j * 2 + 1] -= float(costs[j]) * (j ** 2)
    return rewards


def reward_inception(feat: np.ndarray, means: np.ndarray, std: np.ndarray, rewards: np.ndarray) -> np.ndarray:
    """rllab/sampler/base.py:184-189: variance-normalised squared distance on an Inception endpoint."""
    diff = means - feat
    diff[std == 0] = 0
    diff = diff ** 2 / (std + 1e-5)
    m = np.mean(diff, axis=(1, 2, 3))
    rewards = rewards.copy()
    for j in range(feat.shape[0]):
        rewards[j * 2 + 1] -= float(m[j]) * (j ** 2)
    return rewards


def trainer_batch(bank: np.ndarray, batch_size: int, nlen: int, rng: np.random.RandomState, ntrain: int):
    """scripts/train_script.py:153-159: bank [nlen,N,H,W,C]; returns [3,B,H,W,C] = [src, ctx, tgt]."""
    choicesrc = rng.choice(ntrain, batch_size)
    choicetgt = rng.choice(ntrain, batch_size)
    t = np.arange(0, batch_size) % nlen
    return np.stack([bank[t, choicesrc], bank[0, choicetgt], bank[t, choicetgt]])
