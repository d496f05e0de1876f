"""ORACLE — TEST INFRASTRUCTURE ONLY.  **PARITY UNPINNED** (numerically): CPU restatement of the InceptionV3 forward pass.

Follows nets/inception_v3.py:29-416 (`inception_v3_base`), :484-496 (`PreLogits`) and the arg scope of
nets/inception_utils.py:31-71 in PyTorch (fp32 or fp64).  The reference ships no checkpoint and no numeric golden for
this network; what its own tests pin — every end-point shape for a 299x299 input and the parameter count 21 802 784
(nets/inception_v3_test.py:80-120, 141-144) — is checked in tests/test_oracle.py.  Written block by block,
independently of the table in imitation_from_observation_b200/inception.py, so that a transcription error in either
shows up as a mismatch.  Variable names are slim's: <scope>/weights [kh,kw,Cin,Cout], <scope>/BatchNorm/{beta,
moving_mean, moving_variance} (no gamma: slim.batch_norm's default scale=False).

TF semantics restated: SAME padding is asymmetric (low half first); VALID has none; avg_pool2d(SAME) divides by the
number of in-range elements; batch norm at inference is (x - mean) / sqrt(var + 0.001) + beta; activation ReLU.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

EPS = 1e-3


def _pad_same(x, kh, kw, stride):
    H, W = x.shape[2], x.shape[3]
    oh, ow = math.ceil(H / stride), math.ceil(W / stride)
    th, tw = max((oh - 1) * stride + kh - H, 0), max((ow - 1) * stride + kw - W, 0)
    return F.pad(x, (tw // 2, tw - tw // 2, th // 2, th - th // 2))


class _Net:
    def __init__(self, P: Dict[str, torch.Tensor]):
        self.P = P
        self.used = set()

    def conv(self, x, scope, stride=1, padding="SAME"):
        """slim.conv2d: bias-free conv -> inference BN without gamma -> ReLU.  x is NCHW here."""
        w = self.P[scope + "/weights"]                              # [kh, kw, Cin, Cout]
        kh, kw = w.shape[0], w.shape[1]
        if padding == "SAME":
            x = _pad_same(x, kh, kw, stride)
        y = F.conv2d(x, w.permute(3, 2, 0, 1), stride=stride)
        mean, var, beta = (self.P[scope + "/BatchNorm/" + n] for n in ("moving_mean", "moving_variance", "beta"))
        y = (y - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + EPS) + beta.view(1, -1, 1, 1)
        self.used.add(scope)
        return torch.relu(y)


def _max_pool(x, k=3, stride=2):
    return F.max_pool2d(x, k, stride)                               # VALID


def _avg_pool_same(x, k=3):
    return F.avg_pool2d(x, k, 1, padding=k // 2, count_include_pad=False)


def forward(P: Dict[str, torch.Tensor], images_nhwc: torch.Tensor) -> Dict[str, torch.Tensor]:
    """images_nhwc [N,H,W,3], already normalised.  Returns every end point of the base network + PreLogits, NHWC."""
    n = _Net(P)
    r = "InceptionV3/"
    ep: Dict[str, torch.Tensor] = {}
    x = images_nhwc.permute(0, 3, 1, 2)
    x = ep["Conv2d_1a_3x3"] = n.conv(x, r + "Conv2d_1a_3x3", 2, "VALID")
    x = ep["Conv2d_2a_3x3"] = n.conv(x, r + "Conv2d_2a_3x3", 1, "VALID")
    x = ep["Conv2d_2b_3x3"] = n.conv(x, r + "Conv2d_2b_3x3", 1, "SAME")
    x = ep["MaxPool_3a_3x3"] = _max_pool(x)
    x = ep["Conv2d_3b_1x1"] = n.conv(x, r + "Conv2d_3b_1x1", 1, "VALID")
    x = ep["Conv2d_4a_3x3"] = n.conv(x, r + "Conv2d_4a_3x3", 1, "VALID")
    x = ep["MaxPool_5a_3x3"] = _max_pool(x)

    # 35 x 35: Mixed_5b, 5c, 5d
    for name, b1a, b1b in (("Mixed_5b", "Conv2d_0a_1x1", "Conv2d_0b_5x5"), ("Mixed_5c", "Conv2d_0b_1x1", "Conv_1_0c_5x5"),
                           ("Mixed_5d", "Conv2d_0a_1x1", "Conv2d_0b_5x5")):
        s = r + name + "/"
        b0 = n.conv(x, s + "Branch_0/Conv2d_0a_1x1")
        b1 = n.conv(n.conv(x, s + "Branch_1/" + b1a), s + "Branch_1/" + b1b)
        b2 = n.conv(n.conv(n.conv(x, s + "Branch_2/Conv2d_0a_1x1"), s + "Branch_2/Conv2d_0b_3x3"), s + "Branch_2/Conv2d_0c_3x3")
        b3 = n.conv(_avg_pool_same(x), s + "Branch_3/Conv2d_0b_1x1")
        x = ep[name] = torch.cat([b0, b1, b2, b3], 1)

    # 35 -> 17: Mixed_6a
    s = r + "Mixed_6a/"
    b0 = n.conv(x, s + "Branch_0/Conv2d_1a_1x1", 2, "VALID")
    b1 = n.conv(n.conv(n.conv(x, s + "Branch_1/Conv2d_0a_1x1"), s + "Branch_1/Conv2d_0b_3x3"), s + "Branch_1/Conv2d_1a_1x1", 2, "VALID")
    x = ep["Mixed_6a"] = torch.cat([b0, b1, _max_pool(x)], 1)

    # 17 x 17: Mixed_6b .. 6e (1x7 / 7x1 factorisation)
    for name in ("Mixed_6b", "Mixed_6c", "Mixed_6d", "Mixed_6e"):
        s = r + name + "/"
        b0 = n.conv(x, s + "Branch_0/Conv2d_0a_1x1")
        b1 = n.conv(x, s + "Branch_1/Conv2d_0a_1x1")
        b1 = n.conv(n.conv(b1, s + "Branch_1/Conv2d_0b_1x7"), s + "Branch_1/Conv2d_0c_7x1")
        b2 = n.conv(x, s + "Branch_2/Conv2d_0a_1x1")
        for sc in ("Conv2d_0b_7x1", "Conv2d_0c_1x7", "Conv2d_0d_7x1", "Conv2d_0e_1x7"):
            b2 = n.conv(b2, s + "Branch_2/" + sc)
        b3 = n.conv(_avg_pool_same(x), s + "Branch_3/Conv2d_0b_1x1")
        x = ep[name] = torch.cat([b0, b1, b2, b3], 1)

    # 17 -> 8: Mixed_7a
    s = r + "Mixed_7a/"
    b0 = n.conv(n.conv(x, s + "Branch_0/Conv2d_0a_1x1"), s + "Branch_0/Conv2d_1a_3x3", 2, "VALID")
    b1 = n.conv(x, s + "Branch_1/Conv2d_0a_1x1")
    b1 = n.conv(n.conv(b1, s + "Branch_1/Conv2d_0b_1x7"), s + "Branch_1/Conv2d_0c_7x1")
    b1 = n.conv(b1, s + "Branch_1/Conv2d_1a_3x3", 2, "VALID")
    x = ep["Mixed_7a"] = torch.cat([b0, b1, _max_pool(x)], 1)

    # 8 x 8: Mixed_7b, 7c (split 1x3 / 3x1 heads)
    for name, b1_3x1 in (("Mixed_7b", "Conv2d_0b_3x1"), ("Mixed_7c", "Conv2d_0c_3x1")):
        s = r + name + "/"
        b0 = n.conv(x, s + "Branch_0/Conv2d_0a_1x1")
        t = n.conv(x, s + "Branch_1/Conv2d_0a_1x1")
        b1 = torch.cat([n.conv(t, s + "Branch_1/Conv2d_0b_1x3"), n.conv(t, s + "Branch_1/" + b1_3x1)], 1)
        t = n.conv(n.conv(x, s + "Branch_2/Conv2d_0a_1x1"), s + "Branch_2/Conv2d_0b_3x3")
        b2 = torch.cat([n.conv(t, s + "Branch_2/Conv2d_0c_1x3"), n.conv(t, s + "Branch_2/Conv2d_0d_3x1")], 1)
        b3 = n.conv(_avg_pool_same(x), s + "Branch_3/Conv2d_0b_1x1")
        x = ep[name] = torch.cat([b0, b1, b2, b3], 1)

    ep["PreLogits"] = F.avg_pool2d(x, (x.shape[2], x.shape[3]))     # whole-map VALID pool; dropout = identity
    out = {k: v.permute(0, 2, 3, 1).contiguous() for k, v in ep.items()}
    out["_convs_used"] = len(n.used)
    return out


def reward_inception(feat: torch.Tensor, means: torch.Tensor, std: torch.Tensor) -> torch.Tensor:
    """rllab/sampler/base.py:184-189 for one path: d = means - feat; d[std == 0] = 0; mean_{hwc} d^2 / (std + 1e-5)."""
    d = means - feat
    d = torch.where(std == 0, torch.zeros_like(d), d)
    return (d * d / (std + 1e-5)).reshape(d.shape[0], -1).mean(1)
