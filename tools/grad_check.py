"""Per-parameter gradient error of each precision mode against the fp64 oracle (ContextSkipNew 64x64, B triples):
relative L2 error, and the least-squares scale <g, ref>/<ref, ref> - 1 (a systematic shrink shows up here).
    python tools/grad_check.py [--batch 32] [--modes fp32,tf32x3,mixed,tf32]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imitation_from_observation_b200 import spec as S, synthetic          # noqa: E402
from imitation_from_observation_b200.network import FWD_ALL, ContextNet   # noqa: E402
from oracle import tf_semantics as O                                      # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--modes", default="fp32,tf32x3,mixed,tf32")
a = ap.parse_args()
spec, B = S.context_skip_new(64, 64), a.batch
P = O.init_params(spec, seed=0)
gen = torch.Generator().manual_seed(1)
for k in P:
    if k.endswith("bias") or k.endswith("biases"):
        P[k] = torch.randn(P[k].shape, generator=gen) * 0.01
bank = synthetic.video_bank(25, 16, 64, 64)
image = torch.from_numpy(synthetic.trainer_batch(synthetic.to_float(bank), B, 25, 16, np.random.RandomState(0))).float()
ref, G = O.loss_and_grads(spec, {k: v.double() for k, v in P.items()}, image.double())
_, G32 = O.loss_and_grads(spec, P, image)
res = {}
for mode in a.modes.split(","):
    net = ContextNet(spec, B, precision=mode)
    net.load_state_dict(P)
    net.set_input(image.cuda())
    net.zero_grad()
    net.forward(FWD_ALL)
    net.backward()
    torch.cuda.synchronize()
    res[mode] = ({k: net.G[k].detach().cpu().double() for k in G},
                 {k: net.fetch(k).detach().cpu().double() for k in ("out", "out2", "translated_z", "input_z", "loss")})
res["cpu32"] = ({k: G32[k].double() for k in G}, None)
print("forward max-norm rel err:")
for mode, (_, outs) in res.items():
    if outs:
        print("  %-7s " % mode + " ".join("%s %.2e" % (k, float((outs[k] - ref[k].double()).abs().max() /
                                                                 ref[k].double().abs().max())) for k in outs))
print("%-34s" % "gradient rel-L2 (scale-1)" + "".join("%22s" % m for m in res))
for k in G:
    line = "%-34s" % k
    r = G[k].double().flatten()
    for mode, (g, _) in res.items():
        v = g[k].flatten()
        l2 = float((v - r).norm() / r.norm())
        sc = float((v @ r) / (r @ r)) - 1.0
        line += "   %.2e (%+.1e)" % (l2, sc)
    print(line)
