"""Prints forward / gradient deviations of the 6-layer 128x128 model (BASELINE configs[3]) against the fp64 oracle.
    python tools/deep_check.py [precision] [B]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import rel_err, rel_l2                                   # noqa: E402
from imitation_from_observation_b200 import spec as S                  # noqa: E402
from imitation_from_observation_b200.network import FWD_ALL, ContextNet  # noqa: E402
from oracle import tf_semantics as O                                   # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "mixed"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
spec = S.context_skip_deep(128, 128)
P = O.init_params(spec, seed=7)
image = torch.rand((3, B, 128, 128, 3), generator=torch.Generator().manual_seed(8)) * 2 - 1
ref, G = O.loss_and_grads(spec, {k: v.double() for k, v in P.items()}, image.double())
net = ContextNet(spec, B, precision=prec)
net.load_state_dict(P)
net.set_input(image.cuda())
net.zero_grad()
net.forward(FWD_ALL)
net.backward()
torch.cuda.synchronize()
for k in ("input_z", "translated_z", "target_z", "ctx_z", "out", "out2", "recon1", "recon2", "simloss", "loss"):
    print("%-14s rel_err %.3e   |ref|max %.3e" % (k, rel_err(net.fetch(k).cpu(), ref[k]), float(ref[k].abs().max())))
for k in spec.param_shapes():
    print("%-34s rel_l2 %.3e" % (k, rel_l2(net.G[k].cpu(), G[k])))
