"""Runs a few training steps of the bench workload (ContextSkipNew 64x64x3, B triples) — a short target for ncu.
    python tools/one_step.py [--steps 2] [--batch 256] [--precision tf32x3]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imitation_from_observation_b200 import spec as S, synthetic          # noqa: E402
from imitation_from_observation_b200.network import ContextNet            # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--precision", default="tf32x3")
a = ap.parse_args()
spec = S.context_skip_new(64, 64)
net = ContextNet(spec, a.batch, precision=a.precision)
bank = synthetic.to_float(synthetic.video_bank(25, 32, 64, 64)).astype(np.float32)
img = torch.from_numpy(synthetic.trainer_batch(bank, a.batch, 25, 32, np.random.RandomState(0))).cuda()
for _ in range(a.steps):
    net.set_input(img)
    net.train_step(1e-4)
torch.cuda.synchronize()
print("loss", float(net.fetch("loss")))
