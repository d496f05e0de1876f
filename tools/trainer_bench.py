"""Per-iteration wall time of ModelTrainer.train (train_script.py:153-167) at B=256, 64x64: the reference's host gather +
float feed against the HBM-resident bank (data.DeviceBank).  Time per iteration = (t(N2) - t(N1)) / (N2 - N1), which
cancels graph build and bank generation.
    python tools/trainer_bench.py [N1 N2]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imitation_from_observation_b200 import synthetic                   # noqa: E402
from imitation_from_observation_b200.trainer import ModelTrainer        # noqa: E402

N1 = int(sys.argv[1]) if len(sys.argv) > 1 else 21
N2 = int(sys.argv[2]) if len(sys.argv) > 2 else 101
bank = synthetic.to_float(synthetic.video_bank(25, 64, 64, 64, seed=1))            # float64 [25, 64, 64, 64, 3]


def run(nitr, dev):
    tr = ModelTrainer(idims=(64, 64), nvideos=64, ntrain=64, batch_size=256, model="ContextSkipNew", nitr=nitr,
                      save_every=10 ** 9, nlen=25, nskip=2, data=bank, log=None, device_bank=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tr.train()
    torch.cuda.synchronize()
    return time.perf_counter() - t0


run(5, True)                                                                        # warm the context / allocator
for dev in (False, True):
    a, b = run(N1, dev), run(N2, dev)
    ms = (b - a) / (N2 - N1) * 1e3
    print("device_bank=%-5s  %.2f ms per iteration  = %.0f frames/s   (loss + nn_err fetched every 4th step)"
          % (dev, ms, 3 * 256 / ms * 1e3))
