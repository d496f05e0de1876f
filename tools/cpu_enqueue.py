"""Host time to ENQUEUE one training step (no synchronisation inside) against the GPU time of the step: tells whether the
launch path (Python + ctypes + tensor-map encoding) or the GPU bounds the step.    python tools/cpu_enqueue.py [B]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imitation_from_observation_b200 import spec as S                     # noqa: E402
from imitation_from_observation_b200.network import ContextNet            # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
net = ContextNet(S.context_skip_new(64, 64), B)
x = torch.rand((3, B, 64, 64, 3), device="cuda") * 2 - 1
net.set_input(x)
for _ in range(5):
    net.train_step(1e-4)
torch.cuda.synchronize()
n = 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
cpu = []
e0.record()
for _ in range(n):
    t0 = time.perf_counter()
    net.train_step(1e-4)
    cpu.append(time.perf_counter() - t0)
e1.record()
torch.cuda.synchronize()
cpu.sort()
print("B=%d: GPU %.2f ms/step; host enqueue median %.2f ms, min %.2f ms, max %.2f ms per step"
      % (B, e0.elapsed_time(e1) / n, cpu[n // 2] * 1e3, cpu[0] * 1e3, cpu[-1] * 1e3))
