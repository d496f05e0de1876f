"""Which part of the training step is not capturable for a given net?  python tools/graph_debug.py"""
import os, sys, traceback
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imitation_from_observation_b200 import _lib, spec as S
from imitation_from_observation_b200.network import FWD_ALL, ContextNet

for (name, spec, B, prec) in [("tiny fp32", S.context_skip_new(16, 16, 3, df_dim=4, featsize=16), 6, "fp32"),
                              ("tiny mixed", S.context_skip_new(16, 16, 3, df_dim=4, featsize=16), 6, "mixed"),
                              ("48x48 mixed", S.context_skip_new(48, 48, 3, df_dim=64, featsize=64), 8, "mixed")]:
    net = ContextNet(spec, B, precision=prec, seed=1)
    net._graph_on = False
    x = torch.rand((3, B, spec.height, spec.width, 3), device="cuda") * 2 - 1
    for _ in range(2):
        net.set_input(x); net.train_step(1e-3)
    torch.cuda.synchronize()
    net._step_dev = torch.zeros(1, device="cuda", dtype=torch.int64); net._lr_t_dev = torch.zeros(1, device="cuda")
    for part in ("zero_grad", "forward", "backward", "apply"):
        g = torch.cuda.CUDAGraph()
        try:
            net._capturing = True
            with torch.cuda.graph(g):
                if part == "zero_grad": net.zero_grad()
                elif part == "forward": net.forward(FWD_ALL)
                elif part == "backward": net.backward()
                else: net.apply_gradients(1e-3)
            print(name, part, "captured OK")
        except Exception as ex:
            print(name, part, "FAILED:", str(ex)[:150].replace("\n", " "), "| lib:", _lib.last_error()[:200])
        finally:
            net._capturing = False
        torch.cuda.synchronize()
