import sys, torch
sys.path.insert(0, "/root/repo")
from imitation_from_observation_b200 import ops
from imitation_from_observation_b200._lib import IMPL_AUTO
x = torch.randn(2, 32, 32, 64, device="cuda"); w = torch.randn(5, 5, 64, 128, device="cuda") * 0.05; b = torch.zeros(128, device="cuda")
try:
    y = ops.conv2d(x, w, b, stride=2, act=True, impl=IMPL_AUTO); torch.cuda.synchronize(); print("ok", float(y.abs().mean()))
except Exception as e:
    print("ERR", e)
