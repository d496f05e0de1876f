"""A handful of launches of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imitation_from_observation_b200 import ops, spec as S                        # noqa: E402
from imitation_from_observation_b200._lib import IMPL_AUTO, IMPL_FAST, IMPL_SIMT  # noqa: E402
from imitation_from_observation_b200.network import FWD_ALL, ContextNet           # noqa: E402

g = torch.Generator().manual_seed(0)
r = lambda *s: (torch.randn(s, generator=g) * 0.1).cuda()
for impl in (IMPL_AUTO, IMPL_FAST, IMPL_SIMT):
    ops.conv2d(r(4, 16, 16, 64), r(5, 5, 64, 128), r(128), stride=2, act=True, impl=impl)          # pair / x1 / simt
    ops.conv2d(r(2, 17, 17, 48), r(3, 3, 48, 80), r(80), stride=1, act=True, impl=impl)            # partial tiles, ragged
    ops.conv2d_transpose(r(4, 8, 8, 64), r(5, 5, 64, 128), (16, 16), bias=r(64), stride=2, x2=r(2, 8, 8, 64), act=True, impl=impl)
    ops.conv2d_transpose(r(2, 16, 16, 32), r(5, 5, 3, 64), (32, 32), bias=r(3), stride=2, x2=r(2, 16, 16, 32), act=False, impl=impl)
    ops.conv2d_wgrad(r(4, 16, 16, 64), r(4, 8, 8, 64), 5, 2, small2=r(2, 8, 8, 64), impl=impl)
    ops.gemm(r(256, 512), r(512, 128), bias=r(128), act=True, impl=impl)
    ops.gemm(r(512, 76)[:, :75], r(512, 64), transA=True, accumulate=True, out=torch.zeros(75, 64).cuda(), impl=impl)
torch.cuda.synchronize()
net = ContextNet(S.context_skip_new(16, 16, 3, df_dim=32, featsize=64), 4, precision="mixed")
net.set_input((torch.rand((3, 4, 16, 16, 3), generator=g) * 2 - 1).cuda())
net.train_step(1e-4)
torch.cuda.synchronize()
print("sanitize smoke done, loss", float(net.fetch("loss")))
