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
# forward convolutions on bf16 planes: single CTA (64 / 128 wide), CTA pair (256 wide), transposed phases, merged phases
xp = ops.split_bf16(r(4, 16, 16, 64))
for Co in (64, 128, 256):
    ops.conv2d_bf16x3(xp, ops.split_bf16_conv_filter(r(5, 5, 64, Co)), 5, bias=r(Co), stride=2, act=True,
                      out_planes=torch.empty((2, 4, 8, 8, Co), device="cuda", dtype=torch.bfloat16))
ops.conv2d_transpose_bf16x3(ops.split_bf16(r(4, 8, 8, 64)), ops.split_bf16(r(5, 5, 128, 128)), (16, 16), 5, bias=r(128), stride=2,
                            x2p=ops.split_bf16(r(2, 8, 8, 64)), act=True)
ops.conv2d_transpose_merged_bf16x3(ops.split_bf16(r(4, 8, 8, 64)), ops.conv2d_transpose_merged_pack(r(5, 5, 64, 128), (16, 16), dtype=torch.bfloat16),
                                   (16, 16), 5, 64, bias=r(64), stride=2, x2p=ops.split_bf16(r(2, 8, 8, 64)), act=True)
ops.gather_frames(r(10, 16, 16, 3), torch.tensor([3, 1, 9], device="cuda"))
torch.cuda.synchronize()
# skinny dense layers: cluster split-K when CTX_TC_CLUSTER_K=1 (partials through distributed shared memory), else grid split-K
for impl in (IMPL_AUTO, IMPL_FAST):
    ops.gemm(r(64, 2048), r(2048, 1024), bias=r(1024), act=True, impl=impl)
    ops.gemm(r(256, 1024), r(1024, 1024), transB=True, impl=impl)
    ops.gemm(r(256, 1024), r(256, 512), transA=True, accumulate=True, out=torch.zeros(1024, 512).cuda(), impl=impl)   # += as vector reductions
torch.cuda.synchronize()
# two eager training steps of full-width nets in "mixed": h16_kernel (three products forward, one product backward, merged
# phases, filter gradients), the few-channel kernels at 64x64 (h0, d_h4 + fused loss and their gradients), grad_planes / amax /
# skip_grad_sum, Adam with plane refresh
for (hw, B) in ((32, 8), (64, 4)):
    net = ContextNet(S.context_skip_new(hw, hw, 3, df_dim=64, featsize=64), B, precision="mixed")
    net._graph_on = False
    assert net.bf3 and net.p1
    for _ in range(2):
        net.set_input((torch.rand((3, B, hw, hw, 3), generator=g) * 2 - 1).cuda())
        net.train_step(1e-4)
    torch.cuda.synchronize()
    print("sanitize smoke %dx%d done, loss" % (hw, hw), float(net.fetch("loss")), "few:", net._enc_few[0], net._dec_fewt[-1])
