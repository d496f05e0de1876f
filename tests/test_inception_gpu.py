"""GPU parity of the InceptionV3 feature extractor (nets/inception_v3.py) and the Inception reward against the CPU
oracle, through the C-ABI.  Random (variance-scaling) weights: the reference ships no checkpoint."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from imitation_from_observation_b200 import _lib, inception as I, ops
from oracle import inception_v3 as OI
from test_oracle import _inception_params

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,stride,padding,mode", [(3, 2, "VALID", "max"), (3, 1, "SAME", "avg"), (8, 1, "VALID", "avg"),
                                                   (3, 2, "SAME", "max")])
def test_pool2d(k, stride, padding, mode):
    hw = 8 if k == 8 else 17
    x = torch.randn((2, hw, hw, 12), generator=torch.Generator().manual_seed(3))
    xc = x.permute(0, 3, 1, 2)
    Fn = torch.nn.functional
    if padding == "VALID":
        ref = Fn.max_pool2d(xc, k, stride) if mode == "max" else Fn.avg_pool2d(xc, k, stride)
    elif mode == "avg":
        ref = OI._avg_pool_same(xc, k)                               # divides by the number of in-range elements
    else:                                                            # SAME max pooling: padding never wins
        out = -(-hw // stride)
        tot = max((out - 1) * stride + k - hw, 0)
        ref = Fn.max_pool2d(Fn.pad(xc, (tot // 2, tot - tot // 2, tot // 2, tot - tot // 2), value=float("-inf")), k, stride)
    y = ops.pool2d(x.cuda(), k, stride, padding, mode)
    assert tuple(y.shape) == tuple(ref.permute(0, 2, 3, 1).shape)
    assert rel_err(y.cpu(), ref.permute(0, 2, 3, 1)) < 1e-6


def test_copy_channels_is_concat():
    a, b = torch.randn(3, 4, 5, 8), torch.randn(3, 4, 5, 5)
    out = torch.zeros(3, 4, 5, 13).cuda()
    ops.copy_channels(a.cuda(), out, 0)
    ops.copy_channels(b.cuda(), out, 8)
    assert torch.equal(out.cpu(), torch.cat([a, b], 3))


@pytest.mark.parametrize("kh,kw,stride,padding", [(1, 7, 1, "SAME"), (7, 1, 1, "SAME"), (1, 3, 1, "SAME"), (3, 3, 2, "VALID"),
                                                  (3, 3, 1, "VALID"), (1, 1, 1, "SAME"), (5, 5, 1, "SAME")])
@pytest.mark.parametrize("hw,c,co", [(17, 128, 192), (8, 64, 96), (9, 48, 64)])
def test_conv2d_general_rectangular_and_valid(kh, kw, stride, padding, hw, c, co):
    g = torch.Generator().manual_seed(5)
    x = torch.randn((2, hw, hw, c), generator=g)
    w = torch.randn((kh, kw, c, co), generator=g) * (kh * kw * c) ** -0.5
    b = torch.randn((co,), generator=g) * 0.1
    xc = x.double().permute(0, 3, 1, 2)
    if padding == "SAME":
        xc = OI._pad_same(xc, kh, kw, stride)
    ref = torch.relu(torch.nn.functional.conv2d(xc, w.double().permute(3, 2, 0, 1), b.double(), stride=stride))
    for impl, tol in ((_lib.IMPL_SIMT, 2e-5), (_lib.IMPL_AUTO, 2e-5)):
        y = ops.conv2d_general(x.cuda(), w.cuda(), b.cuda(), stride=stride, padding=padding, act=True, slope=0.0, impl=impl)
        assert rel_err(y.cpu(), ref.permute(0, 2, 3, 1)) < tol, impl


@pytest.mark.parametrize("hw,c,co,k,stride,padding", [
    (147, 32, 64, 3, 1, "SAME"),      # wider than one 128-pixel tile: column tiles (q_tiles > 1)
    (149, 32, 32, 3, 1, "VALID"),     # N = 32 tile
    (73, 80, 192, 3, 1, "VALID"),     # ragged channel block (80 = 2.5 x 32), exact 192-wide tile, 71-wide output
    (35, 288, 384, 3, 2, "VALID"),    # stride-2 parity views on an odd grid
    (35, 48, 64, 5, 1, "SAME"),       # 35-wide grid: two column tiles of 18
    (71, 192, 32, 1, 1, "SAME")])
def test_conv2d_general_inception_stem_shapes(hw, c, co, k, stride, padding):
    g = torch.Generator().manual_seed(7)
    x = torch.randn((2, hw, hw, c), generator=g)
    w = torch.randn((k, k, c, co), generator=g) * (k * k * c) ** -0.5
    b = torch.randn((co,), generator=g) * 0.1
    xc = x.double().permute(0, 3, 1, 2)
    if padding == "SAME":
        xc = OI._pad_same(xc, k, k, stride)
    ref = torch.relu(torch.nn.functional.conv2d(xc, w.double().permute(3, 2, 0, 1), b.double(), stride=stride))
    for impl, tol in ((_lib.IMPL_AUTO, 2e-5), (_lib.IMPL_FAST, 3e-3)):
        y = ops.conv2d_general(x.cuda(), w.cuda(), b.cuda(), stride=stride, padding=padding, act=True, slope=0.0, impl=impl)
        assert int(_lib.load().ctx_last_impl()) != _lib.IMPL_SIMT
        assert rel_err(y.cpu(), ref.permute(0, 2, 3, 1)) < tol, impl
    # ... and into a wider destination (the block's concat buffer)
    dst = torch.full((2, ref.shape[2], ref.shape[3], co + 24), 7.0).cuda()
    ops.conv2d_general(x.cuda(), w.cuda(), b.cuda(), stride=stride, padding=padding, act=True, slope=0.0, into=dst, c_off=8)
    assert rel_err(dst[..., 8:8 + co].cpu(), ref.permute(0, 2, 3, 1)) < 2e-5
    assert float((dst[..., :8] - 7.0).abs().max()) == 0.0 and float((dst[..., 8 + co:] - 7.0).abs().max()) == 0.0


def test_inception_v3_endpoints_match_oracle():
    """B = 2, 299x299 uint8 frames (as base.py:72-76 feeds them): Mixed_5d / Mixed_6e / Mixed_7c / PreLogits vs the
    fp64 oracle, 1e-3 relative (north_star tolerance); 94 folded conv launches."""
    P = _inception_params(seed=11)
    frames = torch.randint(0, 256, (2, 299, 299, 3), generator=torch.Generator().manual_seed(12), dtype=torch.uint8)
    x = ((frames.float() * (1.0 / 255.0)) - 0.5) * 2.0
    ref = OI.forward({k: v.double() for k, v in P.items()}, x.double())
    net = I.InceptionV3(precision="mixed")
    net.load_state_dict(P)
    got = net.forward(frames.cuda(), endpoints=("Mixed_5d", "Mixed_6e", "Mixed_7c", "PreLogits"))
    torch.cuda.synchronize()
    for k in ("Mixed_5d", "Mixed_6e", "Mixed_7c", "PreLogits"):
        assert tuple(got[k].shape) == tuple(ref[k].shape), k
        assert rel_err(got[k].cpu(), ref[k]) < 1e-3, k
    # early exit: asking for an early end point does not run the rest
    n0 = ops.launch_count()
    net.forward(frames.cuda(), endpoints=("Mixed_5b",))
    n1 = ops.launch_count()
    net.forward(frames.cuda(), endpoints=("Mixed_7c",))
    assert n1 - n0 < ops.launch_count() - n1


def test_inception_loads_a_tf_checkpoint(tmp_path):
    """slim's published inception_v3.ckpt is a single-file V1 checkpoint (base.py:130-132 restores it with tf.train.Saver;
    a re-saved copy would be a V2 tensor bundle): InceptionV3.load_tf_checkpoint reads the 4 x 94 variables by name from
    either format and ignores the heads the file also holds."""
    from imitation_from_observation_b200 import tf_checkpoint as T
    P = _inception_params(seed=14)
    sd = {k: v.numpy() for k, v in P.items()}
    sd["InceptionV3/Logits/Conv2d_1c_1x1/weights"] = np.zeros((1, 1, 2048, 1001), np.float32)      # unused head
    sd["InceptionV3/Logits/Conv2d_1c_1x1/biases"] = np.zeros((1001,), np.float32)
    a = I.InceptionV3(precision="mixed")
    a.load_state_dict(P)
    for fmt, writer in (("v1", T.write_checkpoint_v1), ("v2", T.write_checkpoint)):
        prefix = str(tmp_path / fmt / "inception_v3.ckpt")
        writer(prefix, sd)
        b = I.InceptionV3(precision="mixed")
        b.load_tf_checkpoint(prefix)
        for scope in a.folded:
            assert torch.equal(a.folded[scope][0], b.folded[scope][0]) and torch.equal(a.folded[scope][1], b.folded[scope][1])


def test_inception_reward_kernel_matches_reference_loop():
    T, paths = 25, 6
    g = torch.Generator().manual_seed(13)
    feat = torch.rand((paths * T, 4, 4, 32), generator=g)
    means, std = torch.rand((T, 4, 4, 32), generator=g), torch.rand((T, 4, 4, 32), generator=g)
    std[:, :, :, ::5] = 0
    rewards = torch.zeros((paths, 2 * T), dtype=torch.float64)
    ref = rewards.clone()
    for p in range(paths):
        c = OI.reward_inception(feat[p * T:(p + 1) * T].reshape(T, -1), means.reshape(T, -1), std.reshape(T, -1))
        for j in range(T):
            ref[p, 2 * j + 1] -= float(c[j]) * j * j
    r = rewards.cuda()
    ops.reward_inception(feat.cuda(), means.cuda(), std.cuda(), paths, T, rewards=r)
    assert rel_err(r.cpu(), ref) < 1e-5


def test_reward_featurizer_inception_mode_matches_reference_loop():
    """rllab/sampler/base.py:179-189 end to end: uint8 frames -> InceptionV3 end point -> variance-normalised distance
    -> rewards[2j+1] -= m_j * j^2, batched over paths, against the per-path reference loop on the oracle's features."""
    from imitation_from_observation_b200.sampler import RewardFeaturizer
    T, paths_n, hw = 25, 3, 75                                      # 75x75 frames keep the CPU oracle quick
    P = _inception_params(seed=21)
    net = I.InceptionV3(precision="mixed")
    net.load_state_dict(P)
    rng = np.random.RandomState(0)
    paths = [dict(rewards=np.zeros(2 * T), env_infos=dict(imgs=[None if k % 2 else [rng.randint(0, 256, (hw, hw, 3)).astype(np.uint8)]
                                                                for k in range(2 * T)])) for _ in range(paths_n)]
    def feats(frames_u8):
        x = ((torch.from_numpy(frames_u8).float() * (1.0 / 255.0)) - 0.5) * 2.0
        return OI.forward(P, x)["Mixed_5d"]
    all_frames = [np.asarray([im[0] for im in p["env_infos"]["imgs"] if im is not None]) for p in paths]
    f0 = feats(all_frames[0])
    means = (f0 * 0.9).numpy()
    std = (f0.abs() * 0.5).numpy()
    std[..., ::7] = 0
    ref = [p["rewards"].copy() for p in paths]
    for i in range(paths_n):
        c = OI.reward_inception(feats(all_frames[i]).reshape(T, -1), torch.from_numpy(means).reshape(T, -1),
                                torch.from_numpy(std).reshape(T, -1))
        for j in range(T):
            ref[i][2 * j + 1] -= float(c[j]) * j * j
    fz = RewardFeaturizer(mode="inception", name="push", imsize=(hw, hw), layer="Mixed_5d", inception=net, chunk_paths=2)
    fz.initialize()
    fz.set_inception_statistics(means, std)
    fz.process_paths(paths)
    for i in range(paths_n):
        assert rel_err(paths[i]["rewards"], ref[i]) < 1e-3


def test_reward_featurizer_oursinception_mode():
    """base.py:121-132 + 232-257: uint8 frames -> frozen InceptionV3 Mixed_7c -> ContextAEInception2 source encoder ->
    feature + feature-map distance.  75x75 frames (Mixed_7c is 1x1x2048 there) against the two oracles chained."""
    from imitation_from_observation_b200 import spec as S
    from imitation_from_observation_b200.sampler import RewardFeaturizer
    from oracle import tf_semantics as O
    T, paths_n, hw, scale = 25, 2, 75, 0.5
    PI = _inception_params(seed=31)
    inc = I.InceptionV3(precision="mixed")
    inc.load_state_dict(PI)
    fz = RewardFeaturizer(mode="oursinception", name="strike", imsize=(hw, hw), nvp=1, scale=scale, inception=inc)
    fz.initialize()
    spec = fz.model.spec
    assert (spec.height, spec.width, spec.c_dim) == (1, 1, 2048) and spec.filters == (1024, 1024, 512, 512)
    PM = O.init_params(spec, seed=32)
    fz.model.net.load_state_dict(PM)
    rng = np.random.RandomState(1)
    paths = [dict(rewards=np.zeros(2 * T), env_infos=dict(imgs=[None if k % 2 else [rng.randint(0, 256, (hw, hw, 3)).astype(np.uint8)]
                                                                for k in range(2 * T)])) for _ in range(paths_n)]
    g = torch.Generator().manual_seed(33)
    means = torch.randn((T, spec.featsize), generator=g) * 0.01
    imgs = torch.rand((T, 1, 1, 2048), generator=g)
    fz.means, fz.imgs = [means.cuda()], [imgs.cuda()]
    ref = []
    for p in paths:
        fr = np.asarray([im[0] for im in p["env_infos"]["imgs"] if im is not None])
        x = ((torch.from_numpy(fr).float() * (1.0 / 255.0)) - 0.5) * 2.0
        feat = OI.forward(PI, x)["Mixed_7c"]                                       # [T,1,1,2048]
        z = O.forward(spec, PM, torch.stack([feat, feat[:1].expand_as(feat), feat]))["input_z"]
        cost = ((means - z) ** 2).sum(1) + scale * ((imgs - feat) ** 2).reshape(T, -1).sum(1)
        r = np.zeros(2 * T)
        for j in range(T):
            r[2 * j + 1] -= float(cost[j]) * j * j
        ref.append(r)
    fz.process_paths(paths)
    for i in range(paths_n):
        assert rel_err(paths[i]["rewards"], ref[i]) < 1e-3


def test_model_trainer_with_inception_front_end():
    """scripts/train_script.py:100-115: uint8 frames -> frozen InceptionV3 -> Mixed_7c -> ContextAEInception2 trained by
    Adam; only `contextmodel/*` variables change, the first fetched loss equals the oracles chained."""
    from imitation_from_observation_b200.trainer import ModelTrainer
    inc = I.InceptionV3(precision="mixed", seed=41)
    before = {k: v.clone() for k, v in list(inc.params.items())[:3]}
    logs = []
    tr = ModelTrainer(idims=(75, 75), nvideos=6, ntrain=4, batch_size=4, model="ContextAEInception2", nitr=9, save_every=1000,
                      nlen=2, nskip=1, rescale=False, inception=True, strides=[1, 2, 1, 2], kernels=[3, 3, 3, 3],
                      filters=[64, 64, 32, 32], precision="mixed", log=logs.append, inception_net=inc)
    allloss, _ = tr.train()
    assert len(allloss) == 2 and all(np.isfinite(l) for l in allloss)
    assert tr.test.spec.c_dim == 2048 and tr.test.spec.filters == (64, 64, 32, 32)
    for k, v in before.items():
        assert torch.equal(inc.params[k], v)                           # the front-end is frozen
