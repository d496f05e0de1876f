"""CPU tests of the oracle itself: it must agree with an independent direct-loop restatement, satisfy the
adjoint/finite-difference properties TF's ops satisfy, reproduce the reference's shape pins, and reproduce the
committed (self-derived) golden vectors."""
import glob
import math
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, rel_err
from imitation_from_observation_b200 import spec as S, synthetic
from oracle import gen_golden, np_direct as D, tf_semantics as O

GEOMS = [(8, 8, 5, 2, 3, 4), (9, 7, 5, 2, 2, 3), (6, 10, 5, 1, 2, 3), (8, 8, 3, 2, 3, 2), (7, 7, 3, 1, 2, 2),
         (3, 3, 5, 2, 2, 2), (8, 8, 1, 2, 2, 2), (36, 64, 5, 2, 1, 1)]


@pytest.mark.parametrize("H,W,k,st,Ci,Co", GEOMS)
def test_conv_ops_match_direct_loops(H, W, k, st, Ci, Co):
    g = torch.Generator().manual_seed(H * 100 + W)
    x = torch.randn(2, H, W, Ci, generator=g, dtype=torch.float64)
    w = torch.randn(k, k, Ci, Co, generator=g, dtype=torch.float64)
    b = torch.randn(Co, generator=g, dtype=torch.float64)
    y = O.conv2d(x, w, b, st)
    assert y.shape[1:3] == (math.ceil(H / st), math.ceil(W / st))
    assert np.abs(y.numpy() - D.conv2d_same(x.numpy(), w.numpy(), b.numpy(), st)).max() < 1e-12
    P, Q = y.shape[1:3]
    xs = torch.randn(2, P, Q, Co, generator=g, dtype=torch.float64)
    t = O.conv2d_transpose(xs, w, None, (H, W), st)          # w read as [kh,kw,out=Ci,in=Co]
    assert np.abs(t.numpy() - D.conv2d_transpose_same(xs.numpy(), w.numpy(), None, (H, W), st)).max() < 1e-12
    # conv2d_transpose is the adjoint of conv2d (TF defines it as conv2d's input gradient)
    lhs = (O.conv2d(x, w, None, st) * xs).sum()
    rhs = (x * t).sum()
    assert abs(float(lhs - rhs)) < 1e-9 * max(1.0, abs(float(lhs)))
    wr = w.clone().requires_grad_(True)
    (O.conv2d(x, wr, None, st) * xs).sum().backward()
    assert np.abs(wr.grad.numpy() - D.conv2d_wgrad(x.numpy(), xs.numpy(), k, k, st)).max() < 1e-11


def test_same_padding_is_asymmetric():
    # 5x5 s2 -> (1,2); 5x5 s1 -> (2,2); 3x3 s1 -> (1,1); 3x3 s2 on even size -> (0,1)
    assert S.same_pad(64, 5, 2) == (1, 2)
    assert S.same_pad(64, 5, 1) == (2, 2)
    assert S.same_pad(8, 3, 1) == (1, 1)
    assert S.same_pad(8, 3, 2) == (0, 1)
    for size in range(1, 40):
        for k in (1, 3, 5):
            for st in (1, 2):
                out, lo = D.same_geometry(size, k, st)
                assert (out, lo) == (S.same_out(size, st), S.same_pad(size, k, st)[0])


def test_shape_pins_from_reference_notebooks():
    """Notebook printouts: push-sim (B,3,3,512)/(B,1024); strike (B,4,4,512);
    push-real-to-sim (B,9,16,8), (B,36,64,32), (B,36,64,3); parameter counts of SURVEY.md §8e."""
    s = S.context_skip_new(48, 48)
    assert s.enc_sizes()[3] == (3, 3) and s.filters[3] == 512 and s.num_params() == 36634179
    s = S.context_skip_new(64, 64)
    assert s.enc_sizes()[3] == (4, 4) and s.num_params() == 47647811
    assert abs(s.fwd_macs_per_triple() / 1e6 - 1183.6) < 0.1
    assert S.context_skip_new(128, 128).num_params() == 123169859
    r = S.context_ae_real()
    assert r.enc_sizes() == [(36, 64), (18, 32), (18, 32), (9, 16)] and r.filters == (32, 16, 16, 8)
    P = O.init_params(r)
    img = torch.zeros(3, 2, 36, 64, 3)
    out = O.forward(r, P, img)
    assert out["out"].shape == (2, 36, 64, 3) and out["ctx_h"][3].shape == (2, 9, 16, 8)
    assert out["ctx_h"][0].shape == (2, 36, 64, 32) and out["input_z"].shape == (2, 100)
    i2 = S.context_ae_inception2(8, 8, 2048, [1, 2, 1, 2], [3] * 4, [1024, 1024, 512, 512])
    assert i2.enc_sizes() == [(8, 8), (4, 4), (4, 4), (2, 2)]


def test_forward_structure_small():
    s = S.context_skip_new(16, 16, 3, df_dim=4, featsize=16)
    P = gen_golden.golden_params(s, 3)
    img = torch.rand(3, 2, 16, 16, 3) * 2 - 1
    r = O.forward(s, P, img)
    # the ctx encoder's hz is linear (arm_shaping.py:1288), the src/tgt one goes through lrelu (:1297)
    c = O.encode(s, P, "conv_context", img[1], False)
    assert torch.allclose(c[-1], r["ctx_z"])
    assert float(r["loss"]) == pytest.approx(float(r["recon1"] + r["recon2"] + r["simloss"]))
    assert float(r["recon1"]) == pytest.approx(float(((img[2] - r["out"]) ** 2).sum() / 2), rel=1e-6)
    assert float(r["simloss"]) == pytest.approx(float(((r["translated_z"] - r["target_z"]) ** 2).mean() * 1e3), rel=1e-6)


def test_gradients_finite_difference():
    s = S.context_ae_inception2(4, 4, 2, [1, 2, 1, 2], [3] * 4, [3, 3, 2, 2], featsize=4)
    P = {k: v.double() for k, v in gen_golden.golden_params(s, 5).items()}
    img = (torch.rand(3, 2, 4, 4, 2, dtype=torch.float64) * 2 - 1)
    _, G = O.loss_and_grads(s, P, img)
    rng = np.random.RandomState(0)
    for name in ("conv/h1_conv/w", "deconv/d_h2/w", "translate/trans_h0/Matrix", "conv_context/h0_conv/biases",
                 "deconv/d_h0_lin/Matrix"):
        flat = P[name].view(-1)
        for _ in range(3):
            i = int(rng.randint(flat.numel()))
            old = float(flat[i])
            eps = 1e-5
            flat[i] = old + eps
            lp = float(O.forward(s, P, img)["loss"])
            flat[i] = old - eps
            lm = float(O.forward(s, P, img)["loss"])
            flat[i] = old
            fd = (lp - lm) / (2 * eps)
            assert fd == pytest.approx(float(G[name].reshape(-1)[i]), rel=2e-4, abs=1e-6), name


def test_deep_spec_generalises_the_reference_depth():
    """BASELINE configs[3] (6 layers): same graph, one level per entry of `filters`.  At depth 4
    the generic spec IS ContextSkipNew — same names, shapes, order and flop count."""
    a = S.context_skip_new(64, 64)
    b = S.context_skip_deep(64, 64, 3, filters=(64, 128, 256, 512))
    assert list(a.param_shapes().items()) == list(b.param_shapes().items())
    assert a.fwd_macs_per_triple() == b.fwd_macs_per_triple() and a.dec_layers() == b.dec_layers()
    d = S.context_skip_deep()
    assert d.depth == 6 and d.lin_name == "h6_lin" and d.flat_dim() == 2048
    assert d.enc_sizes() == [(64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (2, 2)]
    sh = d.param_shapes()
    assert sh["deconv/d_h6/w"] == (5, 5, 3, 128) and sh["deconv/d_h1/w"] == (5, 5, 512, 1024)
    assert sh["conv/h6_lin/Matrix"] == (2048, 1024) and sh["conv_context/h5_conv/w"] == (5, 5, 512, 512)
    assert d.num_params() == 81199171
    with pytest.raises(ValueError):
        S.context_skip_deep(filters=(4, 8), kernels=(5,))
    # tiny instance through the oracle: shapes of every level, and the loss identity
    t = S.context_skip_deep(64, 64, 3, filters=(4, 8, 8, 8, 16, 16), featsize=16)
    P = gen_golden.golden_params(t, 2)
    img = torch.rand((3, 2, 64, 64, 3), generator=torch.Generator().manual_seed(0)) * 2 - 1
    r = O.forward(t, P, img)
    assert [tuple(h.shape[1:]) for h in r["ctx_h"]] == [(32, 32, 4), (16, 16, 8), (8, 8, 8), (4, 4, 8), (2, 2, 16), (1, 1, 16)]
    assert r["out"].shape == (2, 64, 64, 3)
    assert torch.allclose(r["loss"], r["recon1"] + r["recon2"] + r["simloss"])



@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))))
def test_oracle_reproduces_golden(path):
    name = os.path.basename(path)[:-4]
    spec, B = gen_golden.CASES[name]
    d = np.load(path)
    P = {k[6:]: torch.from_numpy(d[k]).double() for k in d.files if k.startswith("param:")}
    assert set(P) == set(spec.param_shapes())
    r, G = O.loss_and_grads(spec, P, torch.from_numpy(d["image"]).double())
    for k in d.files:
        if k.startswith("out:"):
            assert rel_err(r[k[4:]].numpy(), d[k]) < 1e-9, k
        elif k.startswith("grad:"):
            assert rel_err(G[k[5:]].numpy(), d[k]) < 1e-9, k
    # fp32 oracle vs fp64 golden: the error floor the CUDA path is judged against
    P32 = {k: v.float() for k, v in P.items()}
    r32, G32 = O.loss_and_grads(spec, P32, torch.from_numpy(d["image"]))
    assert rel_err(r32["out"].numpy(), d["out:out"]) < 1e-4
    assert rel_err(G32["deconv/d_h1/w"].numpy(), d["grad:deconv/d_h1/w"]) < 1e-4


def test_tf_adam_matches_manual_formula():
    P = {"w": torch.tensor([1.0, -2.0, 3.0])}
    G = {"w": torch.tensor([0.5, -0.25, 0.0])}
    M = {"w": torch.zeros(3)}
    V = {"w": torch.zeros(3)}
    O.tf_adam_step(P, G, M, V, 1, 1e-2)
    m = 0.1 * G["w"]
    v = 0.001 * G["w"] ** 2
    lr_t = 1e-2 * math.sqrt(1 - 0.999) / (1 - 0.9)
    assert torch.allclose(P["w"], torch.tensor([1.0, -2.0, 3.0]) - lr_t * m / (v.sqrt() + 1e-8), atol=1e-7)
    assert float(P["w"][2]) == 3.0


def test_nn_err_and_rewards_semantics():
    tgt = torch.zeros(4, 2, 2, 1)
    for a in range(4):
        tgt[a] += a
    out = tgt[[1, 1, 3, 0]] + 0.1                     # nearest targets: 1,1,3,0 ; b % nlen = 0,1,0,1 (nlen=2)
    assert O.nn_err(tgt, out, 2) == abs(1 - 0) + abs(1 - 1) + abs(3 - 0) + abs(0 - 1)
    T, F = 3, 4
    z = np.ones((T, F), np.float32)
    x = np.zeros((T, 2, 2, 3), np.float32)
    means = np.zeros((T, F), np.float32)
    mimgs = np.ones((T, 2, 2, 3), np.float32)
    rew = np.zeros(2 * T, np.float64)
    out_r = O.reward_ours(z, x, means, mimgs, 0.5, rew)
    cost = F + 0.5 * 12
    assert np.allclose(out_r[1::2], [-cost * j * j for j in range(T)]) and np.all(out_r[0::2] == 0)
    feat = np.ones((T, 1, 1, 2), np.float32)
    mu = np.zeros((T, 1, 1, 2), np.float32)
    sd = np.array([[[[0.0, 1.0]]]] * T, np.float32)
    r2 = O.reward_inception(feat, mu, sd, np.zeros(2 * T))
    assert np.allclose(r2[1::2], [-(0.5 / (1 + 1e-5)) * j * j for j in range(T)])


def test_normalize_u8_three_roundings():
    x = torch.arange(256, dtype=torch.uint8)
    y = O.normalize_u8(x)
    assert y.dtype == torch.float32 and float(y[0]) == -1.0 and float(y[255]) == 1.0
    ref = ((x.numpy().astype(np.float32) * np.float32(1 / 255.0)) - np.float32(0.5)) * np.float32(2.0)
    assert np.array_equal(y.numpy(), ref)


def test_trainer_batch_assembly_matches_reference_indexing():
    bank = synthetic.video_bank(nlen=5, nvideos=6, height=8, width=8)
    assert bank.shape == (5, 6, 8, 8, 3) and bank.dtype == np.uint8 and bank.max() > 0
    assert np.array_equal(bank, synthetic.video_bank(nlen=5, nvideos=6, height=8, width=8))
    a = synthetic.trainer_batch(bank, 7, 5, 4, np.random.RandomState(3))
    b = O.trainer_batch(bank, 7, 5, np.random.RandomState(3), 4)
    assert np.array_equal(a, b) and a.shape == (3, 7, 8, 8, 3)


# ---- InceptionV3 (nets/inception_v3.py): what the reference's own tests pin --------------------------------------------
def _inception_params(seed=0, dtype=torch.float32):
    from imitation_from_observation_b200 import inception as I          # static table only
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name, shp in I.param_shapes().items():
        if name.endswith("/weights"):
            P[name] = (torch.randn(shp, generator=g) * (2.0 / (shp[0] * shp[1] * shp[2])) ** 0.5).to(dtype)
        elif name.endswith("moving_variance"):
            P[name] = (torch.rand(shp, generator=g) * 0.5 + 0.75).to(dtype)
        else:
            P[name] = (torch.randn(shp, generator=g) * 0.1).to(dtype)
    return P


def test_inception_v3_shape_and_parameter_pins():
    """nets/inception_v3_test.py:80-110 (every end-point shape at 299x299), :112-120 (21 802 784 parameters of the
    base network), :141-144 (PreLogits [B,1,1,2048]) — against BOTH the product's table and the oracle's forward."""
    from imitation_from_observation_b200 import inception as I
    from oracle import inception_v3 as OI
    table = I.conv_table()
    assert len(table) == 94
    assert sum(kh * kw * ci * co + 3 * co for _, kh, kw, ci, co in table) == 21802784
    pins = {"Conv2d_1a_3x3": (149, 149, 32), "Conv2d_2a_3x3": (147, 147, 32), "Conv2d_2b_3x3": (147, 147, 64),
            "MaxPool_3a_3x3": (73, 73, 64), "Conv2d_3b_1x1": (73, 73, 80), "Conv2d_4a_3x3": (71, 71, 192),
            "MaxPool_5a_3x3": (35, 35, 192), "Mixed_5b": (35, 35, 256), "Mixed_5c": (35, 35, 288),
            "Mixed_5d": (35, 35, 288), "Mixed_6a": (17, 17, 768), "Mixed_6b": (17, 17, 768), "Mixed_6c": (17, 17, 768),
            "Mixed_6d": (17, 17, 768), "Mixed_6e": (17, 17, 768), "Mixed_7a": (8, 8, 1280), "Mixed_7b": (8, 8, 2048),
            "Mixed_7c": (8, 8, 2048), "PreLogits": (1, 1, 2048)}
    P = _inception_params()
    x = torch.rand((1, 299, 299, 3), generator=torch.Generator().manual_seed(1)) * 2 - 1
    ep = OI.forward(P, x)
    assert ep.pop("_convs_used") == 94                              # the oracle touched every conv of the table
    assert set(ep) == set(pins) == set(I.ENDPOINTS)
    for k, shp in pins.items():
        assert tuple(ep[k].shape) == (1,) + shp, k
    assert float(ep["Mixed_7c"].abs().max()) > 0 and torch.isfinite(ep["PreLogits"]).all()


def test_inception_reward_formula():
    from oracle import inception_v3 as OI
    T, n = 5, 12
    feat, means = torch.rand(T, n), torch.rand(T, n)
    std = torch.rand(T, n)
    std[:, ::3] = 0
    got = OI.reward_inception(feat, means, std)
    ref = []
    for j in range(T):                                               # base.py:184-189 written out
        d = (means[j] - feat[j]).clone()
        d[std[j] == 0] = 0
        ref.append(float(np.mean((d.numpy() ** 2) / (std[j].numpy() + 1e-5))))
    assert np.allclose(got.numpy(), ref, rtol=1e-6)
