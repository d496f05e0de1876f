"""Reward featurizer throughput (rllab/sampler/base.py:232-257): one TRPO iteration's worth of trajectories —
250 paths x 25 frames, 48x48 (push), nvp = 1 — through RewardFeaturizer.process_paths (host uint8 frames in, float64
rewards out), against the reference's per-path loop restated on the CPU oracle (a bounded sample of paths).
    python tools/reward_bench.py [--paths 250] [--hw 48] [--cpu-paths 10]"""
import argparse, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imitation_from_observation_b200.sampler import RewardFeaturizer   # noqa: E402
from oracle import tf_semantics as O                                    # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--paths", type=int, default=250)
ap.add_argument("--hw", type=int, default=48)
ap.add_argument("--cpu-paths", type=int, default=10)
a = ap.parse_args()
T, hw = 25, a.hw
rng = np.random.RandomState(0)


def make_paths(n):
    return [dict(rewards=np.zeros(2 * T), env_infos=dict(imgs=[None if k % 2 else [rng.randint(0, 256, (hw, hw, 3)).astype(np.uint8)]
                                                               for k in range(2 * T)])) for _ in range(n)]

fz = RewardFeaturizer(mode="ours", name="push", imsize=(hw, hw), nvp=1, scale=1.0)
fz.initialize()
F = fz.model.net.F
fz.means = [torch.randn(T, F, device="cuda") * 0.01]
fz.imgs = [torch.rand(T, hw, hw, 3, device="cuda") * 2 - 1]
paths = make_paths(a.paths)
fz.process_paths(make_paths(a.paths))                                   # warm-up (allocations)
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    fz.process_paths(paths)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
print("B200 process_paths: %d paths x %d frames %dx%d: %.1f ms / iteration, %.0f frames/s (host frames in, rewards out)"
      % (a.paths, T, hw, hw, dt * 1e3, a.paths * T / dt))
# reference loop on the CPU oracle: one B=25 source-encoder forward per path + numpy distance (base.py:232-257)
spec = fz.model.spec
P = {k: torch.from_numpy(v) for k, v in fz.model.net.state_dict().items()}
means, imgs = fz.means[0].cpu().numpy(), fz.imgs[0].cpu().numpy()
torch.set_num_threads(os.cpu_count() or 1)
cp = make_paths(a.cpu_paths)
t0 = time.perf_counter()
with torch.no_grad():
    for p in cp:
        cur = np.asarray([im[0] for im in p["env_infos"]["imgs"] if im is not None])
        x = O.normalize_u8(torch.from_numpy(cur))
        z = O.encode(spec, P, "conv", x, True)[-1].numpy()              # only the source encoder runs (TF prunes the rest)
        cost = np.sum((means - z) ** 2, axis=1) + 1.0 * np.sum((imgs - x.numpy()) ** 2, axis=(1, 2, 3))
        for j in range(T):
            p["rewards"][2 * j + 1] -= cost[j] * j * j
dc = (time.perf_counter() - t0) / a.cpu_paths
print("CPU restatement (per-path loop, %d threads): %.1f ms / path, %.0f frames/s -> %.1f s for %d paths"
      % (os.cpu_count() or 1, dc * 1e3, T / dc, dc * a.paths, a.paths))
