"""ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz with the fp64 restatement oracle.

**These goldens are SELF-DERIVED** (parity unpinned: the reference ships no golden vectors, weights or data for
this path and TensorFlow cannot be installed here — SURVEY.md §8c).  They freeze the oracle's output so that
(a) the oracle cannot drift silently and (b) the GPU box, which has no /root/reference and may run a different
CPU, compares the CUDA path against committed numbers.  Run from the repo root:

    python -m oracle.gen_golden
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imitation_from_observation_b200 import spec as S          # noqa: E402  (static config only)
from oracle import tf_semantics as O                           # noqa: E402

CASES = {
    # name: (spec, batch)
    "skipnew_16x16": (S.context_skip_new(16, 16, 3, df_dim=4, featsize=16), 3),
    "real_12x20": (S.context_ae_real(12, 20, 3, featsize=10, filters=(8, 4, 4, 4)), 2),
    "incep2_6x6": (S.context_ae_inception2(6, 6, 8, [1, 2, 1, 2], [3, 3, 3, 3], [8, 8, 4, 4], featsize=16), 2),
    # BASELINE configs[3]'s 6-layer depth (not a reference model: SURVEY.md §8d config 4), 64 -> 1 in six stride-2 steps
    "deep6_64x64": (S.context_skip_deep(64, 64, 3, filters=(4, 8, 8, 8, 16, 16), featsize=16), 2),
}


def golden_params(spec, seed):
    """He-scaled weights and small non-zero biases so every layer (and every bias path) carries signal."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name, shp in spec.param_shapes().items():
        if name.endswith("biases") or name.endswith("bias"):
            P[name] = (torch.randn(shp, generator=g) * 0.1).float()
        else:
            if name.endswith("Matrix"):
                fan_in = shp[0]
            elif "deconv/" in name:
                fan_in = shp[0] * shp[1] * shp[3] / 4.0
            else:
                fan_in = shp[0] * shp[1] * shp[2]
            P[name] = (torch.randn(shp, generator=g) / np.sqrt(fan_in)).float()
    return P


def make_case(name):
    spec, B = CASES[name]
    P32 = golden_params(spec, seed=len(name) * 7 + 1)
    g = torch.Generator().manual_seed(99)
    image = (torch.rand((3, B, spec.height, spec.width, spec.c_dim), generator=g) * 2 - 1).float()
    P64 = {k: v.double() for k, v in P32.items()}
    r, G = O.loss_and_grads(spec, P64, image.double())
    out = {"image": image.numpy()}
    for k, v in P32.items():
        out["param:" + k] = v.numpy()
    for k in ("input_z", "translated_z", "target_z", "ctx_z", "out", "out2", "simloss", "recon1", "recon2", "loss"):
        out["out:" + k] = r[k].numpy()
    for k, v in G.items():
        out["grad:" + k] = v.numpy()
    return out


def main():
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    os.makedirs(here, exist_ok=True)
    only = sys.argv[1:]
    for name in CASES:
        if only and name not in only:
            continue
        d = make_case(name)
        path = os.path.join(here, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, "%.1f KB" % (os.path.getsize(path) / 1024), "loss", float(d["out:loss"]))


if __name__ == "__main__":
    main()
