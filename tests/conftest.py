import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
# every fresh work buffer of ContextNet starts as NaN under test, so a kernel that reads a buffer nobody wrote fails
# deterministically instead of passing on whatever the caching allocator handed back (network._POISON)
os.environ.setdefault("CTX_POISON", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is visible, e.g. under a plain `pytest tests/` on CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (built on demand on a CPU box; prebuilt .so travels to the GPU box)."""
    from imitation_from_observation_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def rel_err(a, b):
    """max-norm relative error ||a-b||_inf / ||b||_inf (SURVEY.md §8c 'Oracle plan')."""
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / denom


def rel_l2(a, b):
    """relative Frobenius error ||a-b||_2 / ||b||_2 — the metric for whole-graph GRADIENTS.

    lrelu's derivative is discontinuous: a forward perturbation of relative size d flips the mask of a fraction
    ~0.8 d of the units, and every flip changes one term of a gradient sum by 80 %.  The max-norm of a gradient
    therefore jumps by percents on single flips (also between two exact-fp32 implementations that merely sum in a
    different order), while the L2 error follows 0.72 sqrt(d) (measured).  Single ops, forward
    outputs and losses are still held to the max-norm."""
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b)) / max(float(np.linalg.norm(b)), 1e-30)
