"""B200-native context-translation encoder/decoder (the hot path of wyndwarrior/imitation_from_observation).

Host side: Python + torch tensors as containers; compute: hand-written sm_100a CUDA in libctx_b200.so, bound via
ctypes (include/ctx_b200.h).  Importing this package does not need a GPU; running any op does.
"""
from . import spec, synthetic            # pure Python / numpy

__all__ = ["spec", "synthetic", "ops", "network", "models", "compat", "trainer", "sampler", "parallel"]


def __getattr__(name):
    if name in ("ops", "network", "models", "compat", "trainer", "sampler", "parallel", "_lib"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
