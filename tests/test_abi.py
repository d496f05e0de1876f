"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/ctx_b200.h declares,
the ctypes table matches the header, and without a GPU compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from imitation_from_observation_b200 import _lib

HEADER = os.path.join(ROOT, "include", "ctx_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libctx_b200.so does not export %s" % n


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\((.*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), "%s: header has %d parameters, ctypes table %d" % (name, len(params), len(args))


def test_version_and_error_string(lib):
    assert lib.ctx_version() == 200
    assert isinstance(_lib.last_error(), str)


def test_struct_layouts():
    assert ctypes.sizeof(_lib.ConvGeom) == 40
    assert ctypes.sizeof(_lib.Epilogue) == 32
    assert _lib.Epilogue.slope.offset == 16 and _lib.Epilogue.round_tf32.offset == 28


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    a, b, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    rc = lib.ctx_device_info(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    assert rc == -2 and "CUDA" in _lib.last_error()
    assert lib.ctx_tc_available() == 0
    from imitation_from_observation_b200 import ops, network, spec
    with pytest.raises(_lib.CtxError):
        ops.u8_normalize(torch.zeros(4, dtype=torch.uint8))
    with pytest.raises(_lib.CtxError):
        network.ContextNet(spec.context_skip_new(16, 16, 3, 4, 16), 2, device="cpu")


def test_argument_validation_without_device(lib):
    # argument checks come before any CUDA call, so they are testable on CPU
    g = _lib.ConvGeom(1, 8, 8, 4, 4, 5, 5, 0, 1, 1)            # stride 0
    e = _lib.Epilogue(None, None, 0.2, 0, 0, 0)
    rc = lib.ctx_conv2d_fprop(1, 3, 1, 1, 4, ctypes.byref(g), ctypes.byref(e), 0, None)
    assert rc == -1 and "geometry" in _lib.last_error()
    rc = lib.ctx_gemm(0, 0, 4, 4, 4, 1, 2, 1, 4, 1, 4, None, 0, None)
    assert rc == -1 and "leading dimension" in _lib.last_error()
