"""ctypes binding of libctx_b200.so (include/ctx_b200.h).  No torch types cross this boundary: only raw
device pointers (``tensor.data_ptr()``), integer sizes and a ``cudaStream_t``.

There is deliberately no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libctx_b200.so")

IMPL_AUTO, IMPL_SIMT, IMPL_TC, IMPL_TC3, IMPL_FAST, IMPL_BF16X3, IMPL_P16X1 = 0, 1, 2, 3, 4, 5, 6
IMPL_P16X3 = IMPL_BF16X3
P16_BF16, P16_F16 = 0, 1


class CtxError(RuntimeError):
    pass


class ConvGeom(Structure):
    _fields_ = [("N", c_int32), ("H", c_int32), ("W", c_int32), ("P", c_int32), ("Q", c_int32),
                ("kh", c_int32), ("kw", c_int32), ("stride", c_int32), ("pad_h", c_int32), ("pad_w", c_int32)]


class Epilogue(Structure):
    _fields_ = [("bias", c_void_p), ("act_ref", c_void_p), ("slope", c_float), ("act", c_int32),
                ("accumulate", c_int32), ("round_tf32", c_int32)]


class L2Fused(Structure):
    """ctx_l2_fused: the reconstruction loss evaluated inside ctx_conv2d_transpose_few (include/ctx_b200.h)."""
    _fields_ = [("tgt", c_void_p), ("resid", c_void_p), ("weight", c_void_p), ("loss", c_void_p), ("dpred", c_void_p),
                ("dpred_amax", c_void_p), ("pass_images", c_int32)]


class P16(Structure):
    """ctx_p16: element type and arithmetic mode of a 16-bit-planes call (include/ctx_b200.h)."""
    _fields_ = [("dtype", c_int32), ("nprod", c_int32), ("a_scale", c_void_p), ("amax_bits", c_void_p)]


# name -> (restype, argtypes); must list every symbol declared in include/ctx_b200.h
SIGNATURES = {
    "ctx_version": (c_int32, []),
    "ctx_last_error": (c_char_p, []),
    "ctx_device_info": (c_int32, [POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    "ctx_tc_available": (c_int32, []),
    "ctx_launch_count": (c_int64, []),
    "ctx_last_impl": (c_int32, []),
    "ctx_conv2d_fprop": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_int32, POINTER(ConvGeom),
                                   POINTER(Epilogue), c_int32, c_void_p]),
    "ctx_conv2d_fprop_ld": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_int32, POINTER(ConvGeom),
                                      POINTER(Epilogue), c_int32, c_void_p]),
    "ctx_conv2d_transpose": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_int32,
                                       POINTER(ConvGeom), POINTER(Epilogue), c_int32, c_void_p]),
    "ctx_conv2d_wgrad": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_int32, c_void_p,
                                   POINTER(ConvGeom), c_int32, c_void_p]),
    "ctx_gemm": (c_int32, [c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_int32,
                           c_void_p, c_int32, POINTER(Epilogue), c_int32, c_void_p]),
    "ctx_u8_normalize": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p]),
    "ctx_split_bf16": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p]),
    "ctx_split_bf16_transpose": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "ctx_conv2d_fprop_bf16x3_eligible": (c_int32, [c_int32, c_int32, POINTER(ConvGeom)]),
    "ctx_conv2d_fprop_bf16x3": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, POINTER(ConvGeom),
                                          POINTER(Epilogue), c_void_p]),
    "ctx_conv2d_transpose_bf16x3_eligible": (c_int32, [c_int32, c_int32, c_int32, c_int32, POINTER(ConvGeom)]),
    "ctx_conv2d_transpose_bf16x3": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                              c_int32, POINTER(ConvGeom), POINTER(Epilogue), c_void_p]),
    "ctx_conv2d_transpose_merged_bf16x3_eligible": (c_int32, [c_int32, c_int32, c_int32, c_int32, POINTER(ConvGeom)]),
    "ctx_conv2d_transpose_merged_rows": (c_int64, [c_int32, POINTER(ConvGeom)]),
    "ctx_conv2d_transpose_merged_pack": (c_int32, [c_void_p, c_int32, c_int32, POINTER(ConvGeom), c_void_p, c_void_p]),
    "ctx_conv2d_transpose_merged_bf16x3": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_void_p,
                                                     c_void_p, c_int32, POINTER(ConvGeom), POINTER(Epilogue), c_void_p]),
    "ctx_split_planes": (c_int32, [c_void_p, c_int64, c_float, c_void_p, c_int32, c_int32, c_void_p]),
    "ctx_split_planes_transpose": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_float, c_void_p, c_int32, c_int32,
                                             c_void_p]),
    "ctx_conv2d_fprop_p16_eligible": (c_int32, [c_int32, c_int32, POINTER(ConvGeom)]),
    "ctx_conv2d_fprop_p16": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, POINTER(ConvGeom),
                                       POINTER(Epilogue), POINTER(P16), c_void_p]),
    "ctx_conv2d_transpose_p16_eligible": (c_int32, [c_int32, c_int32, c_int32, c_int32, POINTER(ConvGeom)]),
    "ctx_conv2d_transpose_p16": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                           c_int32, c_int32, POINTER(ConvGeom), POINTER(Epilogue), POINTER(P16), c_void_p]),
    "ctx_conv2d_transpose_merged_p16_eligible": (c_int32, [c_int32, c_int32, c_int32, c_int32, POINTER(ConvGeom)]),
    "ctx_conv2d_transpose_merged_pack_p16": (c_int32, [c_void_p, c_int32, c_int32, POINTER(ConvGeom), c_float, c_void_p, c_int32,
                                                       c_int32, c_void_p]),
    "ctx_conv2d_transpose_merged_p16": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_void_p,
                                                  c_void_p, c_int32, c_int32, POINTER(ConvGeom), POINTER(Epilogue),
                                                  POINTER(P16), c_void_p]),
    "ctx_conv2d_wgrad_p16_eligible": (c_int32, [c_int32, c_int32, c_int32, c_int32, POINTER(ConvGeom)]),
    "ctx_conv2d_wgrad_p16": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_int32, c_void_p,
                                       POINTER(ConvGeom), POINTER(P16), c_void_p]),
    "ctx_grad_planes": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                  c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "ctx_amax": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p]),
    "ctx_skip_grad_sum": (c_int32, [c_void_p, c_int32, c_int64, c_int32, c_int32, c_void_p, c_float, c_void_p, c_int32,
                                    c_void_p]),
    "ctx_conv2d_few_eligible": (c_int32, [c_int32, c_int32, POINTER(ConvGeom)]),
    "ctx_conv2d_few_pack_bytes": (c_int64, [c_int32, POINTER(ConvGeom), c_int32]),
    "ctx_conv2d_few_pack": (c_int32, [c_void_p, c_int32, c_int32, POINTER(ConvGeom), c_void_p, c_int32, c_void_p]),
    "ctx_conv2d_few_fprop": (c_int32, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                       POINTER(ConvGeom), POINTER(Epilogue), c_int32, c_void_p, c_void_p, c_void_p]),
    "ctx_conv2d_few_wgrad": (c_int32, [c_void_p, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_int32, c_void_p,
                                       POINTER(ConvGeom), c_void_p, c_void_p, c_void_p]),
    "ctx_conv2d_transpose_few_eligible": (c_int32, [c_int32, c_int32, c_int32, c_int32, POINTER(ConvGeom)]),
    "ctx_conv2d_transpose_few_pack_bytes": (c_int64, [c_int32, POINTER(ConvGeom)]),
    "ctx_conv2d_transpose_few_pack": (c_int32, [c_void_p, c_int32, c_int32, POINTER(ConvGeom), c_void_p, c_void_p]),
    "ctx_conv2d_transpose_few": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_int32,
                                           POINTER(ConvGeom), POINTER(Epilogue), POINTER(L2Fused), c_void_p]),
    "ctx_adam_schedule": (c_int32, [c_void_p, c_float, c_float, c_float, c_void_p, c_void_p]),
    "ctx_adam_step_dev": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_float, c_float, c_float,
                                    c_float, c_void_p, c_void_p]),
    "ctx_gae_tail": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_double, c_double, c_void_p, c_void_p, c_void_p]),
    "ctx_crc32c": (C.c_uint32, [c_void_p, C.c_uint64, C.c_uint32]),
    "ctx_gather_frames": (c_int32, [c_void_p, c_int32, c_void_p, c_int64, c_int64, c_void_p, c_int32, c_void_p, c_void_p]),
    "ctx_lrelu_bwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_int32, c_void_p]),
    "ctx_bias_grad": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "ctx_l2_loss_fwd_bwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int64, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    "ctx_mse_scaled_fwd_bwd": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_int64, c_int32, c_float, c_float,
                                         c_float, c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int32, c_void_p]),
    "ctx_split_skip_grad": (c_int32, [c_void_p, c_int32, c_int64, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                      c_float, c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "ctx_add_bcast": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "ctx_im2col": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, POINTER(ConvGeom), c_int32, c_void_p]),
    "ctx_pool2d": (c_int32, [c_void_p, c_void_p, c_int32, POINTER(ConvGeom), c_int32, c_void_p]),
    "ctx_copy_channels": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int64, c_void_p]),
    "ctx_adam_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                                c_float, c_void_p, c_void_p]),
    "ctx_round_tf32": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p]),
    "ctx_transpose_round": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "ctx_nn_err": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ctx_reward_ours": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int64,
                                  c_float, c_float, c_float, c_void_p, c_int64, c_void_p, c_void_p]),
    "ctx_reward_inception": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int64, c_void_p, c_int64,
                                       c_void_p, c_void_p]),
}

_lib = None


def load():
    """Load (once) and return the ctypes library with typed entry points.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CtxError(
            "libctx_b200.so not found at %s — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C imitation_from_observation_b200/csrc`.  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError if the symbol is missing -> loud failure
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().ctx_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = ""):
    if rc != 0:
        raise CtxError("%s failed (code %d): %s" % (what or "libctx_b200 call", rc, last_error()))
