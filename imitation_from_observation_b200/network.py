"""Graph executor for the context-translation network: explicit forward, backward and TF-Adam over flat buffers.

This is the B200 replacement for "TF builds a graph, sess.run executes a pruned subgraph": the graph is static
(SURVEY.md §3.4), so forward and backward are written out once, every activation / gradient lives in a
pre-allocated HBM buffer, and each node is one call into libctx_b200.so on the current CUDA stream.

Design points:
  * frames that share encoder weights are stacked on the batch axis and run as one launch
    ([src; tgt] through scope "conv" — arm_shaping.py:1290-1307; the two decoder passes [trans_z; tgt_z]
    through scope "deconv" — :1314-1343), so shared weights are read once and their gradients are one wgrad;
  * tf.concat([h, skip], 3) (:1323-1329) is never materialised: the transposed-conv kernel takes two sources;
  * lrelu' is fused into the epilogue of the kernel that produces each data gradient;
  * parameters, gradients and Adam moments are three flat fp32 buffers with identical layout (ordered so that
    gradients complete front-to-back during backward -> bucketed all-reduce), one fused Adam launch.
"""
from __future__ import annotations

import math
import os
import weakref
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops
from ._lib import IMPL_AUTO, IMPL_FAST, IMPL_SIMT
from .spec import NetSpec

FWD_SRC, FWD_TRANS, FWD_ALL = 1, 2, 3
_ALIGN = 64          # floats; every parameter starts on a 256-byte boundary (TMA / 128-bit access friendly)


_POISON = os.environ.get("CTX_POISON", "0") == "1"      # fill every fresh work buffer with NaN (set by tests/conftest.py)


def _numel(shape) -> int:
    n = 1
    for d in shape:
        n *= int(d)
    return n


_LIVE_NETS: "weakref.WeakSet" = weakref.WeakSet()


def release_all_graphs():
    """Destroy the captured training-step graphs of every live net.  A data-parallel graph holds NCCL kernel nodes: it has to
    go BEFORE the process group does (destroying the communicator under a live graph hangs at exit), so
    parallel.shutdown() and an atexit hook call this first."""
    for net in list(_LIVE_NETS):
        net.release_graphs()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


class ContextNet:
    """One network instance bound to a device, a per-rank batch size and a precision mode.

    precision = "mixed" (default): forward fp32-grade — convolutions on fp16 (hi, lo) operand planes, three products
                        (csrc/tc_h16.cu), the 3-pass TF32 split where planes do not apply; backward (data and filter
                        gradients) single-product on the fp16 hi planes = TF32's 11-bit significand ROUNDED TO NEAREST
                        by the producer (gradients scaled per tensor into fp16's range), single-pass TF32 for the dense
                        and few-channel layers.  The forward pass decides every lrelu mask, so it keeps fp32-grade
                        accuracy; the backward GEMMs are smooth in their operands, where the unbiased 2^-12 rounding
                        averages out.
    precision = "tf32x3": tcgen05 kernels with the 3-pass TF32 split (hi*hi + hi*lo + lo*hi, fp32
                        accumulate; ~2^-21 per product, i.e. fp32-faithful) wherever the shape qualifies, the exact
                        fp32 SIMT kernels elsewhere.  The strictest tensor-core mode.
    precision = "fp32": every GEMM-shaped op uses the exact-fp32 SIMT kernels.
    precision = "tf32": single-pass TF32 (operands rounded to nearest by the producing epilogue) wherever the
                        shape qualifies.  Fast mode: forward outputs stay within ~1e-3, but lrelu masks flip for
                        ~3e-4 of the units, so gradients only agree in aggregate (relative L2 ~1e-2).
    """

    def __init__(self, spec: NetSpec, batch: int, device="cuda", precision: str = "mixed", seed: int = 0,
                 global_batch: Optional[int] = None, ablation_type: str = "None", params_from: "ContextNet" = None):
        if precision not in ("fp32", "tf32x3", "tf32", "mixed"):
            raise ValueError("precision must be 'mixed', 'tf32x3', 'tf32' or 'fp32'")
        _lib.load()                                   # fail loudly if the extension is missing
        self.spec = spec
        self.B = int(batch)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CtxError("ContextNet needs a CUDA device: the B200 path has no CPU fallback")
        self.precision = precision
        self.impl = {"fp32": IMPL_SIMT, "tf32x3": IMPL_AUTO, "tf32": IMPL_FAST, "mixed": IMPL_AUTO}[precision]
        self.impl_b = IMPL_FAST if precision == "mixed" else self.impl       # backward GEMMs
        self.round = precision == "tf32"                                      # round forward activations to TF32
        self.round_b = precision in ("tf32", "mixed")                         # round backward gradients to TF32
        self.use_tc = precision != "fp32"
        self.global_batch = int(global_batch) if global_batch else self.B
        self.ablation_type = ablation_type
        self._w_recon = {"None": (1.0, 1.0), "L2": (1.0, 1.0), "L2L3": (1.0, 0.0), "L1": (0.0, 1.0)}[ablation_type]
        self._w_sim = {"None": 1.0, "L2": 0.0, "L2L3": 0.0, "L1": 1.0}[ablation_type]
        self.step_count = 0
        self.concurrent = os.environ.get("CTX_CONCURRENT", "1") != "0"
        self._s2 = None
        self._s3 = None
        self._s4 = None                                 # backward: the context encoder's chain (data gradients) ...
        self._s5 = None                                 # ... and its filter gradients
        self._copy_stream = None
        self._in_events = None
        self._in_slot = 0                               # input double buffering (prefetch_input)
        self._slot_free = [None, None]
        self._prefetched = None
        self.allreduce = None          # parallel.BucketedAllReduce when data-parallel (attach_allreduce)
        self._bufs: Dict[str, torch.Tensor] = {}
        _LIVE_NETS.add(self)

        # ---- flat parameter / gradient / moment buffers -----------------------------------------------------
        self.shapes = spec.param_shapes()
        self.offsets: Dict[str, int] = {}
        off = 0
        for name, shp in self.shapes.items():
            self.offsets[name] = off
            off += -(-_numel(shp) // _ALIGN) * _ALIGN
        self.flat_size = off
        if params_from is not None:
            if params_from.spec.param_shapes() != self.shapes:
                raise ValueError("params_from has a different architecture")
            self.flat_p = params_from.flat_p
        else:
            self.flat_p = torch.zeros(off, device=self.device, dtype=torch.float32)
        self.flat_g = None
        self.flat_m = None
        self.flat_v = None
        self.P = {n: self._view(self.flat_p, n) for n in self.shapes}
        self.G: Dict[str, torch.Tensor] = {}
        # Single-pass tensor-core GEMMs read a shadow copy of the parameters rounded to the nearest TF32 (refreshed by
        # the Adam kernel): the hardware truncates fp32 operands, which would shrink every product by ~2^-11.
        self.flat_pr = None
        fwd_single, bwd_single = precision == "tf32", precision in ("tf32", "mixed")
        if fwd_single or bwd_single:
            if params_from is not None and params_from.flat_pr is not None:
                self.flat_pr = params_from.flat_pr
            else:
                self.flat_pr = torch.zeros(off, device=self.device, dtype=torch.float32)
        PR = {n: self._view(self.flat_pr, n) for n in self.shapes} if self.flat_pr is not None else self.P
        self.Wf = PR if fwd_single else self.P          # weights read by forward GEMMs
        self.Wb = PR if bwd_single else self.P          # weights read by backward (data-gradient) GEMMs
        if params_from is None:
            self.init_params(seed)
        else:
            self.sync_shadow()
        self.ctx_scope = "conv_context" if spec.separate_ctx_encoder else "conv"
        self.sizes = [(spec.height, spec.width)] + spec.enc_sizes()
        self.chans = spec.enc_channels()
        self.F = spec.featsize
        self.flat = spec.flat_dim()
        self.dec = spec.dec_layers()
        self.L = spec.depth                              # encoder convs = decoder transposed convs (4 in the reference)
        self.lin = spec.lin_name                         # "h4_lin" at the reference's depth
        # Forward convolutions on pre-split 16-bit planes (csrc/tc_h16.cu): the fp32-grade forward of "mixed" runs
        # hi*hi + hi*lo + lo*hi on the 16-bit tensor cores wherever channels % 64 == 0 and the pixel grid tiles exactly;
        # the other layers (3-channel image side, dense layers) stay on the 3-pass TF32 kernel.  CTX_FWD_BF3=0 disables.
        # Planes are fp16 (22 significant bits in hi + lo; CTX_PLANES=bf16 selects bf16: 16 bits, fp32's range); the hi
        # plane alone is the backward pass's operand (_enc_p1 / _dec_p1 below, fp16 only).
        self.pdt = torch.bfloat16 if os.environ.get("CTX_PLANES", "fp16") == "bf16" else torch.float16
        self._wsc = 256.0 if self.pdt == torch.float16 else 1.0
        self._enc_bf3 = [False] * self.L
        self._dec_bf3 = [False] * self.L
        self._enc_p1 = [False] * self.L                  # backward of encoder conv j on fp16 hi planes (one product)
        self._enc_p1_merged = [False] * self.L           # ... its data gradient in the phase-merged form (Cin = 64 / 128)
        self._dec_p1 = [False] * self.L                  # backward of decoder level i on fp16 hi planes
        self._enc_few = [False] * self.L                 # few-channel-input conv (the image): patches built on chip
        self._dec_few = [False] * self.L                 # few-channel-output level (the reconstruction): its backward
        self._dec_fewt = [False] * self.L                # ... and its forward (channel GEMM + col2im gather, fused loss)
        self._wb: Dict[str, torch.Tensor] = {}           # backward weight planes (one fp16 plane), refreshed with _wp
        self._in_planes: Dict[tuple, torch.Tensor] = {}  # (tag, j) -> planes of the input of encoder conv j (forward)
        self._dec_planes: Dict[int, tuple] = {}          # i -> (state planes, skip planes) of decoder level i (forward)
        self._dec_merged = [False] * self.L              # stride-2 levels with 64 output channels: phases merged along N
        self._wp: Dict[str, torch.Tensor] = {}           # weight planes, refreshed after every parameter update
        self._wp_stale = True
        self._wp_shared = params_from is not None        # someone else updates the parameters: refresh every forward
        if precision == "mixed" and os.environ.get("CTX_FWD_BF3", "1") != "0" and ops.tc_available():
            for j in range(self.L):
                self._enc_bf3[j] = ops.conv2d_bf16x3_eligible(2, self.sizes[j], self.chans[j], self.chans[j + 1],
                                                              spec.kernels[j], spec.strides[j])
            if spec.separate_ctx_encoder:                # (a shared encoder's skip rows are a slice: no plane view)
                for i, Ld in enumerate(self.dec):
                    args = (2 * self.B, Ld["out_hw"], Ld["c_state"], Ld["c_skip"], self.B, Ld["c_out"], Ld["k"], Ld["stride"])
                    self._dec_merged[i] = (os.environ.get("CTX_BF3_MERGE", "1") != "0" and
                                           ops.conv2d_transpose_merged_eligible(*args))
                    self._dec_bf3[i] = self._dec_merged[i] or ops.conv2d_transpose_bf16x3_eligible(*args)
            if (spec.separate_ctx_encoder and self.pdt == torch.float16 and os.environ.get("CTX_BWD_P16", "1") != "0"):
                few_on = os.environ.get("CTX_FEW", "1") != "0"
                for j in range(self.L):
                    Cin, Co, k, st = self.chans[j], self.chans[j + 1], spec.kernels[j], spec.strides[j]
                    rows = (2 * self.B, self.B)          # scope "conv" rows / context rows
                    if j == 0:                           # no data gradient below the first layer
                        self._enc_few[0] = few_on and ops.conv2d_few_eligible(2 * self.B, self.sizes[0], Cin, Co, k, st)
                        self._enc_p1[0] = self._enc_bf3[0] and all(
                            ops.conv2d_wgrad_planes_eligible(n, self.sizes[0], Cin, Co, 0, 0, k, st) for n in rows)
                        continue
                    mg = all(ops.conv2d_transpose_merged_eligible(n, self.sizes[j], Co, 0, 0, Cin, k, st) for n in rows) \
                        and os.environ.get("CTX_BF3_MERGE", "1") != "0"
                    dg = mg or all(ops.conv2d_transpose_planes_eligible(n, self.sizes[j], Co, 0, 0, Cin, k, st) for n in rows)
                    wg = all(ops.conv2d_wgrad_planes_eligible(n, self.sizes[j], Cin, Co, 0, 0, k, st) for n in rows)
                    self._enc_p1[j] = self._enc_bf3[j] and dg and wg
                    self._enc_p1_merged[j] = self._enc_p1[j] and mg
                for i, Ld in enumerate(self.dec):
                    Cs, Csk, Co, k, st = Ld["c_state"], Ld["c_skip"], Ld["c_out"], Ld["k"], Ld["stride"]
                    self._dec_p1[i] = (self._dec_bf3[i] and
                                       ops.conv2d_planes_eligible(2 * self.B, Ld["out_hw"], Co, Cs + Csk, k, st) and
                                       ops.conv2d_wgrad_planes_eligible(2 * self.B, Ld["out_hw"], Co, Cs, Csk, self.B, k, st))
                    self._dec_few[i] = (few_on and not self._dec_bf3[i] and Cs % 64 == 0 and Csk % 64 == 0 and
                                        ops.conv2d_few_eligible(2 * self.B, Ld["out_hw"], Co, Cs + Csk, k, st))
                    self._dec_fewt[i] = (self._dec_few[i] and os.environ.get("CTX_FEW_T", "1") != "0" and
                                         ops.conv2d_transpose_few_eligible(2 * self.B, Ld["out_hw"], Cs, Csk, self.B, Co, k, st))
        self.bf3 = any(self._enc_bf3) or any(self._dec_bf3) or any(self._enc_few) or any(self._dec_few)
        self.p1 = any(self._enc_p1) or any(self._dec_p1) or any(self._enc_few) or any(self._dec_few)
        self._level = 0
        # The whole training step (zero_grad + forward + backward + Adam over three streams) replayed as ONE CUDA graph: the
        # graph is static, so after two eager steps have allocated every buffer the ~120 launches of a step
        # cost one cudaGraphLaunch of host time.  One graph per (input slot, learning rate).  CTX_GRAPH=0 disables.
        self._graph_on = os.environ.get("CTX_GRAPH", "1") != "0"
        self._graphs: Dict[tuple, "torch.cuda.CUDAGraph"] = {}
        self._graph_meta: Dict[tuple, tuple] = {}
        self._capturing = False
        self._in_fresh = False
        self._eager_steps = 0
        self._step_dev = None                            # int64[1]: Adam's step counter on the device (graph replays)
        self._lr_t_dev = None
        self._step_dev_host = -1
        self._l2_fused = False                           # the last forward evaluated the reconstruction loss inside d_h<L>

    # ---- second stream -----------------------------------------------------------------------------------------
    # Independent launches run on a side stream so that one kernel's last, partly filled wave overlaps another
    # kernel: the context encoder beside the source/target encoder in forward, every filter gradient (+ bias gradient)
    # beside the data gradient of the same layer in backward.  Off while bench.py times single launches.
    def _side(self):
        """Context manager: run the enclosed launches on the side stream, after everything enqueued so far."""
        import contextlib
        if not self.concurrent or ops.PROFILE is not None:
            return contextlib.nullcontext()
        if self._s2 is None:
            self._s2 = torch.cuda.Stream(device=self.device)
        self._s2.wait_stream(torch.cuda.current_stream(self.device))
        return torch.cuda.stream(self._s2)

    def _join(self):
        """The current stream waits for everything enqueued on the side stream."""
        if self._s2 is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._s2)

    # ---- buffers ------------------------------------------------------------------------------------------------
    def _view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        o = self.offsets[name]
        shp = self.shapes[name]
        return flat[o:o + _numel(shp)].view(shp)

    def _buf(self, name: str, shape, dtype=torch.float32) -> torch.Tensor:
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(tuple(shape), device=self.device, dtype=dtype)
            if _POISON and t.is_floating_point():
                t.fill_(float("nan"))                      # tests: a read-before-write shows up as NaN, not as luck
            self._bufs[name] = t
        return t

    def _ensure_grads(self):
        if self.flat_g is None:
            self.flat_g = torch.zeros(self.flat_size, device=self.device, dtype=torch.float32)
            self.G = {n: self._view(self.flat_g, n) for n in self.shapes}

    def _ensure_moments(self):
        if self.flat_m is None:
            self.flat_m = torch.zeros(self.flat_size, device=self.device, dtype=torch.float32)
            self.flat_v = torch.zeros(self.flat_size, device=self.device, dtype=torch.float32)

    # ---- parameters ---------------------------------------------------------------------------------------------
    def init_params(self, seed: int = 0):
        """arm_shaping.py:25-29,52-55,67-79: conv truncated_normal(0.02), deconv/linear random_normal(0.02),
        biases 0.  (Host-side torch RNG; initialisation is not on the hot path.)"""
        g = torch.Generator().manual_seed(seed)
        for name, shp in self.shapes.items():
            if name.endswith("biases") or name.endswith("bias"):
                t = torch.zeros(shp)
            else:
                t = torch.randn(shp, generator=g) * 0.02
                if "_conv/w" in name:
                    for _ in range(8):
                        bad = t.abs() > 0.04
                        if not bad.any():
                            break
                        t = torch.where(bad, torch.randn(shp, generator=g) * 0.02, t)
                    t = t.clamp(-0.04, 0.04)
            self.P[name].copy_(t.to(self.device))
        self.sync_shadow()

    def sync_shadow(self):
        """Refresh the TF32-rounded shadow parameters after the fp32 masters were written from outside."""
        if self.flat_pr is not None:
            ops.round_tf32(self.flat_p, out=self.flat_pr)
        self._wp_stale = True

    def state_dict(self) -> Dict[str, np.ndarray]:
        """TF-checkpoint-keyed parameters (SURVEY.md §8b 'Checkpoint keys')."""
        return {n: self.P[n].detach().cpu().numpy().copy() for n in self.shapes}

    def load_state_dict(self, sd: Dict[str, "np.ndarray | torch.Tensor"], prefix: str = ""):
        for n in self.shapes:
            key = prefix + n
            if key not in sd:
                raise KeyError("missing parameter %s" % key)
            v = sd[key]
            v = torch.as_tensor(np.asarray(v)) if not torch.is_tensor(v) else v
            if tuple(v.shape) != tuple(self.shapes[n]):
                raise ValueError("parameter %s has shape %s, expected %s" % (key, tuple(v.shape), self.shapes[n]))
            self.P[n].copy_(v.to(device=self.device, dtype=torch.float32))
        self.sync_shadow()

    # ---- inputs ---------------------------------------------------------------------------------------------------
    def _input_buffers(self, slot: Optional[int] = None):
        """xe: frames through scope "conv" ([src; tgt] and, for a shared encoder, [src; tgt; ctx]); xc: ctx frames.
        Two slots: the step computes on `self._in_slot` while prefetch_input() fills the other one."""
        B, (H, W), C = self.B, self.sizes[0], self.chans[0]
        n_e = 2 * B if self.spec.separate_ctx_encoder else 3 * B
        tag = "" if (self._in_slot if slot is None else slot) == 0 else "_alt"
        xe = self._buf("xe" + tag, (n_e, H, W, C))
        xc = self._buf("xc" + tag, (B, H, W, C)) if self.spec.separate_ctx_encoder else xe[2 * B:]
        return xe, xc

    def _copy_roles(self, image: torch.Tensor, dst, after):
        """Pinned host batch -> device, role by role on the copy stream, context frames first (the context encoder runs
        on the side stream and can start while the source / target frames are still in flight).  `after`: what the copy
        stream waits for before overwriting `dst`.  Returns (event: ctx landed, event: everything landed)."""
        if image.dtype != torch.float32 or not image.is_pinned():
            raise _lib.CtxError("a host input must be a pinned float32 tensor (uint8 frames are staged on the device)")
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        cs = self._copy_stream
        if isinstance(after, torch.cuda.Event):
            cs.wait_event(after)
        elif after is not None:
            cs.wait_stream(after)
        with torch.cuda.stream(cs):
            dst[1].copy_(image[1], non_blocking=True)
            ev_ctx = cs.record_event()
            dst[0].copy_(image[0], non_blocking=True)
            dst[2].copy_(image[2], non_blocking=True)
            ev_rest = cs.record_event()
        return ev_ctx, ev_rest

    def prefetch_input(self, image: torch.Tensor):
        """Start copying the NEXT step's pinned host batch into the idle input slot while the current step computes; the
        following set_input(image, use_prefetched=True) only flips the slots (any other set_input drops the prefetch).
        Returns the event of the last copy."""
        B = self.B
        if image.is_cuda or tuple(image.shape) != (3, B) + self.sizes[0] + (self.chans[0],):
            raise _lib.CtxError("prefetch_input expects a pinned host batch [3,%d,...]" % B)
        slot = 1 - self._in_slot
        xe, xc = self._input_buffers(slot)
        # the idle slot was last read by the step before the current one; its "free" event was recorded at the flip
        after = self._slot_free[slot]
        if after is None:
            after = torch.cuda.current_stream(self.device)
        evs = self._copy_roles(image, (xe[:B], xc, xe[B:2 * B]), after)
        self._prefetched = (image.data_ptr(), slot, evs)
        return evs[1]

    def set_input(self, image: torch.Tensor, use_prefetched: bool = False):
        """image [3,B,H,W,C] (0 = src, 1 = ctx, 2 = tgt — arm_shaping.py:1278-1280), uint8 or fp32, on the device.
        uint8 is normalised on the fly (base.py:117-119)."""
        B = self.B
        pre, self._prefetched = self._prefetched, None
        self._in_fresh = False
        if use_prefetched:
            if pre is None or not torch.is_tensor(image) or pre[0] != image.data_ptr():
                raise _lib.CtxError("set_input(use_prefetched=True): no matching prefetch_input() is pending")
            # this batch is already on its way into the other slot (prefetch_input): flip
            self._slot_free[self._in_slot] = torch.cuda.current_stream(self.device).record_event()
            self._in_slot = pre[1]
            self._in_events = pre[2]
            return
        xe, xc = self._input_buffers()
        dst = (xe[:B], xc, xe[B:2 * B])
        self._in_events = None
        if isinstance(image, (list, tuple)):             # three device role batches (data.DeviceBank gathers)
            want = (B,) + self.sizes[0] + (self.chans[0],)
            if len(image) != 3 or any(tuple(r.shape) != want or not r.is_cuda for r in image):
                raise _lib.CtxError("a role list must hold three CUDA tensors of shape %s" % (want,))
            for r in range(3):
                if image[r].dtype == torch.uint8:
                    ops.u8_normalize(image[r].contiguous(), out=dst[r])
                else:
                    dst[r].copy_(image[r])
            return
        if tuple(image.shape) != (3, B) + self.sizes[0] + (self.chans[0],):
            raise _lib.CtxError("input must be [3,%d,%d,%d,%d], got %s"
                                % ((B,) + self.sizes[0] + (self.chans[0],) + (tuple(image.shape),)))
        if not image.is_cuda:
            # pinned host batch (Session.run): copy role by role on a copy stream, context frames first, so that the
            # context encoder (side stream) starts while the source / target frames are still in flight
            # (the copy stream first waits for the current stream: the previous step is done with the input buffers)
            self._in_events = self._copy_roles(image, dst, torch.cuda.current_stream(self.device))
            self._in_fresh = True                          # copies just issued: an eager step starts under them, role by role
            return
        for r in range(3):
            if image.dtype == torch.uint8:
                ops.u8_normalize(image[r].contiguous(), out=dst[r])
            else:
                dst[r].copy_(image[r])

    def _wait_input(self, which: int):
        """Make the current stream wait for the staged host input: 0 = context frames, 1 = everything."""
        if self._in_events is not None and not self._capturing:
            torch.cuda.current_stream(self.device).wait_event(self._in_events[which])

    # ---- 16-bit (hi, lo) planes --------------------------------------------------------------------------------------
    def _planes(self, name: str, shape, n: int = 2) -> torch.Tensor:
        return self._buf(name, (n,) + tuple(shape), dtype=self.pdt)

    def _wscale(self) -> Optional[torch.Tensor]:
        """Device scalar the forward plane layers divide their result by: the (hi, lo) filter planes hold 2^8 * w when the
        planes are fp16 — filters are ~N(0, 0.02), whose lo parts would otherwise fall into fp16's subnormals (19 bits)."""
        if self._wsc == 1.0:
            return None
        t = self._bufs.get("wscale")
        if t is None:
            t = self._bufs["wscale"] = torch.full((1,), self._wsc, device=self.device, dtype=torch.float32)
        return t

    def _split(self, x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        return ops.split_planes(x, out=out, nplanes=out.shape[0])

    def _refresh_weight_planes(self, scope: Optional[str] = None):
        """Operand planes of the filters, from the fp32 masters (one small launch per filter and layout): (hi, lo) planes
        for the three-product forward layers, one hi plane in the data gradient's layout for the backward layers."""
        if not self.bf3:
            return
        sp = self.spec
        scopes = ["conv"] + (["conv_context"] if sp.separate_ctx_encoder else [])
        for sc in scopes:
            if scope not in (None, sc):
                continue
            for j in range(self.L):
                name = "%s/h%d_conv/w" % (sc, j)
                k = sp.kernels[j]
                if self._enc_few[j]:
                    self._wp[name] = self._bufs["wp:" + name] = ops.conv2d_few_pack(self.P[name], 2,
                                                                                     out=self._bufs.get("wp:" + name))
                if self._enc_bf3[j]:
                    buf = self._planes("wp:" + name, (k * k * self.chans[j + 1], self.chans[j]))
                    self._wp[name] = ops.split_planes_transpose(self.P[name], out=buf, scale=self._wsc)
                if self._enc_p1[j]:
                    # data gradient = transposed conv of dY with the HWIO filter as stored: rows (tap, Cin), K = Cout
                    if self._enc_p1_merged[j]:
                        self._wb[name] = ops.conv2d_transpose_merged_pack(self.P[name], self.sizes[j], sp.strides[j],
                                                                          out=self._bufs.get("wb:" + name), nplanes=1,
                                                                          dtype=self.pdt)
                        self._bufs["wb:" + name] = self._wb[name]
                    else:
                        buf = self._planes("wb:" + name, (k * k * self.chans[j], self.chans[j + 1]), n=1)
                        self._wb[name] = ops.split_planes(self.P[name], out=buf, nplanes=1)
        if scope in (None, "deconv"):
            for i, Ld in enumerate(self.dec):
                name = "deconv/%s/w" % Ld["name"]
                Ck = Ld["c_state"] + Ld["c_skip"]
                if self._dec_bf3[i]:
                    if self._dec_merged[i]:
                        self._wp[name] = ops.conv2d_transpose_merged_pack(self.P[name], Ld["out_hw"], Ld["stride"],
                                                                          out=self._bufs.get("wp:" + name), dtype=self.pdt,
                                                                          scale=self._wsc)
                        self._bufs["wp:" + name] = self._wp[name]
                    else:
                        buf = self._planes("wp:" + name, (Ld["k"] * Ld["k"] * Ld["c_out"], Ck))
                        self._wp[name] = ops.split_planes(self.P[name], out=buf, scale=self._wsc)
                if self._dec_fewt[i]:
                    self._wp[name] = self._bufs["wp:" + name] = ops.conv2d_transpose_few_pack(
                        self.P[name], out=self._bufs.get("wp:" + name))
                if self._dec_few[i]:                      # data gradient: few-channel-input conv of dY, filter read as HWIO
                    self._wb[name] = self._bufs["wb:" + name] = ops.conv2d_few_pack(self.P[name], 1,
                                                                                     out=self._bufs.get("wb:" + name))
                if self._dec_p1[i]:
                    # data gradient = conv of dY with the [kh,kw,Cout,Cin] filter read as HWIO: rows (tap, Cin), K = Cout
                    buf = self._planes("wb:" + name, (Ld["k"] * Ld["k"] * Ck, Ld["c_out"]), n=1)
                    self._wb[name] = ops.split_planes_transpose(self.P[name], out=buf, nplanes=1)

    # ---- forward pieces (also the functional enc1 / enc2 / translate / dec surface) ------------------------------
    def encode_convs(self, scope: str, x: torch.Tensor, outs: Sequence[torch.Tensor], tag: Optional[str] = None,
                     skip_planes: bool = False):
        """L x (conv SAME + bias + lrelu) — arm_shaping.py:1283-1286 (L = 4 there).
        tag: forward() passes a buffer tag, which enables the bf16-plane layers (the functional enc1 / enc2 surface runs
        the fp32-input kernels).  Returns the planes of each level's output where they were produced (else None);
        skip_planes: also produce them for the levels the decoder reads as skips."""
        h, hp = x, None
        L = self.L
        produced: List[Optional[torch.Tensor]] = [None] * L
        for j in range(L):
            w, b = self.Wf["%s/h%d_conv/w" % (scope, j)], self.P["%s/h%d_conv/biases" % (scope, j)]
            k, Cin, Co = self.spec.kernels[j], h.shape[3], w.shape[3]
            use_bf = tag is not None and self._enc_bf3[j]
            want_p = tag is not None and ((j + 1 < L and self._enc_bf3[j + 1]) or
                                          (skip_planes and (self._dec_bf3[L - 1 - j] or self._dec_few[L - 1 - j])))
            op = self._planes("%s_h%d_p%d" % (tag, j, h.shape[0]), outs[j].shape) if want_p else None
            if tag is not None and self._enc_few[j]:
                # the image: patch matrix built in shared memory, three products, result + its planes in one pass
                ops.conv2d_few(h, self._wp["%s/h%d_conv/w" % (scope, j)], k, Co, bias=b, stride=self.spec.strides[j],
                               act=True, out=outs[j], out_planes=op, nprod=3, amax=self._fwd_amax())
                h, hp = outs[j], op
                produced[j] = op
                continue
            if use_bf:
                if hp is None:
                    hp = self._split(h, self._planes("%s_in%d_p%d" % (tag, j, h.shape[0]), h.shape))
                self._in_planes[(tag, j)] = hp             # backward's filter gradient reads the hi plane
                ops.conv2d_planes(hp, self._wp["%s/h%d_conv/w" % (scope, j)], k, bias=b, stride=self.spec.strides[j],
                                  act=True, out=outs[j], out_planes=op, nprod=3, amax=self._fwd_amax(),
                                  a_scale=self._wscale())
                h, hp = outs[j], op
                produced[j] = op
                continue
            if self._few(Cin, k) and Co % 4 == 0:
                K = k * k * Cin
                # (forward()'s patch matrix is reused by backward; the functional enc1 / enc2 surface gets its own buffer)
                pt = self._patches("patch_%s_h%d_%d%s" % (scope, j, h.shape[0], "" if tag is not None else "_fn"), h, k,
                                   self.spec.strides[j])
                ops.gemm(pt[:, :K], w.view(K, Co), out=outs[j].view(-1, Co), bias=b, act=True, round_tf32=self.round,
                         impl=self.impl)
            else:
                ops.conv2d(h, w, b, stride=self.spec.strides[j], act=True, out=outs[j], round_tf32=self.round,
                           impl=self.impl)
            h = outs[j]
            hp = self._split(h, op) if want_p else None
            produced[j] = hp
        return produced

    def _fwd_amax(self) -> torch.Tensor:
        """int32 word: bit pattern of the largest |activation| the plane layers produced this step (fp16 planes saturate at
        65504: fetch('loss') turns NaN instead of training on clipped activations)."""
        return self._buf("losses", (8,))[4:5].view(torch.int32)

    def _gemm(self, tA, tB, M, N, K, A, lda, Bm, ldb, Cm, ldc, bias=None, act=False, act_ref=None,
              accumulate=False, round_=None, bwd=False):
        if round_ is None:
            round_ = self.round_b if bwd else self.round
        epi = _lib.Epilogue(None if bias is None else bias.data_ptr(), act_ref, ops.LRELU_SLOPE, int(act),
                            int(accumulate), int(round_))
        ops.gemm_raw(tA, tB, M, N, K, A, lda, Bm, ldb, Cm, ldc, epi, self.impl_b if bwd else self.impl)

    # ---- few-channel layers (the 3-channel image side): patch matrix + dense tensor-core GEMM ----------------------
    def _few(self, C: int, k: int) -> bool:
        """True when a conv over a C-channel tensor cannot take the TMA tap path (C % 32) but its whole filter
        footprint fits a short GEMM K: then  im2col (ctx_im2col) + ctx_gemm  replaces the SIMT implicit GEMM."""
        return self.use_tc and C % 32 != 0 and C * k * k <= 256 and ops.tc_available()

    @staticmethod
    def _patch_ld(C: int, k: int) -> int:
        return -(-(C * k * k) // 4) * 4

    def _patches(self, name: str, x: torch.Tensor, k: int, stride: int, round_tf32: bool = False,
                 reuse: bool = False) -> torch.Tensor:
        """Patch matrix of x in a named buffer; reuse=True returns the buffer forward() filled for the same x."""
        n, H, W_, C = x.shape
        ld = self._patch_ld(C, k)
        rows = n * (-(-H // stride)) * (-(-W_ // stride))
        have = self._bufs.get(name)
        if reuse and have is not None and tuple(have.shape) == (rows, ld):
            return have
        buf = self._buf(name, (rows, ld))
        ops.im2col(x, k, stride, ld, out=buf, round_tf32=round_tf32)
        return buf

    def _fwd_encoder(self, scope: str, x: torch.Tensor, hs: List[torch.Tensor], h4: torch.Tensor,
                     tag: Optional[str] = None, skip_planes: bool = False):
        n = x.shape[0]
        produced = self.encode_convs(scope, x, hs, tag=tag, skip_planes=skip_planes)
        self._gemm(False, False, n, self.F, self.flat, hs[-1].data_ptr(), self.flat,
                   self.Wf["%s/%s/Matrix" % (scope, self.lin)].data_ptr(), self.F, h4.data_ptr(), self.F,
                   bias=self.P["%s/%s/bias" % (scope, self.lin)], act=True)
        return produced

    def _fwd_hz(self, scope: str, h4_rows: torch.Tensor, out_ptr: int, ldo: int, act: bool):
        self._gemm(False, False, h4_rows.shape[0], self.F, self.F, h4_rows.data_ptr(), self.F,
                   self.Wf[scope + "/hz_lin/Matrix"].data_ptr(), self.F, out_ptr, ldo,
                   bias=self.P[scope + "/hz_lin/bias"], act=act)

    def _enc_buffers(self, n_rows: int, tag: str):
        hs = [self._buf("%s_h%d" % (tag, j), (n_rows,) + self.sizes[j + 1] + (self.chans[j + 1],)) for j in range(self.L)]
        h4 = self._buf("%s_lin" % tag, (n_rows, self.F))      # (not "_h4": that is conv level 4 of a deeper net)
        return hs, h4

    def forward(self, level: int = FWD_ALL):
        """Run the graph up to `level`:
             FWD_SRC   — source encoder only (`input_z`; the reward path, base.py:232-235)
             FWD_TRANS — + context encoder, translator, decoder on trans_z (`translated_z`, `out`; base.py:216)
             FWD_ALL   — + target encoder, decoder on tgt_z, the three losses (training, train_script.py:163)
        """
        sp, B, F = self.spec, self.B, self.F
        if self.bf3 and (self._wp_stale or self._wp_shared):
            self._refresh_weight_planes()
            self._wp_stale = False
        xe, xc = self._input_buffers()
        n_e = xe.shape[0]
        e_hs, e_h4 = self._enc_buffers(n_e, "e")
        c_planes = None
        self._buf("losses", (8,)).zero_()               # [recon1, recon2, simloss, -, bits of max |activation|, ...]
        cat = self._buf("cat", (B, 2 * F))
        z2 = self._buf("z2", (2 * B, F))
        f4 = 4  # bytes per float

        if level == FWD_ALL:
            rows = n_e                                   # [src; tgt] (+ ctx when the encoder is shared)
        elif level == FWD_TRANS and not sp.separate_ctx_encoder:
            rows = None                                  # src rows and ctx rows, two launches
        else:
            rows = B
        side_ctx = level >= FWD_TRANS and sp.separate_ctx_encoder
        if side_ctx:
            c_hs, c_h4 = self._enc_buffers(B, "c")
            with self._side():                           # the context encoder has its own weights and input
                self._wait_input(0)
                c_planes = self._fwd_encoder("conv_context", xc, c_hs, c_h4, tag="c", skip_planes=True)
                self._fwd_hz(self.ctx_scope, c_h4, cat.data_ptr() + F * f4, 2 * F, act=sp.ctx_hz_lrelu)  # ctx z
        self._wait_input(1)
        if rows is not None:
            self._fwd_encoder("conv", xe[:rows], [h[:rows] for h in e_hs], e_h4[:rows], tag="e")
        else:
            self._fwd_encoder("conv", xe[:B], [h[:B] for h in e_hs], e_h4[:B], tag="e")
            self._fwd_encoder("conv", xe[2 * B:], [h[2 * B:] for h in e_hs], e_h4[2 * B:], tag="e2")
        self._fwd_hz("conv", e_h4[:B], cat.data_ptr(), 2 * F, act=True)                    # src z -> cat[:, :F]
        if level >= FWD_TRANS:
            if side_ctx:
                self._join()
            else:
                c_hs, c_h4 = [h[2 * B:] for h in e_hs], e_h4[2 * B:]
                self._fwd_hz(self.ctx_scope, c_h4, cat.data_ptr() + F * f4, 2 * F, act=sp.ctx_hz_lrelu)  # ctx z
            th0 = self._buf("th0", (B, F))
            if level == FWD_ALL:
                with self._side():                         # independent of the translator: rows B.. of the decoder's input
                    self._fwd_hz("conv", e_h4[B:2 * B], z2.data_ptr() + B * F * f4, F, act=True)
            self._gemm(False, False, B, F, 2 * F, cat.data_ptr(), 2 * F,
                       self.Wf["translate/trans_h0/Matrix"].data_ptr(), F, th0.data_ptr(), F,
                       bias=self.P["translate/trans_h0/bias"], act=True)
            self._gemm(False, False, B, F, F, th0.data_ptr(), F, self.Wf["translate/trans_z/Matrix"].data_ptr(), F,
                       z2.data_ptr(), F, bias=self.P["translate/trans_z/bias"], act=False)      # trans_z -> z2[:B]
            nd = B
            if level == FWD_ALL:
                self._join()                               # tgt z (issued beside the translator, below) -> z2[B:]
                nd = 2 * B
            d = self._dec_buffers()
            l2 = None
            if level == FWD_ALL:                           # the reconstruction loss can ride on the last decoder layer
                losses = self._buf("losses", (8,))
                l2 = dict(tgt=xe[B:2 * B], resid=xc if sp.residual_out else None, weight=self._recon_weights(),
                          loss=losses[0:2], dpred=self._buf("g_out", tuple(d[self.L].shape)),
                          amax=losses[5:6].view(torch.int32), pass_images=B)
            self._fwd_decoder(z2[:nd], [c[:B] for c in c_hs], [t[:nd] for t in d], skip_planes=c_planes, tag="d", l2=l2)
            if level == FWD_ALL:
                self._fwd_losses(d[self.L], xe[B:2 * B], xc)
        self._level = level

    def _dec_buffers(self):
        B = self.B
        d = [self._buf("d0", (2 * B,) + self.sizes[self.L] + (self.chans[self.L],))]
        for i, L in enumerate(self.dec):
            d.append(self._buf("d%d" % (i + 1), (2 * B,) + tuple(L["out_hw"]) + (L["c_out"],)))
        return d

    def _fwd_decoder(self, z: torch.Tensor, skips: Sequence[torch.Tensor], d: Sequence[torch.Tensor],
                     skip_planes: Optional[Sequence[Optional[torch.Tensor]]] = None, tag: Optional[str] = None,
                     l2: Optional[dict] = None):
        """arm_shaping.py:1314-1329 for all stacked passes at once; skips = ctx (h0..h_{L-1}), batch B.
        tag (forward() only) enables the bf16-plane levels; skip_planes[j] = planes of skips[j] where the context
        encoder produced them."""
        n = z.shape[0]
        self._l2_fused = False
        dp = None                                          # planes of the running decoder state
        self._gemm(False, False, n, self.flat, self.F, z.data_ptr(), self.F,
                   self.Wf["deconv/d_h0_lin/Matrix"].data_ptr(), self.flat, d[0].data_ptr(), self.flat,
                   bias=self.P["deconv/d_h0_lin/bias"], act=True)
        for i, L in enumerate(self.dec):
            last = i == self.L - 1
            if tag is not None and self._dec_bf3[i] and skips[self.L - 1 - i].shape[0] == self.B:
                j = self.L - 1 - i
                if dp is None:
                    dp = self._split(d[i], self._planes("%s%d_in_p%d" % (tag, i, n), d[i].shape))
                sk = skip_planes[j] if skip_planes is not None else None
                if sk is None:
                    sk = self._split(skips[j].contiguous(), self._planes("%s_skip%d_p" % (tag, j), skips[j].shape))
                self._dec_planes[i] = (dp, sk)             # backward's filter gradient reads the hi planes
                op = None
                if not last and (self._dec_bf3[i + 1] or self._dec_few[i + 1]):
                    op = self._planes("%s%d_p%d" % (tag, i + 1, n), d[i + 1].shape)
                if self._dec_merged[i]:
                    ops.conv2d_transpose_merged_planes(dp, self._wp["deconv/%s/w" % L["name"]], L["out_hw"], L["k"],
                                                       L["c_out"], bias=self.P["deconv/%s/biases" % L["name"]],
                                                       stride=L["stride"], x2p=sk, act=not last, out=d[i + 1],
                                                       out_planes=op, nprod=3, amax=self._fwd_amax(),
                                                       a_scale=self._wscale())
                else:
                    ops.conv2d_transpose_planes(dp, self._wp["deconv/%s/w" % L["name"]], L["out_hw"], L["k"],
                                                bias=self.P["deconv/%s/biases" % L["name"]], stride=L["stride"], x2p=sk,
                                                act=not last, out=d[i + 1], out_planes=op, nprod=3,
                                                amax=self._fwd_amax(), a_scale=self._wscale())
                dp = op
                continue
            if tag is not None and self._dec_few[i] and skips[self.L - 1 - i].shape[0] == self.B:
                j = self.L - 1 - i                         # backward's filter gradient reads the hi planes of [state ‖ skip]
                if dp is None:
                    dp = self._split(d[i], self._planes("%s%d_in_p%d" % (tag, i, n), d[i].shape))
                sk = skip_planes[j] if skip_planes is not None else None
                if sk is None:
                    sk = self._split(skips[j].contiguous(), self._planes("%s_skip%d_p" % (tag, j), skips[j].shape))
                self._dec_planes[i] = (dp, sk)
                if self._dec_fewt[i]:
                    # few output channels: channel contraction on the tensor cores, col2im as a gather from shared memory,
                    # and (training) the reconstruction loss + d loss / d out in the same pass
                    ops.conv2d_transpose_few(dp, self._wp["deconv/%s/w" % L["name"]], L["out_hw"], L["k"], L["c_out"],
                                             bias=self.P["deconv/%s/biases" % L["name"]], stride=L["stride"], x2p=sk,
                                             act=not last, out=d[i + 1], l2=l2 if last else None)
                    self._l2_fused = last and l2 is not None
                    dp = None
                    continue
            dp = None
            ops.conv2d_transpose(d[i], self.Wf["deconv/%s/w" % L["name"]], L["out_hw"],
                                 bias=self.P["deconv/%s/biases" % L["name"]], stride=L["stride"], x2=skips[self.L - 1 - i],
                                 act=not last, out=d[i + 1], round_tf32=self.round and not last, impl=self.impl)

    def _recon_weights(self) -> torch.Tensor:
        w = self._bufs.get("w_recon")
        if w is None:
            w = self._bufs["w_recon"] = torch.tensor(self._w_recon, device=self.device, dtype=torch.float32)
        return w

    def _fwd_losses(self, d4: torch.Tensor, tgt: torch.Tensor, ctx: torch.Tensor):
        B, F = self.B, self.F
        losses = self._buf("losses", (8,))
        if not self._l2_fused:
            g_d4 = self._buf("g_out", tuple(d4.shape))
            ops.l2_loss_fwd_bwd(d4, tgt, losses[0:2], dpred=g_d4, resid=ctx if self.spec.residual_out else None,
                                weight=self._recon_weights())
        z2 = self._bufs["z2"]
        ops.mse_scaled_fwd_bwd(z2[:B], z2[B:], 1e3, 1.0 / (B * F), loss=losses[2:3])      # reported: local mean

    # ---- backward -----------------------------------------------------------------------------------------------------
    def zero_grad(self):
        self._ensure_grads()
        self.flat_g.zero_()

    def backward(self, on_stage=None, on_dense=None):
        """Gradient of the (ablated) loss w.r.t. every parameter, accumulated into the flat gradient buffer.
        `on_stage(scope)` is called as soon as all gradients of a parameter scope ("deconv", "translate", "conv",
        "conv_context" — the order of the flat buffer) are enqueued, so a bucketed all-reduce can start early.
        `on_dense(scope)` is called inside an encoder scope as soon as the gradients of its dense layers (hz_lin, h<L>_lin:
        the first two thirds of the scope's parameters) AND the last readers of those weights are enqueued — the former on
        the side stream current at that moment (self._s2), the latter on the current stream."""
        if self._level != FWD_ALL:
            raise _lib.CtxError("backward() needs forward(FWD_ALL) first")
        self._ensure_grads()
        sp, B, F, flat = self.spec, self.B, self.F, self.flat
        P, G, bufs, Wb = self.P, self.G, self._bufs, self.Wb
        f4 = 4
        xe, xc = self._input_buffers()
        n_e = xe.shape[0]
        e_hs, e_h4 = self._enc_buffers(n_e, "e")
        if sp.separate_ctx_encoder:
            c_hs, c_h4 = self._enc_buffers(B, "c")
        else:
            c_hs, c_h4 = [h[2 * B:] for h in e_hs], e_h4[2 * B:]
        d = self._dec_buffers()
        cat, z2, th0 = bufs["cat"], bufs["z2"], bufs["th0"]

        # gradient buffers of the encoder activations (pre-activation gradients, i.e. lrelu' already applied)
        ge_hs = [self._buf("g_e_h%d" % j, tuple(e_hs[j].shape)) for j in range(self.L)]
        ge_h4 = self._buf("g_e_lin", tuple(e_h4.shape))
        if sp.separate_ctx_encoder:
            gc_hs = [self._buf("g_c_h%d" % j, tuple(c_hs[j].shape)) for j in range(self.L)]
            gc_h4 = self._buf("g_c_lin", tuple(c_h4.shape))
        else:
            gc_hs, gc_h4 = [g[2 * B:] for g in ge_hs], ge_h4[2 * B:]

        # ---- activations read by the single-pass TF32 dense filter gradients ----------------------------------------------------
        # The tensor core TRUNCATES fp32 operands to TF32: gradients and weights reach it rounded to nearest (producer
        # epilogues, shadow parameters), but the forward activations are exact fp32, and truncating them shrinks every dense
        # filter gradient by ~6e-4 (profiles/r2_grad_check_p16_v1.txt).  A rounded copy costs one pass over a few MB.
        def rn(t):
            if not self.round_b or self.round:
                return t
            buf = self._buf("rn:%x" % t.data_ptr(), tuple(t.shape))
            return ops.round_tf32(t, out=buf)

        # ---- gradients as fp16 operand planes (the one-product layers) ----------------------------------------------------
        # A gradient tensor that feeds a one-product GEMM exists as ONE fp16 plane scaled by a power of two chosen from the
        # tensor's max |.| (fp16 keeps TF32's significand but only 5 exponent bits).  The max comes from the epilogue of the
        # kernel that produced the fp32 gradient (or one ctx_amax pass); ctx_grad_planes applies lrelu', rounds to nearest
        # and forms the bias gradient in the same pass.
        slots = scales = None
        if self.p1:
            slots = self._buf("g_amax", (64,), dtype=torch.int32)
            slots.zero_()
            scales = self._buf("g_scale", (64,))
        nslot = [0]

        def new_slot():
            i = nslot[0]
            nslot[0] += 1
            return slots[i:i + 1], scales[i:i + 1]

        def to_planes(x, C_, name, ldx=None, ref=None, db=None, slot=None):
            """fp32 gradient rows -> (fp16 plane [1, rows, C_] holding scale * x * lrelu'(ref), scale); slot: (amax word the
            producer of x filled, scale word) — without it one pass over x finds the max."""
            if slot is None:
                slot = new_slot()
                ops.amax(x, slot[0])
            rows = x.numel() // (ldx or C_)
            gp = self._buf(name, (1, rows, C_), dtype=torch.float16)
            ops.grad_planes(x, C_, planes=gp, ref=ref, amax=slot[0], scale=slot[1], db=db, ldx=ldx)
            return gp, slot[1]

        # ---- decoder: d_h<L> .. d_h1 ----------------------------------------------------------------------------
        Ld = self.L
        g_out = bufs["g_out"]                             # d loss / d d4 (written by the loss kernel)
        gp = None                                          # (plane, scale) of g_out when the level runs on planes
        bias_done = False                                  # this level's bias gradient came with its planes
        ib, rb = self.impl_b, self.round_b
        for i in reversed(range(Ld)):
            L = self.dec[i]
            wname = "deconv/%s/w" % L["name"]
            state, skip = d[i], c_hs[Ld - 1 - i]
            Cs, Ck, Co, k = L["c_state"], L["c_state"] + L["c_skip"], L["c_out"], L["k"]
            dcat = self._buf("dcat%d" % i, (2 * B,) + tuple(L["in_hw"]) + (Ck,))
            slot_d = None
            if self._dec_p1[i]:
                if gp is None:
                    gp = to_planes(g_out, Co, "gp_d%d" % (i + 1), db=G["deconv/%s/biases" % L["name"]])
                gp5 = gp[0].view((1,) + tuple(d[i + 1].shape))
                dpl, skp = self._dec_planes[i]
                with self._side():                         # filter gradient: big = dY planes, small = [state ‖ skip] hi planes
                    ops.conv2d_wgrad_planes(gp5, dpl[:1], k, L["stride"], small2p=skp[:1], dw=G[wname], nprod=1,
                                            a_scale=gp[1])
                slot_d = new_slot()
                ops.conv2d_planes(gp5, self._wb[wname], k, stride=L["stride"], out=dcat, nprod=1, a_scale=gp[1],
                                  amax=slot_d[0])
            elif self._dec_few[i] and i in self._dec_planes:
                # few output channels (the reconstruction): both gradients are few-channel-INPUT convolutions over dY, patch
                # matrix built on chip; dY is scaled into fp16's range from its max
                slot_g = new_slot()
                if i == Ld - 1 and self._l2_fused:         # max |d loss / d out| came with the fused loss
                    slot_g = (bufs["losses"][5:6].view(torch.int32), slot_g[1])
                else:
                    ops.amax(g_out, slot_g[0])
                dpl, skp = self._dec_planes[i]
                with self._side():
                    ops.conv2d_few_wgrad(g_out, dpl[:1], k, L["stride"], small2p=skp[:1], dw=G[wname], x_amax=slot_g[0])
                    if not bias_done:
                        ops.bias_grad(g_out, G["deconv/%s/biases" % L["name"]])
                slot_d = new_slot()
                ops.conv2d_few(g_out, self._wb[wname], k, Ck, stride=L["stride"], out=dcat, nprod=1, x_amax=slot_g[0],
                               amax=slot_d[0])
            elif self._few(Co, k):
                # few output channels (the image): patch matrix of g_out, then three dense GEMMs
                K = k * k * Co
                pt = self._patches("patch_g_d%d" % (i + 1), g_out, k, L["stride"], round_tf32=rb)[:, :K]
                rows = pt.shape[0] // 2                    # rows of one decoder pass (the skip is shared by both)
                dw2 = G[wname].view(K, Ck)
                with self._side():
                    ops.gemm(pt, state.view(-1, Cs), out=dw2[:, :Cs], transA=True, accumulate=True, impl=ib)
                    for g_ in range(2):
                        ops.gemm(pt[g_ * rows:(g_ + 1) * rows], skip.view(-1, Ck - Cs), out=dw2[:, Cs:], transA=True,
                                 accumulate=True, impl=ib)
                    if not bias_done:
                        ops.bias_grad(g_out, G["deconv/%s/biases" % L["name"]])
                ops.gemm(pt, Wb[wname].view(K, Ck), out=dcat.view(-1, Ck), impl=ib)
            else:
                # filter gradient: big = g_out [2B, out_hw, Cout], small = [state ‖ skip] (side stream)
                with self._side():
                    ops.conv2d_wgrad(g_out, state, k, L["stride"], small2=skip, dw=G[wname], impl=ib)
                    if not bias_done:
                        ops.bias_grad(g_out, G["deconv/%s/biases" % L["name"]])
                # data gradient of the concatenated input = SAME conv of g_out with the [kh,kw,Cout,Cin] filter as HWIO
                ops.conv2d(g_out, Wb[wname], None, stride=L["stride"], act=False, out=dcat, impl=ib)
            if i > 0 and self._dec_p1[i - 1]:
                # the level below runs on planes: state half -> lrelu' -> scaled fp16 plane + its bias gradient in one
                # pass; skip half -> sum of the two decoder passes -> lrelu' -> the context encoder's fp32 gradient
                gp = to_planes(dcat, Cs, "gp_d%d" % i, ldx=Ck, ref=state, slot=slot_d,
                               db=G["deconv/%s/biases" % self.dec[i - 1]["name"]])
                with self._side():                         # only the context encoder's chain reads it, much later
                    ops.skip_grad_sum(dcat, 2, Cs, Ck - Cs, gc_hs[Ld - 1 - i], skip_ref=skip)
                g_out, bias_done = None, True
                continue
            g_state = self._buf("g_d%d" % i, tuple(state.shape))
            ops.split_skip_grad(dcat, 2, B, Cs, Ck - Cs, g_state, gc_hs[Ld - 1 - i],
                                state_ref=state, skip_ref=skip, accumulate_skip=False, round_tf32=rb)
            g_out, gp, bias_done = g_state, None, False
        # ---- d_h0_lin ------------------------------------------------------------------------------------------------
        g_d0 = g_out                                       # [2B, flat], lrelu' applied
        Wd0 = Wb["deconv/d_h0_lin/Matrix"]
        with self._side():
            self._gemm(True, False, F, flat, 2 * B, rn(z2).data_ptr(), F, g_d0.data_ptr(), flat,
                       G["deconv/d_h0_lin/Matrix"].data_ptr(), flat, accumulate=True, round_=False, bwd=True)
            ops.bias_grad(g_d0.view(2 * B, flat), G["deconv/d_h0_lin/bias"])
        g_z2 = self._buf("g_z2", (2 * B, F))
        self._gemm(False, True, 2 * B, F, flat, g_d0.data_ptr(), flat, Wd0.data_ptr(), flat, g_z2.data_ptr(), F, bwd=True)
        if on_stage:
            self._join()                                   # every decoder gradient is enqueued and ordered
            on_stage("deconv")
        # ---- feature-alignment loss gradient (mean over the GLOBAL batch so DP ranks sum correctly) ---------------
        if self._w_sim != 0.0:
            ops.mse_scaled_fwd_bwd(z2[:B], z2[B:], 1e3, 1.0 / (self.global_batch * F), da=g_z2[:B], db=g_z2[B:],
                                   grad_weight=self._w_sim, accumulate=True)
        g_transz = g_z2[:B]                                # trans_z is linear
        g_tz = g_z2[B:]
        ops.lrelu_bwd(g_tz, z2[B:], out=g_tz)              # tgt_z = lrelu(hz_lin(.))
        if self.round_b:
            # the loss gradient was added in exact fp32: back to nearest-TF32 before the single-pass GEMMs below read it
            # (the tensor core would truncate instead, a one-sided error that every upstream gradient inherits)
            ops.round_tf32(g_z2, out=g_z2)
        # ---- translator -------------------------------------------------------------------------------------------------
        # (every filter / bias gradient of these small layers goes to the side stream: the data-gradient chain below is
        # a sequence of ~20 us launches that leave most SMs idle)
        with self._side():
            self._gemm(True, False, F, F, B, rn(th0).data_ptr(), F, g_transz.data_ptr(), F,
                       G["translate/trans_z/Matrix"].data_ptr(), F, accumulate=True, round_=False, bwd=True)
            ops.bias_grad(g_transz, G["translate/trans_z/bias"])
        g_th0 = self._buf("g_th0", (B, F))
        self._gemm(False, True, B, F, F, g_transz.data_ptr(), F, Wb["translate/trans_z/Matrix"].data_ptr(), F,
                   g_th0.data_ptr(), F, act_ref=th0.data_ptr(), bwd=True)
        with self._side():
            self._gemm(True, False, 2 * F, F, B, rn(cat).data_ptr(), 2 * F, g_th0.data_ptr(), F,
                       G["translate/trans_h0/Matrix"].data_ptr(), F, accumulate=True, round_=False, bwd=True)
            ops.bias_grad(g_th0, G["translate/trans_h0/bias"])
        g_cat = self._buf("g_cat", (B, 2 * F))
        Wt0 = Wb["translate/trans_h0/Matrix"]               # [2F, F]
        # src half: masked by lrelu'(src_z); ctx half: masked only if the ctx hz_lin has an lrelu
        self._gemm(False, True, B, F, F, g_th0.data_ptr(), F, Wt0.data_ptr(), F, g_cat.data_ptr(), 2 * F,
                   act_ref=cat.data_ptr(), bwd=True)
        self._gemm(False, True, B, F, F, g_th0.data_ptr(), F, Wt0.data_ptr() + F * F * f4, F,
                   g_cat.data_ptr() + F * f4, 2 * F,
                   act_ref=(cat.data_ptr() + F * f4) if sp.ctx_hz_lrelu else None, bwd=True)
        # The context encoder's backward chain (hz_lin, h<L>_lin, its convs) depends on nothing the
        # source / target encoder's chain produces: with its own weights it runs on its own pair of streams (data gradients /
        # filter gradients) beside that chain.  Its kernels see B images — 32 ... 128 CTAs for the upper layers — so alone they
        # leave most SMs idle.
        fine = on_stage is not None and getattr(getattr(on_stage, "__self__", None), "fine", False)
        # (under the bucketed all-reduce — `fine` is set whenever on_stage is BucketedAllReduce.stage_done — the chain stays on
        # the main stream unless CTX_DP_CTX_STREAM=1: with NCCL's spinning kernels and more streams than the default eight
        # hardware queues the 2-GPU bench stalled; CUDA_DEVICE_MAX_CONNECTIONS=32 cured it (4.90 vs 5.00 ms per step, one run))
        ctx_par = (sp.separate_ctx_encoder and self.concurrent and ops.PROFILE is None and
                   (not fine or os.environ.get("CTX_DP_CTX_STREAM", "0") == "1") and
                   os.environ.get("CTX_BWD_CTX_STREAM", "1") != "0")
        cur = torch.cuda.current_stream(self.device)
        # (every reader of the translator's weights is enqueued on THIS stream above: the optimiser slice on_stage("translate")
        # starts may overwrite them as soon as what is enqueued here has run)
        self._join()                                       # filter gradients so far, and the skip-gradient sums the chain reads
        if ctx_par:
            if self._s4 is None:
                self._s4 = torch.cuda.Stream(device=self.device)
                self._s5 = torch.cuda.Stream(device=self.device)
            self._s4.wait_stream(cur)                      # everything the chain reads (g_cat, the skip gradients) is enqueued
        if on_stage:
            on_stage("translate")
        # ---- hz_lin of the three roles -> gradient w.r.t. h4 (lrelu' fused) -------------------------------------------
        def hz_backward(scope, g_ptr, ldg, h4_rows, gh4_rows):
            n = h4_rows.shape[0]
            with self._side():
                self._gemm(True, False, F, F, n, rn(h4_rows).data_ptr(), F, g_ptr, ldg,
                           G[scope + "/hz_lin/Matrix"].data_ptr(), F, accumulate=True, round_=False, bwd=True)
                check = _lib.check
                check(_lib.load().ctx_bias_grad(g_ptr, n, F, ldg, G[scope + "/hz_lin/bias"].data_ptr(), ops._stream()),
                      "ctx_bias_grad")
            self._gemm(False, True, n, F, F, g_ptr, ldg, Wb[scope + "/hz_lin/Matrix"].data_ptr(), F,
                       gh4_rows.data_ptr(), F, act_ref=h4_rows.data_ptr(), bwd=True)

        hz_backward("conv", g_cat.data_ptr(), 2 * F, e_h4[:B], ge_h4[:B])                     # src
        hz_backward("conv", g_tz.data_ptr(), F, e_h4[B:2 * B], ge_h4[B:2 * B])               # tgt
        if not ctx_par:
            hz_backward(self.ctx_scope, g_cat.data_ptr() + F * f4, 2 * F, c_h4, gc_h4)       # ctx

        # ---- encoders: h<L>_lin then the convs -----------------------------------------------------------------------------
        # (finer all-reduce stages, data parallel: issued from the stream that wrote the gradients, no join needed)
        def encoder_backward(scope, x, hs, h4, g_hs, g_h4, groups, tag):
            """groups: [(row0, rows, accumulate)] — ctx rows already hold the skip gradients, so they accumulate.
            tag: forward()'s buffer tag of this encoder pass (its operand planes)."""
            n = x.shape[0]
            with self._side():
                self._gemm(True, False, flat, F, n, rn(hs[-1]).data_ptr(), flat, g_h4.data_ptr(), F,
                           G["%s/%s/Matrix" % (scope, self.lin)].data_ptr(), F, accumulate=True, round_=False, bwd=True)
                ops.bias_grad(g_h4, G["%s/%s/bias" % (scope, self.lin)])
                if fine:                                   # hz_lin + h<L>_lin gradients are final (both on this stream)
                    on_stage(scope + "/dense")
            for (r0, nr, acc) in groups:
                self._gemm(False, True, nr, flat, F, g_h4[r0:r0 + nr].data_ptr(), F,
                           Wb["%s/%s/Matrix" % (scope, self.lin)].data_ptr(), F, g_hs[-1][r0:r0 + nr].data_ptr(), flat,
                           act_ref=hs[-1][r0:r0 + nr].data_ptr(), accumulate=acc, bwd=True)
            if on_dense is not None:
                on_dense(scope)
            gp, bias_done = None, False                    # (plane, scale) of g_hs[j] when conv j runs on planes
            for j in reversed(range(Ld)):
                wname = "%s/h%d_conv/w" % (scope, j)
                src = hs[j - 1] if j > 0 else x
                kj, Cin = sp.kernels[j], src.shape[3]
                if j == 0 and self._enc_few[0]:            # the image: filter gradient with the patch matrix built on chip
                    if gp is None:
                        gp = to_planes(g_hs[0], g_hs[0].shape[3], "gp_%s_h0" % tag, db=G["%s/h0_conv/biases" % scope])
                    with self._side():
                        ops.conv2d_few_wgrad(src, gp[0], kj, sp.strides[0], dw=G[wname], r_scale=gp[1])
                    continue
                if self._enc_p1[j]:                        # (one row group: separate context encoder)
                    if gp is None:
                        gp = to_planes(g_hs[j], g_hs[j].shape[3], "gp_%s_h%d" % (tag, j),
                                       db=G["%s/h%d_conv/biases" % (scope, j)])
                    gp5 = gp[0].view((1,) + tuple(g_hs[j].shape))
                    with self._side():                     # filter gradient: big = the conv's input (hi plane), small = dY
                        ops.conv2d_wgrad_planes(self._in_planes[(tag, j)][:1], gp5, kj, sp.strides[j], dw=G[wname],
                                                nprod=1, a_scale=gp[1])
                        if fine and j == Ld - 1:
                            on_stage(scope + "/top")
                    if j == 0:
                        continue
                    slot = new_slot()
                    acc = groups[0][2]
                    if self._enc_p1_merged[j]:
                        ops.conv2d_transpose_merged_planes(gp5, self._wb[wname], self.sizes[j], kj, Cin,
                                                           stride=sp.strides[j], out=g_hs[j - 1], nprod=1,
                                                           act_ref=hs[j - 1], accumulate=acc, a_scale=gp[1], amax=slot[0])
                    else:
                        ops.conv2d_transpose_planes(gp5, self._wb[wname], self.sizes[j], kj, stride=sp.strides[j],
                                                    out=g_hs[j - 1], nprod=1, act_ref=hs[j - 1], accumulate=acc,
                                                    a_scale=gp[1], amax=slot[0])
                    if self._enc_p1[j - 1] or (j == 1 and self._enc_few[0]):
                        gp = to_planes(g_hs[j - 1], Cin, "gp_%s_h%d" % (tag, j - 1), slot=slot,
                                       db=G["%s/h%d_conv/biases" % (scope, j - 1)])
                        bias_done = True
                    else:
                        gp, bias_done = None, False
                    continue
                with self._side():                         # filter + bias gradient beside the data gradient below
                    if self._few(Cin, kj) and g_hs[j].shape[3] % 4 == 0:
                        K = kj * kj * Cin
                        pt = self._patches("patch_%s_h%d_%d" % (scope, j, n), src, kj, sp.strides[j], reuse=True)
                        ops.gemm(pt[:, :K], g_hs[j].view(-1, g_hs[j].shape[3]), out=G[wname].view(K, -1), transA=True,
                                 accumulate=True, impl=self.impl_b)
                    else:
                        ops.conv2d_wgrad(src, g_hs[j], kj, sp.strides[j], dw=G[wname], impl=self.impl_b)
                    if not bias_done:
                        ops.bias_grad(g_hs[j], G["%s/h%d_conv/biases" % (scope, j)])
                    if fine and j == Ld - 1:
                        on_stage(scope + "/top")
                bias_done = False
                if j > 0:
                    for (r0, nr, acc) in groups:
                        ops.conv2d_transpose(g_hs[j][r0:r0 + nr], Wb[wname], self.sizes[j], stride=sp.strides[j],
                                             out=g_hs[j - 1][r0:r0 + nr], act_ref=hs[j - 1][r0:r0 + nr],
                                             accumulate=acc, round_tf32=self.round_b, impl=self.impl_b)

        if sp.separate_ctx_encoder and ctx_par:
            main_side = self._s2
            with torch.cuda.stream(self._s4):
                self._s2 = self._s5                        # (_side / _join of the enclosed calls use the chain's own side stream)
                try:
                    hz_backward(self.ctx_scope, g_cat.data_ptr() + F * f4, 2 * F, c_h4, gc_h4)
                    encoder_backward("conv_context", xc, c_hs, c_h4, gc_hs, gc_h4, [(0, B, True)], "c")
                    self._join()                           # s4 waits for the chain's filter gradients
                finally:
                    self._s2 = main_side
            encoder_backward("conv", xe, e_hs, e_h4, ge_hs, ge_h4, [(0, 2 * B, False)], "e")
            if on_stage:
                self._join()
                on_stage("conv")
            cur.wait_stream(self._s4)
            if on_stage:
                on_stage("conv_context")
        elif sp.separate_ctx_encoder:
            encoder_backward("conv", xe, e_hs, e_h4, ge_hs, ge_h4, [(0, 2 * B, False)], "e")
            if on_stage:
                self._join()
                on_stage("conv")
            encoder_backward("conv_context", xc, c_hs, c_h4, gc_hs, gc_h4, [(0, B, True)], "c")
            if on_stage:
                self._join()
                on_stage("conv_context")
        else:
            encoder_backward("conv", xe, e_hs, e_h4, ge_hs, ge_h4, [(0, 2 * B, False), (2 * B, B, True)], "e")
            if on_stage:
                self._join()
                on_stage("conv")
        self._join()                                       # gradients complete on the caller's stream

    # ---- optimiser -------------------------------------------------------------------------------------------------------
    def adam_step(self, lr: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
        """tf.train.AdamOptimizer (train_script.py:128), one fused launch over the flat buffers."""
        self._ensure_moments()
        if self._capturing:                                # graph replays: the step-dependent rate lives on the device
            ops.adam_schedule(self._step_dev, lr, b1, b2, self._lr_t_dev)
            lr_t = self._lr_t_dev
        else:
            self.step_count += 1
            self._eager_steps += 1
            t = self.step_count
            lr_t = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
        ops.adam_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, lr_t, b1, b2, eps, p_tf32=self.flat_pr)
        self._wp_stale = True

    def attach_allreduce(self, group=None):
        """Data parallelism: after this, every optimiser step sum-all-reduces the flat gradient (bucketed, launched
        from backward's stage callbacks) before Adam.  Build the net with global_batch = world * batch."""
        from . import parallel
        self._ensure_grads()
        # CTX_DP_FINE=1 also splits each encoder scope (dense layers / top conv / rest); measured on 4 GPUs the extra
        # collectives cost more than the shorter tail saves (6.76 vs 6.69 ms per step), so three buckets are the default
        self.allreduce = parallel.BucketedAllReduce(self.flat_g, parallel.scope_ranges(self.offsets, self.flat_size),
                                                    splits=(parallel.fine_splits(self.offsets, self.L)
                                                            if os.environ.get("CTX_DP_FINE", "0") == "1" else None),
                                                    group=group)
        return self.allreduce

    def apply_gradients(self, lr: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
        """backward (+ bucketed gradient all-reduce when data-parallel) + Adam; forward(FWD_ALL) must have run."""
        if self.allreduce is not None and getattr(self.allreduce, "enabled", False) and ops.PROFILE is None:
            # Data parallel: buckets are all-reduced as backward completes them; the Adam slice of a bucket runs on the
            # optimiser stream as soon as its reduction is done, under the rest of backward / the later reductions.
            ar = self.allreduce
            self._ensure_moments()
            if self._capturing:                            # graph replays: the step-dependent rate lives on the device
                ops.adam_schedule(self._step_dev, lr, b1, b2, self._lr_t_dev)
                lr_t = self._lr_t_dev
            else:
                self.step_count += 1
                self._eager_steps += 1
                t = self.step_count
                lr_t = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
            if self._s3 is None:
                self._s3 = torch.cuda.Stream(device=self.device)
            cur = torch.cuda.current_stream(self.device)

            from . import parallel
            scopes = parallel.scope_ranges(self.offsets, self.flat_size)

            def adam_range(a, b):                          # runs with the optimiser stream current, after the bucket
                # every kernel that READS these parameters (the data gradients of their layers) is already enqueued — on
                # the main stream, or on the context encoder's chain — when a later bucket is launched: the update must not
                # overtake it
                self._s3.wait_stream(cur)
                if self._s4 is not None:
                    self._s3.wait_stream(self._s4)
                ops.adam_step(self.flat_p[a:b], self.flat_g[a:b], self.flat_m[a:b], self.flat_v[a:b], lr_t, b1, b2, eps,
                              p_tf32=None if self.flat_pr is None else self.flat_pr[a:b])
                # next step's operand planes of every scope this bucket completes, under the remaining backward /
                # reductions (as the single-GPU branch does) instead of serially at the start of the next forward
                for sc, sa, sb in scopes:
                    if a < sb <= b:                        # the bucket holds the END of the scope: its last part
                        self._refresh_weight_planes(sc)

            ar.on_reduced, ar.reduced_stream = adam_range, self._s3
            self.backward(on_stage=ar.stage_done)
            ar.finish()
            cur.wait_stream(self._s3)
            self._wp_stale = False
            return
        if self.allreduce is not None:
            self.backward(on_stage=self.allreduce.stage_done)
            self.allreduce.finish()
            self.adam_step(lr, b1, b2, eps)
            return
        if not self.concurrent or ops.PROFILE is not None:
            self.backward()
            self.adam_step(lr, b1, b2, eps)
            return
        # Single process: a parameter scope is final (and its weights no longer read by this step) as soon as backward
        # leaves it, so its slice of the HBM-bound Adam pass runs on a third stream under the remaining backward GEMMs.
        from . import parallel
        self._ensure_moments()
        if self._capturing:                                # graph replays: the step-dependent rate lives on the device
            ops.adam_schedule(self._step_dev, lr, b1, b2, self._lr_t_dev)
            lr_t = self._lr_t_dev
        else:
            self.step_count += 1
            self._eager_steps += 1
            t = self.step_count
            lr_t = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
        ranges = {sc: (a, b) for sc, a, b in parallel.scope_ranges(self.offsets, self.flat_size)}
        if self._s3 is None:
            self._s3 = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream(self.device)

        dense_done: Dict[str, int] = {}                    # scope -> end of the part adam_dense already updated

        def adam_range(a, b):
            ops.adam_step(self.flat_p[a:b], self.flat_g[a:b], self.flat_m[a:b], self.flat_v[a:b], lr_t, b1, b2, eps,
                          p_tf32=None if self.flat_pr is None else self.flat_pr[a:b])

        def adam_scope(scope):                             # called after backward's join: gradients of `scope` are final
            a, b = ranges[scope]
            a = dense_done.get(scope, a)
            self._s3.wait_stream(cur)
            with torch.cuda.stream(self._s3):
                adam_range(a, b)
                self._refresh_weight_planes(scope)         # next step's forward operands, under the remaining backward

        def adam_dense(scope):
            # (CTX_ADAM_DENSE_EARLY=1; measured neutral at B = 256 and B = 32, so off by default) the dense layers hold two
            # thirds of an encoder scope and are done first: their slice runs under the scope's conv backward instead of after it (the weights' last readers — the hz_lin / h<L>_lin data gradients — are on the
            # current stream, their filter gradients on the current side stream)
            top = "%s/h%d_conv/w" % (scope, self.L - 1)
            if scope not in ranges or top not in self.offsets:
                return
            a, b = ranges[scope][0], self.offsets[top]
            if not a < b <= ranges[scope][1]:
                return
            self._s3.wait_stream(torch.cuda.current_stream(self.device))
            if self._s2 is not None:
                self._s3.wait_stream(self._s2)
            with torch.cuda.stream(self._s3):
                adam_range(a, b)
            dense_done[scope] = b

        self.backward(on_stage=adam_scope, on_dense=adam_dense if os.environ.get("CTX_ADAM_DENSE_EARLY", "0") == "1" else None)
        cur.wait_stream(self._s3)
        self._wp_stale = False

    def release_graphs(self):
        """Drop the captured step graphs (the next train_step captures again after two eager steps)."""
        self._graphs.clear()
        self._graph_meta.clear()
        self._eager_steps = 0

    def train_step(self, lr: float = 1e-4, allreduce=None):
        """forward + backward + (optional bucketed gradient all-reduce) + Adam: one `sess.run(optimizer, ...)`.
        `allreduce` is a parallel.BucketedAllReduce (or anything with stage_done(scope) / finish())."""
        if allreduce is not None:
            self.allreduce = allreduce
        if self.step_graphed(lr):
            return
        self.zero_grad()
        self.forward(FWD_ALL)
        self.apply_gradients(lr)

    def step_graphed(self, lr: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> bool:
        """zero_grad + forward(FWD_ALL) + apply_gradients on the current input as one CUDA-graph launch.  Returns False when
        the step has to run eagerly (first two steps, profiling, CTX_GRAPH=0 / CTX_DP_GRAPH=0, capture not possible)."""
        dp = self.allreduce is not None and getattr(self.allreduce, "enabled", False)
        if (not self._graph_on or self._eager_steps < 2 or not self.concurrent or ops.PROFILE is not None or
                (dp and os.environ.get("CTX_DP_GRAPH", "1") == "0")):
            return False
        if self._in_fresh:
            # a host batch whose copies were issued just now (no prefetch): the eager step overlaps them with the context
            # encoder role by role; a graph would have to wait for the whole batch first
            return False
        cur = torch.cuda.current_stream(self.device)
        if self._in_events is not None:                    # staged host input: the graph starts after the copies
            cur.wait_event(self._in_events[0])
            cur.wait_event(self._in_events[1])
        key = (self._in_slot, float(lr), float(b1), float(b2), float(eps))
        g = self._graphs.get(key)
        if g is None:
            if self._step_dev is None:
                self._step_dev = torch.zeros(1, device=self.device, dtype=torch.int64)
                self._lr_t_dev = torch.zeros(1, device=self.device, dtype=torch.float32)
            try:
                g = torch.cuda.CUDAGraph()
                self._capturing = True
                l0 = ops.launch_count()
                # data parallel: the bucketed NCCL all-reduces are captured with the step (NCCL's watchdog thread polls
                # events, which only a thread-local capture tolerates)
                with torch.cuda.graph(g, capture_error_mode="thread_local" if dp else "global"):
                    self.zero_grad()
                    self.forward(FWD_ALL)
                    self.apply_gradients(lr, b1, b2, eps)
                # kernels of this library one replay runs; whether the step leaves the weight planes stale
                self._graph_meta[key] = (ops.launch_count() - l0, self._wp_stale)
            except Exception as ex:                        # noqa: BLE001  (anything not capturable: stay eager, say so once)
                self._graph_on = False
                self._capturing = False
                torch.cuda.synchronize(self.device)
                import traceback
                import warnings
                tb = traceback.format_exc().strip().splitlines()
                warnings.warn("CUDA-graph capture of the training step failed (%s | %s | library: %s); running eagerly"
                              % (str(ex)[:200], " <- ".join(t.strip() for t in tb[-6:-1])[:600], _lib.last_error()[:200]))
                return False
            finally:
                self._capturing = False
            self._graphs[key] = g
        if self._step_dev_host != self.step_count:         # eager steps ran in between: realign the device counter
            self._step_dev.fill_(self.step_count)
        g.replay()
        launches, stale = self._graph_meta[key]
        ops.count_replayed(launches)
        self.step_count += 1
        self._step_dev_host = self.step_count
        self._level = FWD_ALL
        self._wp_stale = stale
        return True

    # ---- fetches -----------------------------------------------------------------------------------------------------------
    def fetch(self, name: str) -> torch.Tensor:
        """Device tensor for a reference attribute name (arm_shaping.py:1296,1312,1349-1354)."""
        B, F, bufs = self.B, self.F, self._bufs
        need = {"input_z": FWD_SRC, "translated_z": FWD_TRANS, "out": FWD_TRANS, "ctx_z": FWD_TRANS,
                "image_trans": 0}.get(name, FWD_ALL)
        if self._level < need:
            raise _lib.CtxError("fetch(%r) needs forward(level>=%d)" % (name, need))
        if name == "input_z":
            return bufs["cat"][:, :F]
        if name == "ctx_z":
            return bufs["cat"][:, F:]
        if name == "translated_z":
            return bufs["z2"][:B]
        if name == "target_z":
            return bufs["z2"][B:]
        if name in ("out", "out2"):
            d4 = bufs["d%d" % self.L]
            rows = d4[:B] if name == "out" else d4[B:]
            if self.spec.residual_out:
                _, xc = self._input_buffers()
                return ops.add_bcast(rows.contiguous(), xc.contiguous())
            return rows
        if name == "recon1":
            return bufs["losses"][0]
        if name == "recon2":
            return bufs["losses"][1]
        if name == "simloss":
            return bufs["losses"][2]
        if name == "loss":
            l = bufs["losses"]
            loss = self._w_recon[0] * l[0] + self._w_recon[1] * l[1] + self._w_sim * l[2]
            if self.bf3 and self.pdt == torch.float16:
                # an activation beyond fp16's range was clipped in its operand planes: not a loss worth reporting
                loss = torch.where(l[4:5].view(torch.int32)[0] >= 0x477FE000, torch.full_like(loss, float("nan")), loss)
            return loss
        if name == "image_trans":
            xe, xc = self._input_buffers()
            return torch.stack([xe[:B], xc, xe[B:2 * B]])
        raise KeyError(name)

    def nn_err(self, nlen: int) -> int:
        """scripts/train_script.py:148 on the current batch (integer)."""
        xe, _ = self._input_buffers()
        err, _ = ops.nn_err(xe[self.B:2 * self.B].contiguous(), self.fetch("out").contiguous(), nlen)
        return int(err.item())
