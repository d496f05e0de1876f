"""Data parallelism for the trainer: one process per GPU, batch elements (demo-video pairs) sharded over ranks,
one sum all-reduce of the flat fp32 gradient per step, issued as contiguous buckets as soon as
each parameter scope's gradients are complete so NCCL over NVLink overlaps the rest of backward.

The reference has no multi-GPU code at all; this module is new.  It only touches torch.distributed (NCCL on
GPUs, gloo in the CPU tests) and flat tensors, so its logic is testable without a GPU.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

SCOPE_ORDER = ("deconv", "translate", "conv", "conv_context")


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; initialises the default group if needed."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            # NCCL prints its version banner to stdout when the first communicator is created; callers such as
            # bench.py own stdout (one JSON line), so create the communicator with fd 1 pointed at stderr
            import sys
            sys.stdout.flush()
            saved = os.dup(1)
            try:
                os.dup2(2, 1)
                dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
                warm = torch.zeros(1, device=torch.device("cuda", local))
                dist.all_reduce(warm)
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    global _SHUTDOWN_HOOKED
    if world > 1 and not _SHUTDOWN_HOOKED:
        import atexit
        atexit.register(shutdown)
        _SHUTDOWN_HOOKED = True
    return rank, world, local


_SHUTDOWN_HOOKED = False


def shutdown():
    """Orderly end of a data-parallel process: captured step graphs (they hold NCCL kernel nodes) are destroyed first, the
    device drains, then the process group goes.  Idempotent; registered with atexit by init_distributed."""
    from . import network
    network.release_all_graphs()
    if dist.is_available() and dist.is_initialized():
        try:
            dist.destroy_process_group()
        except Exception:                                  # noqa: BLE001  (exit path: nothing useful to do with it)
            pass


def scope_ranges(offsets: Dict[str, int], flat_size: int) -> List[Tuple[str, int, int]]:
    """Contiguous [start, end) range of the flat buffer per parameter scope, in buffer order.  Relies on the
    parameter order of spec.NetSpec.param_shapes() (decoder, translator, encoder(s))."""
    names = list(offsets.keys())
    firsts: List[Tuple[str, int]] = []
    for n in names:
        sc = n.split("/")[0]
        if not firsts or firsts[-1][0] != sc:
            if any(f[0] == sc for f in firsts):
                raise ValueError("parameters of scope %r are not contiguous in the flat buffer" % sc)
            firsts.append((sc, offsets[n]))
    out = []
    for i, (sc, start) in enumerate(firsts):
        end = firsts[i + 1][1] if i + 1 < len(firsts) else flat_size
        out.append((sc, start, end))
    if out and out[0][1] != 0:
        raise ValueError("flat buffer does not start at offset 0")
    return out


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) shard of `n_items` (demo videos / trajectories) for `rank`."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def fine_splits(offsets: Dict[str, int], depth: int) -> Dict[str, List[Tuple[str, int]]]:
    """Split points inside each encoder scope, in buffer (= gradient completion) order: the dense layers (hz_lin,
    h<L>_lin: two thirds of the scope's parameters) are final long before the convs are, and the top conv holds most of
    the rest.  Returns scope -> [(stage name that completes the PRECEDING part, offset where the next part starts)]."""
    out: Dict[str, List[Tuple[str, int]]] = {}
    for sc in ("conv", "conv_context"):
        top, nxt = "%s/h%d_conv/w" % (sc, depth - 1), "%s/h%d_conv/w" % (sc, depth - 2)
        if top in offsets:
            pts = [(sc + "/dense", offsets[top])]
            if depth >= 2 and nxt in offsets:
                pts.append((sc + "/top", offsets[nxt]))
            out[sc] = pts
    return out


class BucketedAllReduce:
    """Sum all-reduce of a flat gradient in contiguous buckets, launched from backward's stage callbacks as soon as a
    bucket's gradients are final, so NCCL over NVLink runs under the rest of backward.

    Buckets: decoder + translator, then each encoder scope — optionally split at `splits` (fine_splits) into its dense
    layers, its top conv and the rest, which leaves only the small last part exposed after backward ends.
    `on_reduced(a, b)`, if set, is called once per bucket, in order, as soon as a LATER bucket is launched (and for the
    remaining ones in finish()): the caller runs the optimiser on [a, b) there; `reduced_stream` is the CUDA stream
    made to wait for the bucket before the callback runs (None: the current stream)."""

    fine = True        # backward() may deliver the finer "<scope>/dense", "<scope>/top" stages

    def __init__(self, flat_g: torch.Tensor, ranges: Sequence[Tuple[str, int, int]], group=None,
                 merge: Sequence[Sequence[str]] = (("deconv", "translate"), ("conv",), ("conv_context",)),
                 splits: Optional[Dict[str, List[Tuple[str, int]]]] = None):
        self.flat_g, self.group = flat_g, group
        by = {sc: (a, b) for sc, a, b in ranges}
        self.buckets: List[Tuple[str, int, int]] = []
        for grp in merge:
            present = [s for s in grp if s in by]
            if not present:
                continue
            a, b = by[present[0]][0], by[present[-1]][1]
            pts = (splits or {}).get(present[-1], []) if len(present) == 1 else []
            for stage, off in pts:
                if a < off < b:
                    self.buckets.append((stage, a, off))
                    a = off
            self.buckets.append((present[-1], a, b))        # fires when the LAST scope (part) of the bucket is done
        covered = sum(b - a for _, a, b in self.buckets)
        if covered != flat_g.numel():
            raise ValueError("buckets cover %d of %d gradient elements" % (covered, flat_g.numel()))
        self._work: List[tuple] = []                         # launched, not yet handed to on_reduced: (work, a, b)
        self.on_reduced = None
        self.reduced_stream = None
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    def _drain(self):
        """Hand every launched bucket to on_reduced, in order; the wait is a stream dependency, not a host block (NCCL)."""
        for w, a, b in self._work:
            if self.reduced_stream is not None:
                with torch.cuda.stream(self.reduced_stream):
                    w.wait()
                    if self.on_reduced is not None:
                        self.on_reduced(a, b)
            else:
                w.wait()
                if self.on_reduced is not None:
                    self.on_reduced(a, b)
        self._work = []

    def stage_done(self, scope: str):
        if not self.enabled:
            return
        for trigger, a, b in self.buckets:
            if trigger == scope:
                if self.on_reduced is not None:
                    self._drain()                            # earlier buckets: optimiser under this bucket's reduction
                self._work.append((dist.all_reduce(self.flat_g[a:b], op=dist.ReduceOp.SUM, group=self.group,
                                                   async_op=True), a, b))

    def finish(self):
        self._drain()
