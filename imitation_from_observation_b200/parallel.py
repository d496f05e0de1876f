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
    return rank, world, local


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


class BucketedAllReduce:
    """Sum all-reduce of a flat gradient in per-scope buckets, launched from backward's stage callbacks."""

    def __init__(self, flat_g: torch.Tensor, ranges: Sequence[Tuple[str, int, int]], group=None,
                 merge: Sequence[Sequence[str]] = (("deconv", "translate"), ("conv",), ("conv_context",))):
        self.flat_g, self.group = flat_g, group
        by = {sc: (a, b) for sc, a, b in ranges}
        self.buckets: List[Tuple[str, int, int]] = []
        for grp in merge:
            present = [s for s in grp if s in by]
            if not present:
                continue
            a, b = by[present[0]][0], by[present[-1]][1]
            self.buckets.append((present[-1], a, b))        # fires when the LAST scope of the bucket is done
        covered = sum(b - a for _, a, b in self.buckets)
        if covered != flat_g.numel():
            raise ValueError("buckets cover %d of %d gradient elements" % (covered, flat_g.numel()))
        self._work = []
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    def stage_done(self, scope: str):
        if not self.enabled:
            return
        for trigger, a, b in self.buckets:
            if trigger == scope:
                self._work.append(dist.all_reduce(self.flat_g[a:b], op=dist.ReduceOp.SUM, group=self.group,
                                                  async_op=True))

    def finish(self):
        for w in self._work:
            w.wait()
        self._work = []
