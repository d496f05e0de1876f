"""World-size-2 gloo tests (CPU) of the data-parallel logic: bucketed all-reduce == one all-reduce of the whole
flat gradient, and per-rank gradients with the *global* simloss denominator sum to the full-batch gradient
 — checked with the oracle standing in for the per-rank compute."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from imitation_from_observation_b200 import parallel, spec as S


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _offsets(spec):
    off, o = {}, 0
    for n, shp in spec.param_shapes().items():
        off[n] = o
        o += -(-int(np.prod(shp)) // 64) * 64
    return off, o


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_distributed("gloo")
    assert (r, w) == (rank, world)
    from oracle import gen_golden, tf_semantics as O
    spec = S.context_skip_new(16, 16, 3, df_dim=4, featsize=16)
    off, total = _offsets(spec)
    P = {k: v.double() for k, v in gen_golden.golden_params(spec, 11).items()}
    g = torch.Generator().manual_seed(5)
    Bg = 4
    image = torch.rand((3, Bg, 16, 16, 3), generator=g, dtype=torch.float64) * 2 - 1
    lo, hi = parallel.shard_range(Bg, rank, world)
    local = image[:, lo:hi]
    # per-rank loss with the GLOBAL mean for simloss: recon sums + 1e3 * sum(diff^2) / (Bg * F)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    res = O.forward(spec, Pg, local)
    sim_global = ((res["translated_z"] - res["target_z"]) ** 2).sum() * 1e3 / (Bg * spec.featsize)
    loss = res["recon1"] + res["recon2"] + sim_global
    grads = torch.autograd.grad(loss, list(Pg.values()))
    flat = torch.zeros(total, dtype=torch.float64)
    for (name, shp), gr in zip(spec.param_shapes().items(), grads):
        flat[off[name]:off[name] + gr.numel()] = gr.reshape(-1)
    whole = flat.clone()
    flat_local = flat.clone()
    dist.all_reduce(whole)
    ar = parallel.BucketedAllReduce(flat, parallel.scope_ranges(off, total))
    assert ar.enabled
    for scope in ("deconv", "translate", "conv", "conv_context"):
        ar.stage_done(scope)
    ar.finish()
    ok_bucket = bool(torch.equal(flat, whole))
    # finer buckets (dense layers / top conv / rest of each encoder) + the per-bucket "reduced" callback, in order
    flat2 = flat_local.clone()
    ar2 = parallel.BucketedAllReduce(flat2, parallel.scope_ranges(off, total), splits=parallel.fine_splits(off, 4))
    seen = []
    ar2.on_reduced = lambda a, b: seen.append((a, b))
    stages = ["deconv", "translate", "conv/dense", "conv/top", "conv", "conv_context/dense", "conv_context/top",
              "conv_context"]
    for st in stages:
        ar2.stage_done(st)
    ar2.finish()
    ok_bucket = ok_bucket and bool(torch.equal(flat2, whole)) and len(ar2.buckets) == 7
    ok_bucket = ok_bucket and seen == [(a, b) for _, a, b in ar2.buckets] and seen[0][0] == 0 and seen[-1][1] == total
    ok_bucket = ok_bucket and ar2.buckets[1][2] == off["conv/h3_conv/w"] and ar2.buckets[2][2] == off["conv/h2_conv/w"]
    # full-batch oracle gradient
    _, G = O.loss_and_grads(spec, P, image)
    err = 0.0
    for name in spec.param_shapes():
        a = flat[off[name]:off[name] + G[name].numel()].reshape(G[name].shape)
        err = max(err, float((a - G[name]).abs().max() / (G[name].abs().max() + 1e-30)))
    if rank == 0:
        q.put((ok_bucket, err))
    dist.barrier()
    # orderly end of a data-parallel process: graphs (none here) before the process group; idempotent, and what atexit runs again
    parallel.shutdown()
    assert not dist.is_initialized()
    parallel.shutdown()


@pytest.mark.timeout(180)
def test_bucketed_allreduce_equals_full_batch_gradient():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok_bucket, err = q.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok_bucket
    assert err < 1e-10
