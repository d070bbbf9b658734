"""world_size-2 CPU test (gloo) of the multi-GPU host logic: view sharding and the gradient all-reduce that a
data-parallel loop over the replicated rasterizer needs (gaussianeditor_b200/distributed.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaussianeditor_b200 import distributed as D


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        views = D.view_shard(7, world, rank)
        g = torch.Generator().manual_seed(100)            # same "cloud" on every rank
        base = [torch.randn(50, 3, generator=g), torch.randn(50, 16, 3, generator=g), torch.randn(50, 1, generator=g)]
        # each rank's gradient = sum over its own views of (view+1) * base  -> the all-reduce must give sum over all views
        grads = [sum((v + 1) for v in views) * b.clone() for b in base]
        D.allreduce_gradients(grads)
        want = [sum(v + 1 for v in range(7)) * b for b in base]
        ok = all(torch.allclose(a, b, rtol=1e-6, atol=1e-6) for a, b in zip(grads, want))
        t = D.max_over_ranks(1.0 + rank)
        out[rank] = (ok, views, t)
    finally:
        dist.destroy_process_group()


def test_view_shard_partition():
    for n in (1, 7, 8, 48):
        for w in (1, 2, 4, 8):
            allv = sorted(v for r in range(w) for v in D.view_shard(n, w, r))
            assert allv == list(range(n))
            sizes = [len(D.view_shard(n, w, r)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == D.steps_per_rank(n, w)


def test_gloo_world2_allreduce_and_timing():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        ok, views, t = out[r]
        assert ok and views == list(range(r, 7, world)) and t == 2.0
