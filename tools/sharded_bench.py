"""Config 4 (5M Gaussians, SH degree 3, 1920x1080) through the Gaussian-sharded path, one process per GPU.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_bench.py \
      [--config c4] [--P N] [--steps 10] [--warmup 3] [--single]

Per step: forward (preprocess shard -> all-gather -> order -> render owned tiles -> all-reduce) + backward (render
backward -> reduce-scatter -> preprocess backward) for one camera of the ring (cycled). Timed with CUDA events on the
compute stream, max over ranks. Also prints a per-phase breakdown (events between the phases, mean over steps) and,
with --single, the single-GPU rasterizer on the full cloud on rank 0 for the strong-scaling ratio.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gaussianeditor_b200 import sharded as S, synth  # noqa: E402
from gaussianeditor_b200.rasterizer import GaussianRasterizer  # noqa: E402
from util import cloud_tensors, settings_from  # noqa: E402

PHASES = ["preprocess", "all_gather", "order", "render", "all_reduce", "render_bwd", "reduce_scatter", "preprocess_bwd"]


def sharded_step(plan, ex, rs, loc, dL, ev=None, ws=None):
    """One forward+backward through the step functions (what _ShardedRasterize does), with optional phase events."""
    mark = (lambda: ev.append(torch.cuda.Event(enable_timing=True)) or ev[-1].record()) if ev is not None else (lambda: None)
    empty = torch.empty(0, device=dL.device)
    mark()
    buf = S.shard_preprocess(plan, rs, loc["means3D"], loc["shs"], empty, loc["opacities"], loc["scales"],
                             loc["rotations"], empty, ws=ws)
    mark()
    if buf.peer is not None:
        ex.barrier(dL.device)
    else:
        ex.all_gather_inplace(S.exchange_view(buf))
    mark()
    S.shard_order(buf)
    mark()
    frame = torch.zeros(4, buf.H, buf.W, dtype=torch.float32, device=dL.device)
    S.shard_render(buf, frame[:3], frame[3:])
    mark()
    ex.all_reduce_sum(frame)
    mark()
    acc = S.shard_backward_render(buf, dL)   # loss = (color * dL).sum()  ->  dL/dcolor = dL
    mark()
    acc_slice = buf.alloc("acc_slice", plan.slice_len * S.ACC_STRIDE, torch.float32, dL.device).view(plan.slice_len, S.ACC_STRIDE)
    ex.reduce_scatter_sum(acc, acc_slice)
    mark()
    grads = S.shard_backward_preprocess(buf, acc_slice)
    mark()
    return frame, grads, buf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4")
    ap.add_argument("--P", type=int, default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--single", action="store_true")
    ap.add_argument("--p2p", action="store_true", help="fused preprocess + all-gather over peer memory")
    ap.add_argument("--no-pool", action="store_true", help="allocate every scratch buffer per step (diagnostic)")
    ap.add_argument("--sync-each", action="store_true", help="host-synchronise after every timed step (diagnostic)")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    S.init_distributed("nccl", dev)
    cloud, cams = synth.make_config_cached(a.config, P=a.P)
    P = cloud.means3D.shape[0]
    H, W = cams[0].image_height, cams[0].image_width
    bg = (0.0, 0.0, 0.0)
    settings = [settings_from(c, bg, cloud.sh_degree, dev) for c in cams]
    dL = torch.from_numpy(np.random.default_rng(7).random((3, H, W), dtype=np.float32)).to(dev)
    ex = S.Exchange()
    plan = S.ShardPlan(P, world, rank)
    full = cloud_tensors(cloud, dev)
    loc = {k: S.shard_slice(v, plan).clone() for k, v in full.items()}
    if not (a.single and rank == 0):
        del full
    torch.cuda.empty_cache()

    from gaussianeditor_b200.rasterizer import _geometry_bytes
    from gaussianeditor_b200 import _lib
    pool = S.WorkspacePool(ex, _geometry_bytes(_lib.load(), plan.P_pad), a.p2p) if not a.no_pool else None
    ws = pool.take(dev) if pool is not None else None   # one step in flight at a time: one workspace
    for i in range(a.warmup):
        sharded_step(plan, ex, settings[i % len(settings)], loc, dL, ws=ws)
    torch.cuda.synchronize(); dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ndev0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    t0.record()
    Rs = []
    for i in range(a.steps):
        _, _, buf = sharded_step(plan, ex, settings[i % len(settings)], loc, dL, ws=ws)
        Rs.append(buf.R)
        if a.sync_each:
            torch.cuda.synchronize()
    t1.record()
    torch.cuda.synchronize()
    ndev1 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    ms = torch.tensor([t0.elapsed_time(t1) / a.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)

    # phase breakdown (separate pass so the event records do not perturb the timed region)
    acc_ms = np.zeros(len(PHASES))
    for i in range(a.steps):
        ev = []
        sharded_step(plan, ex, settings[i % len(settings)], loc, dL, ev, ws=ws)
        torch.cuda.synchronize()
        acc_ms += np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(len(PHASES))])
    phase = torch.tensor(acc_ms / a.steps, device=dev)
    phase_max = phase.clone()
    dist.all_reduce(phase_max, op=dist.ReduceOp.MAX)

    single_ms = None
    if a.single and rank == 0:
        rast = [GaussianRasterizer(s) for s in settings]
        t = {k: v.clone().requires_grad_(True) for k, v in full.items()}
        m2 = torch.zeros_like(t["means3D"], requires_grad=True)

        def one(i):
            color, _, _ = rast[i % len(rast)](means3D=t["means3D"], means2D=m2, opacities=t["opacities"], shs=t["shs"],
                                              scales=t["scales"], rotations=t["rotations"])
            (color * dL).sum().backward()
            for v in list(t.values()) + [m2]:
                v.grad = None
        for i in range(a.warmup):
            one(i)
        torch.cuda.synchronize()
        t0.record()
        for i in range(a.steps):
            one(i)
        t1.record()
        torch.cuda.synchronize()
        single_ms = t0.elapsed_time(t1) / a.steps
    dist.barrier()
    if rank == 0:
        m = float(ms)
        print(json.dumps({
            "metric": "forward+backward Mpixels/s, Gaussian-sharded", "value": W * H / (m * 1e-3) / 1e6, "unit": "Mpixels/s",
            "n_gpus": world, "ms_per_step": m, "scaling": "strong",
            "config": {"workload": f"{a.config}: P={P}, SH degree {cloud.sh_degree}, {W}x{H}, {len(cams)} ring cameras cycled",
                       "parallelism": f"gaussian-shard x{world} + tile-row interleave",
                       "exchange": "fused peer stores (TMA) + barrier" if a.p2p else "NCCL all-gather"},
            "R_rank0_mean": float(np.mean(Rs)),
            "phase_ms_max_over_ranks": {n: round(float(v), 4) for n, v in zip(PHASES, phase_max.tolist())},
            "single_gpu_ms_per_step": single_ms, "cudaMallocs_in_timed_region": ndev1 - ndev0,
            "reserved_GB": round(torch.cuda.memory_reserved(dev) / 2**30, 2),
            "exchange_bytes_per_step": {"all_gather": plan.P_pad * 48, "all_reduce": 16 * W * H,
                                        "reduce_scatter": plan.P_pad * 48},
        }))
    if pool is not None:
        del buf
        pool.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
