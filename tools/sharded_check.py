"""One process per GPU (torchrun): the Gaussian-sharded rasterizer must reproduce the single-GPU rasterizer.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_check.py [--config c2 --P N]

Every rank renders the full cloud alone (single-GPU path) and then its shard through ShardedGaussianRasterizer;
images and radii must be bit-identical, gradients within 2e-5 relative L2. Prints SHARDED_CHECK_OK on rank 0.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

from gaussianeditor_b200 import sharded as S, synth  # noqa: E402
from gaussianeditor_b200.rasterizer import GaussianRasterizer  # noqa: E402
from util import cloud_tensors, rel_l2, settings_from  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--P", type=int, default=None)
    ap.add_argument("--p2p", action="store_true", help="dense mode: fused preprocess + all-gather over peer memory")
    ap.add_argument("--mode", default="dense", choices=["dense", "sparse"])
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    S.init_distributed("nccl", dev)
    cloud, cams = synth.make_config(a.config, P=a.P)
    cam = cams[0]
    bg = (0.2, 0.5, 0.1)
    H, W = cam.image_height, cam.image_width
    P = cloud.means3D.shape[0]
    rs = settings_from(cam, bg, cloud.sh_degree, dev)
    dL = torch.from_numpy(np.random.default_rng(3).random((3, H, W), dtype=np.float32)).to(dev)

    full = cloud_tensors(cloud, dev, requires_grad=True)
    m2 = torch.zeros_like(full["means3D"], requires_grad=True)
    color, radii, depth = GaussianRasterizer(rs)(means3D=full["means3D"], means2D=m2, opacities=full["opacities"],
                                                shs=full["shs"], scales=full["scales"], rotations=full["rotations"])
    (color * dL).sum().backward()

    rast = S.ShardedGaussianRasterizer(rs, P, p2p=(True if a.mode == 'sparse' else a.p2p), mode=a.mode)
    plan = rast.plan
    loc = {k: S.shard_slice(v.detach(), plan).clone().requires_grad_(True) for k, v in full.items()}
    lm2 = torch.zeros_like(loc["means3D"], requires_grad=True)
    scolor, sradii, sdepth = rast(means3D=loc["means3D"], means2D=lm2, opacities=loc["opacities"], shs=loc["shs"],
                                  scales=loc["scales"], rotations=loc["rotations"])
    (scolor * dL).sum().backward()
    torch.cuda.synchronize()

    # GaussianEditor's step shape: two forwards alive at once (SH render + mask render with precomputed colours), then
    # ONE backward through the first -- exercises the workspace pool with two workspaces in flight
    torch.manual_seed(7)
    mask_col = torch.rand(plan.count, 3, device=dev)
    full_col = torch.zeros(P, 3, device=dev)
    full_col[plan.base:plan.base + plan.count] = mask_col
    dist.all_reduce(full_col)
    for v in list(loc.values()) + [lm2]:
        v.grad = None
    c1, _, _ = rast(means3D=loc["means3D"], means2D=lm2, opacities=loc["opacities"], shs=loc["shs"],
                    scales=loc["scales"], rotations=loc["rotations"])
    c2, _, _ = rast(means3D=loc["means3D"].detach(), means2D=torch.zeros_like(lm2), opacities=loc["opacities"].detach(),
                    colors_precomp=mask_col, scales=loc["scales"].detach(), rotations=loc["rotations"].detach())
    (c1 * dL).sum().backward()
    ref_c2, _, _ = GaussianRasterizer(rs)(means3D=full["means3D"].detach(), means2D=torch.zeros_like(m2),
                                          opacities=full["opacities"].detach(), colors_precomp=full_col,
                                          scales=full["scales"].detach(), rotations=full["rotations"].detach())
    two_ok = torch.equal(c1, color) and torch.equal(c2, ref_c2)
    del c1, c2
    if a.p2p or a.mode == 'sparse':  # second round through the recycled peer workspace
        for v in list(loc.values()) + [lm2]:
            v.grad = None
        del scolor, sradii, sdepth
        scolor, sradii, sdepth = rast(means3D=loc["means3D"], means2D=lm2, opacities=loc["opacities"], shs=loc["shs"],
                                      scales=loc["scales"], rotations=loc["rotations"])
        (scolor * dL).sum().backward()
        torch.cuda.synchronize()
    ok = two_ok and torch.equal(scolor, color) and torch.equal(sdepth, depth) and torch.equal(sradii, S.shard_slice(radii, plan))
    worst = 0.0
    pairs = [(loc[k].grad, S.shard_slice(full[k].grad, plan)) for k in loc] + [(lm2.grad, S.shard_slice(m2.grad, plan))]
    for g, w in pairs:
        worst = max(worst, rel_l2(g.cpu().numpy(), w.cpu().numpy()))
    ok = ok and worst <= 2e-5
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        extra = ""
        if a.mode == "sparse":
            from gaussianeditor_b200 import sparse_sharded as SS
            extra = f" sparse={SS._SparseShardedRasterize.last} redo={rast.sparse_pool.redo}"
        else:
            extra = f" R_rank0={S._ShardedRasterize.last_R}"
        print(f"world={world} mode={a.mode} p2p={a.p2p} P={P}{extra} worst_grad_rel_l2={worst:.2e}")
        print("SHARDED_CHECK_OK" if int(flag) == 1 else "SHARDED_CHECK_FAILED")
    del scolor, sradii, sdepth
    import gc
    gc.collect()
    rast.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
