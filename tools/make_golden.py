"""Generate the committed golden vectors under tests/golden/ by running the REFERENCE's own CUDA build
(oracle/_ref/libdgr_ref.so = the unmodified reference .cu files, see oracle/Makefile) on small seeded scenes.
Must run on a GPU box:   gpurun -- python tools/make_golden.py   (writes gpurun_out/golden/*.npz; copy them to
tests/golden/ and commit).  The reference has no tests or fixtures of its own; these files
are what pins the CPU oracle and our kernels to the reference's actual outputs when no GPU reference is at hand."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

from gaussianeditor_b200 import synth
import test_parity_gpu as T

OUT = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)


def case(name, cloud, cam, bg, colors_precomp=None, scale_modifier=1.0, seed=0):
    dL = np.random.default_rng(seed).uniform(size=(3, cam.image_height, cam.image_width)).astype(np.float32)
    ref = T._ref_run(cloud, cam, bg, dL=dL, colors_precomp=colors_precomp, scale_modifier=scale_modifier)
    ref2 = T._ref_run(cloud, cam, bg, dL=dL, colors_precomp=colors_precomp, scale_modifier=scale_modifier)
    s = ref["state"]
    g = ref["grads"]
    noise = {k: float((g[k] - ref2["grads"][k]).norm() / (g[k].norm() + 1e-30)) for k in g}
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        # inputs
        means3D=cloud.means3D, scales=cloud.scales, rotations=cloud.rotations, opacities=cloud.opacities,
        shs=cloud.shs, sh_degree=cloud.sh_degree, colors_precomp=np.zeros((0, 3), np.float32) if colors_precomp is None else colors_precomp,
        viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, campos=cam.campos, tanfovx=cam.tanfovx,
        tanfovy=cam.tanfovy, H=cam.image_height, W=cam.image_width, bg=np.asarray(bg, np.float32),
        scale_modifier=scale_modifier, dL=dL,
        # reference outputs
        color=ref["color"].cpu().numpy(), depth=ref["depth"].cpu().numpy(), radii=ref["radii"].cpu().numpy(),
        R=ref["R"], final_T=s["final_T"].cpu().numpy(), n_contrib=s["n_contrib"].cpu().numpy(),
        ranges=s["ranges"].cpu().numpy(), point_list=s["point_list"].cpu().numpy(),
        tiles_touched=s["tiles_touched"].cpu().numpy(), means2D=s["means2D"].cpu().numpy(),
        conic_opacity=s["conic_opacity"].cpu().numpy(), rgb=s["rgb"].cpu().numpy(), depths=s["depths"].cpu().numpy(),
        **{k: v.cpu().numpy() for k, v in g.items()},
        noise=np.array([noise[k] for k in sorted(noise)]), noise_keys=np.array(sorted(noise)))
    print(name, "R", ref["R"], "visible", int((ref["radii"] > 0).sum()), "noise", max(noise.values()))


c3, _ = synth.make_config("c3", P=400)
case("c3_p400_deg3", c3, synth.ring_cameras(8, 4.5, 15.0, 72, 56, 61.0)[2], (0.1, 0.3, 0.6), seed=1)
c1, _ = synth.make_config("c1", P=300)
case("c1_p300_deg0", c1, synth.look_at_camera((0, 0, -3.5), (0, 0, 0), (0, -1, 0), 64, 48, fovy_deg=50.0), (0, 0, 0), seed=2)
c2, _ = synth.make_config("c2", P=300)
cp = np.random.default_rng(5).uniform(size=(300, 3)).astype(np.float32)
case("c2_p300_colors_mod07", c2, synth.look_at_camera((0, 0, -3.5), (0, 0, 0), (0, -1, 0), 50, 35, fovy_deg=50.0),
     (1, 1, 1), colors_precomp=cp, scale_modifier=0.7, seed=3)
print("golden written to", OUT)
