"""what the per-render PyTorch prologue (sigmoid, exp, normalize, cat) and its autograd epilogue cost
around the rasterizer at BASELINE config 3, and what folding them into the preprocess kernels saves.

    python tools/fused_act_bench.py [--config c3] [--steps 30]

Both arms start from the scene model's raw leaf parameters and time forward + backward (CUDA events):
  prologue : activations in PyTorch -> GaussianRasterizer.forward -> autograd through the activations
  fused    : GaussianRasterizer.forward_raw
Prints one JSON line.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gaussianeditor_b200 import synth  # noqa: E402
from gaussianeditor_b200.rasterizer import GaussianRasterizer  # noqa: E402
from util import settings_from  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    dev = torch.device("cuda")
    cloud, cams = synth.make_config(a.config)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    op = t(cloud.opacities).clamp(1e-6, 1 - 1e-6)
    raw = dict(xyz=t(cloud.means3D), opacity=torch.log(op / (1 - op)), dc=t(cloud.shs[:, :1, :]).contiguous(),
               rest=t(cloud.shs[:, 1:, :]).contiguous(), scaling=torch.log(t(cloud.scales)), rotation=t(cloud.rotations) * 1.7)
    raw = {k: v.requires_grad_(True) for k, v in raw.items()}
    rasts = [GaussianRasterizer(settings_from(c, (0, 0, 0), cloud.sh_degree, dev)) for c in cams]
    H, W = cams[0].image_height, cams[0].image_width
    G = torch.rand(3, H, W, device=dev)

    def step(i, fused):
        for v in raw.values():
            v.grad = None
        m2 = torch.zeros_like(raw["xyz"], requires_grad=True)
        r = rasts[i % len(rasts)]
        if fused:
            color, _, _ = r.forward_raw(means3D=raw["xyz"], means2D=m2, opacity_logits=raw["opacity"], features_dc=raw["dc"],
                                        features_rest=raw["rest"], log_scales=raw["scaling"], raw_rotations=raw["rotation"])
        else:
            color, _, _ = r(means3D=raw["xyz"], means2D=m2, opacities=torch.sigmoid(raw["opacity"]),
                            shs=torch.cat((raw["dc"], raw["rest"]), dim=1), scales=torch.exp(raw["scaling"]),
                            rotations=torch.nn.functional.normalize(raw["rotation"]))
        (color * G).sum().backward()

    out = {}
    for name, fused in (("prologue", False), ("fused", True)):
        for i in range(5):
            step(i, fused)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            step(i, fused)
        e1.record()
        torch.cuda.synchronize()
        out[name + "_ms_per_step"] = e0.elapsed_time(e1) / a.steps
    out["saved_ms"] = out["prologue_ms_per_step"] - out["fused_ms_per_step"]
    out["config"] = f"{a.config}: P={cloud.means3D.shape[0]}, SH degree {cloud.sh_degree}, {W}x{H}"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
