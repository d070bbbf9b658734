"""Shared helpers of the test-suite."""
import math

import numpy as np
import torch

from gaussianeditor_b200 import synth
from gaussianeditor_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer


def settings_from(cam: synth.Camera, bg, sh_degree, device, scale_modifier=1.0, debug=False):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)
    return GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=t(np.asarray(bg, np.float32)), scale_modifier=scale_modifier, viewmatrix=t(cam.viewmatrix),
        projmatrix=t(cam.projmatrix), sh_degree=sh_degree, campos=t(cam.campos), prefiltered=False, debug=debug)


def cloud_tensors(cloud: synth.Cloud, device, requires_grad=False):
    t = lambda a: torch.from_numpy(a).to(device).requires_grad_(requires_grad)
    return dict(means3D=t(cloud.means3D), opacities=t(cloud.opacities), shs=t(cloud.shs), scales=t(cloud.scales),
                rotations=t(cloud.rotations))


def run_ours(cloud, cam, bg=(0, 0, 0), dL=None, colors_precomp=None, scale_modifier=1.0, device="cuda"):
    """Forward (+ backward if dL is given) through the public GaussianRasterizer API. Returns dict."""
    from gaussianeditor_b200.rasterizer import _RasterizeGaussians, forward_state_views
    ct = cloud_tensors(cloud, device, requires_grad=dL is not None)
    rs = settings_from(cam, bg, cloud.sh_degree, device, scale_modifier)
    rast = GaussianRasterizer(rs)
    means2D = torch.zeros_like(ct["means3D"], requires_grad=dL is not None)
    kw = dict(means3D=ct["means3D"], means2D=means2D, opacities=ct["opacities"], scales=ct["scales"],
              rotations=ct["rotations"])
    cp = None
    if colors_precomp is not None:
        cp = torch.from_numpy(colors_precomp).to(device).requires_grad_(dL is not None)
        kw["colors_precomp"] = cp
    else:
        kw["shs"] = ct["shs"]
    color, radii, depth = rast(**kw)
    state = _RasterizeGaussians.last_state
    out = dict(color=color.detach(), radii=radii, depth=depth.detach(), R=state.num_rendered,
               views=forward_state_views(state), state=state)
    if dL is not None:
        (color * torch.from_numpy(dL).to(device)).sum().backward()
        out["grads"] = dict(dmean3D=ct["means3D"].grad, dmean2D=means2D.grad, dopacity=ct["opacities"].grad,
                            dscale=ct["scales"].grad, drot=ct["rotations"].grad,
                            dsh=None if cp is not None else ct["shs"].grad, dcolor=None if cp is None else cp.grad)
    return out


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
