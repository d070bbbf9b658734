"""Host-side mirror of the reference's render boundary
(/root/reference/gaussiansplatting/gaussian_renderer/__init__.py: ``camera2rasterizer`` :21-42, ``render`` :45-150).

Same names, arguments, result dictionary and conventions, so GaussianEditor's edit/add/delete loops that call
``render(cam, gaussians, pipe, bg)`` work on top of the B200 rasterizer without touching the callers.  The
optional Python-side SH->RGB / covariance paths of the reference (``pipe.convert_SHs_python`` /
``pipe.compute_cov3D_python``, both False in every GaussianEditor config, arguments/__init__.py:63-67) are
supported through the same rasterizer arguments.
"""
from __future__ import annotations

import math

import torch

from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer

C0 = 0.28209479177387814


def camera2rasterizer(viewpoint_camera, bg_color: torch.Tensor, sh_degree: int = 0):
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx,
        tanfovy=tanfovy,
        bg=bg_color,
        scale_modifier=1.0,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=sh_degree,
        campos=viewpoint_camera.camera_center,
        prefiltered=False,
        debug=False,
    )
    return GaussianRasterizer(raster_settings=raster_settings)


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None,
           fused_activations: bool = False):
    """Render the scene; background tensor must be on the GPU. Returns the reference's dictionary:
    render [3,H,W], viewspace_points [P,3] (grad sink for densification), visibility_filter, radii, depth_3dgs.

    ``fused_activations=True`` (opt-in, not in the reference) hands the scene model's RAW parameters
    (``pc._opacity, pc._features_dc, pc._features_rest, pc._scaling, pc._rotation``) to the rasterizer, which applies
    sigmoid / exp / normalize and reads the SH row from the two feature arrays inside its preprocess kernels: the
    per-render PyTorch prologue (five elementwise kernels, a [P,16,3] ``torch.cat``) and its autograd epilogue go away.
    Only valid for the plain SH path (no override colour, no Python-side SH / covariance)."""
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True,
                                          device=pc.get_xyz.device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass

    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx,
        tanfovy=tanfovy,
        bg=bg_color,
        scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center,
        prefiltered=False,
        debug=False,
    )
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)

    means3D = pc.get_xyz
    means2D = screenspace_points
    if fused_activations:
        if override_color is not None or getattr(pipe, "convert_SHs_python", False) or \
                getattr(pipe, "compute_cov3D_python", False):
            raise ValueError("fused_activations needs the plain SH + scale/rotation path")
        rendered_image, radii, depth = rasterizer.forward_raw(
            means3D=means3D, means2D=means2D, opacity_logits=pc._opacity, features_dc=pc._features_dc,
            features_rest=pc._features_rest, log_scales=pc._scaling, raw_rotations=pc._rotation)
        return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
                "radii": radii, "depth_3dgs": depth}
    opacity = pc.get_opacity

    scales = rotations = cov3D_precomp = None
    if getattr(pipe, "compute_cov3D_python", False):
        cov3D_precomp = pc.get_covariance(scaling_modifier)
    else:
        scales = pc.get_scaling
        rotations = pc.get_rotation

    shs = colors_precomp = None
    if override_color is None:
        if getattr(pipe, "convert_SHs_python", False):
            from .sh_utils import eval_sh
            shs_view = pc.get_features.transpose(1, 2).view(-1, 3, (pc.max_sh_degree + 1) ** 2)
            dir_pp = pc.get_xyz - viewpoint_camera.camera_center.repeat(pc.get_features.shape[0], 1)
            dir_pp_normalized = dir_pp / dir_pp.norm(dim=1, keepdim=True)
            sh2rgb = eval_sh(pc.active_sh_degree, shs_view, dir_pp_normalized)
            colors_precomp = torch.clamp_min(sh2rgb + 0.5, 0.0)
        else:
            shs = pc.get_features.float()
    else:
        colors_precomp = override_color

    rendered_image, radii, depth = rasterizer(
        means3D=means3D.float(),
        means2D=means2D.float(),
        shs=shs,
        colors_precomp=colors_precomp,
        opacities=opacity.float(),
        scales=None if scales is None else scales.float(),
        rotations=None if rotations is None else rotations.float(),
        cov3D_precomp=cov3D_precomp,
    )
    return {
        "render": rendered_image,
        "viewspace_points": screenspace_points,
        "visibility_filter": radii > 0,
        "radii": radii,
        "depth_3dgs": depth,
    }
