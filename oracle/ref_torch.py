"""ORACLE -- TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's Python surface (`GaussianRasterizer`, autograd included) on top of oracle/_ref (the reference's own
CUDA kernels), so the config-5 edit-loop harness can run the SAME loop against the reference for the A/B
render-time fraction.  Mirrors DGR/diff_gaussian_rasterization/__init__.py:50-309 with `_C` replaced by
oracle/ref_cuda.ReferenceRasterizer.  GPU only."""
import torch
import torch.nn as nn

from . import ref_cuda

_R = None


def _ref():
    global _R
    if _R is None:
        _R = ref_cuda.ReferenceRasterizer()
    return _R


class _RefRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, rs):
        R = _ref()
        c = lambda t: t.contiguous().float()
        args = dict(means3D=c(means3D), shs=c(sh) if sh.numel() else None,
                    colors_precomp=c(colors_precomp) if colors_precomp.numel() else None, scales=c(scales),
                    rotations=c(rotations), cov3D_precomp=None, bg=c(rs.bg), viewmatrix=c(rs.viewmatrix),
                    projmatrix=c(rs.projmatrix), campos=c(rs.campos), tanfovx=rs.tanfovx, tanfovy=rs.tanfovy,
                    sh_degree=rs.sh_degree, scale_modifier=rs.scale_modifier)
        color, radii, depth, n = R.forward(opacities=c(opacities), image_height=rs.image_height,
                                           image_width=rs.image_width, **args)
        ctx.args, ctx.n, ctx.radii = args, n, radii
        ctx.mark_non_differentiable(radii)
        return color, radii, depth

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth):
        g = _ref().backward(dL_dcolor=g_color.contiguous(), radii=ctx.radii, R=ctx.n, **ctx.args)
        sh = ctx.args["shs"]
        return (g["dL_dmeans3D"], g["dL_dmeans2D"], g["dL_dsh"] if sh is not None else None,
                g["dL_dcolors"] if ctx.args["colors_precomp"] is not None else None, g["dL_dopacity"],
                g["dL_dscales"], g["dL_drotations"], None)


class RefGaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        e = torch.empty(0, device=means3D.device)
        return _RefRasterize.apply(means3D, means2D, e if shs is None else shs,
                                   e if colors_precomp is None else colors_precomp, opacities, scales, rotations,
                                   self.raster_settings)
