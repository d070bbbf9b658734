"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of ``oracle/liboracle_cpu.so`` (the CPU restatement in
``oracle/cpu_rasterizer.cpp``).  Only tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module.

Numpy in, numpy out; mirrors the call shape of the reference's private ``_C`` module
(/root/reference/gaussiansplatting/submodules/diff-gaussian-rasterization/ext.cpp:15-20,
rasterize_points.cu:35-157): ``forward`` returns every intermediate buffer the reference keeps
(so tests can compare radii / tile ranges / sorted lists bit for bit), ``backward`` returns the
nine gradient tensors.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_cpu.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cpu_rasterizer.cpp")
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "cpu"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_forward.restype = C.c_void_p
        _lib.oracle_num_rendered.restype = C.c_int64
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _fp(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


class Forward:
    """One forward pass; keeps the native handle alive for ``backward``/``apply_weights``."""

    def __init__(self, *, means3D, opacities, viewmatrix, projmatrix, campos, tanfovx, tanfovy,
                 image_height, image_width, bg, shs=None, colors_precomp=None, scales=None,
                 rotations=None, cov3D_precomp=None, sh_degree=0, scale_modifier=1.0, f32=True,
                 render=True):
        L = lib()
        self.f32 = 1 if f32 else 0
        self.F = np.float32 if f32 else np.float64
        means3D = _f32(means3D)
        P = means3D.shape[0]
        shs = _f32(shs) if shs is not None and np.size(shs) else None
        colors_precomp = _f32(colors_precomp) if colors_precomp is not None and np.size(colors_precomp) else None
        scales = _f32(scales) if scales is not None and np.size(scales) else None
        rotations = _f32(rotations) if rotations is not None and np.size(rotations) else None
        cov3D_precomp = _f32(cov3D_precomp) if cov3D_precomp is not None and np.size(cov3D_precomp) else None
        M = shs.shape[1] if shs is not None else 0
        self.P, self.M, self.W, self.H = P, M, int(image_width), int(image_height)
        opacities = _f32(opacities)
        bg, viewmatrix, projmatrix, campos = _f32(bg), _f32(viewmatrix), _f32(projmatrix), _f32(campos)
        self.h = C.c_void_p(L.oracle_forward(
            self.f32, P, int(sh_degree), M, self.W, self.H, _fp(bg), _fp(means3D), _fp(shs),
            _fp(colors_precomp), _fp(opacities), _fp(scales), C.c_float(scale_modifier), _fp(rotations),
            _fp(cov3D_precomp), _fp(viewmatrix), _fp(projmatrix), _fp(campos), C.c_float(tanfovx),
            C.c_float(tanfovy), 1 if render else 0))
        self.num_rendered = int(L.oracle_num_rendered(self.f32, self.h))
        R, F, W, H = self.num_rendered, self.F, self.W, self.H
        ntile = ((W + 15) // 16) * ((H + 15) // 16)
        o = self
        o.color = np.zeros((3, H, W), F); o.depth = np.zeros((1, H, W), F)
        o.radii = np.zeros(P, np.int32); o.final_T = np.zeros((H, W), F)
        o.n_contrib = np.zeros((H, W), np.uint32); o.depths = np.zeros(P, F)
        o.means2D = np.zeros((P, 2), F); o.conic_opacity = np.zeros((P, 4), F)
        o.rgb = np.zeros((P, 3), F); o.cov3D = np.zeros((P, 6), F)
        o.clamped = np.zeros((P, 3), np.uint8); o.tiles_touched = np.zeros(P, np.uint32)
        o.point_offsets = np.zeros(P, np.uint32); o.keys = np.zeros(R, np.uint64)
        o.point_list = np.zeros(R, np.uint32); o.ranges = np.zeros((ntile, 2), np.uint32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        L.oracle_get_forward(self.f32, self.h, vp(o.color), vp(o.depth), vp(o.radii), vp(o.final_T),
                             vp(o.n_contrib), vp(o.depths), vp(o.means2D), vp(o.conic_opacity), vp(o.rgb),
                             vp(o.cov3D), vp(o.clamped), vp(o.tiles_touched), vp(o.point_offsets), vp(o.keys),
                             vp(o.point_list), vp(o.ranges))
        pairs, hits = C.c_int64(0), C.c_int64(0)
        L.oracle_stats(self.f32, self.h, C.byref(pairs), C.byref(hits))
        self.pairs, self.hits = pairs.value, hits.value

    def backward(self, dL_dpix):
        """dL_dpix [3,H,W] -> dict of the reference's gradient tensors (rasterize_points.cu:120-128)."""
        L = lib()
        dL = np.ascontiguousarray(dL_dpix, dtype=np.float32)
        assert dL.shape == (3, self.H, self.W)
        P, M, F = self.P, self.M, self.F
        g = dict(dmean2D=np.zeros((P, 3), F), dconic=np.zeros((P, 2, 2), F), dopacity=np.zeros((P, 1), F),
                 dcolor=np.zeros((P, 3), F), dmean3D=np.zeros((P, 3), F), dcov3D=np.zeros((P, 6), F),
                 dsh=np.zeros((P, M, 3), F), dscale=np.zeros((P, 3), F), drot=np.zeros((P, 4), F))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        L.oracle_backward(self.f32, self.h, vp(dL), vp(g["dmean2D"]), vp(g["dconic"]), vp(g["dopacity"]),
                          vp(g["dcolor"]), vp(g["dmean3D"]), vp(g["dcov3D"]), vp(g["dsh"]), vp(g["dscale"]),
                          vp(g["drot"]))
        return g

    def apply_weights(self, weights, cnt, image_weights):
        """In-place semantic tracing (apply_weights.cu:240-356); f32 handle with colors_precomp only."""
        assert self.f32
        CH = image_weights.shape[0]
        assert weights.dtype == np.float32 and cnt.dtype == np.int32
        iw = np.ascontiguousarray(image_weights, dtype=np.float32)
        lib().oracle_apply_weights(self.h, weights.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p),
                                   iw.ctypes.data_as(C.c_void_p), CH)

    def close(self):
        if self.h:
            lib().oracle_free(self.f32, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def mark_visible(means3D, viewmatrix):
    means3D, viewmatrix = _f32(means3D), _f32(viewmatrix)
    out = np.zeros(means3D.shape[0], np.uint8)
    lib().oracle_mark_visible(means3D.shape[0], _fp(means3D), _fp(viewmatrix), out.ctypes.data_as(C.c_void_p))
    return out.astype(bool)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def forward_from(cloud, cam, bg=(0.0, 0.0, 0.0), **kw):
    """Convenience: run a synth.Cloud through a synth.Camera."""
    args = dict(means3D=cloud.means3D, opacities=cloud.opacities, scales=cloud.scales, rotations=cloud.rotations,
                shs=cloud.shs, sh_degree=cloud.sh_degree, viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix,
                campos=cam.campos, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, image_height=cam.image_height,
                image_width=cam.image_width, bg=np.asarray(bg, np.float32))
    args.update(kw)
    if args.get("colors_precomp") is not None:
        args["shs"] = None
    return Forward(**args)
