"""The on-disk format either side of the hot path: 3D-Gaussian-splatting ``point_cloud.ply`` files, without `plyfile`.

Format (what /root/reference/gaussiansplatting/scene/gaussian_model.py writes, :396-445, and reads, :455-551):
one ``vertex`` element whose properties are all ``float`` (f4), in this order

    x y z  nx ny nz  f_dc_0..2  f_rest_0..(3*(K)-1)  opacity  scale_0..2  rot_0..3        K = (D+1)^2 - 1

holding the RAW (pre-activation) parameters: opacity logits, log-scales, unnormalised quaternions (r,x,y,z). The SH
coefficients are stored CHANNEL-MAJOR: ``f_rest_{c*K + k}`` is coefficient k+1 of colour channel c, i.e. the in-memory
``features_rest [P, K, 3]`` transposed to [P, 3, K] and flattened (``f_dc`` likewise, [P,1,3] -> [P,3]). Normals are
written as zeros and ignored on load. The reader is tolerant the way the reference is: properties are found by NAME
(``f_rest_*``, ``scale_*``, ``rot*`` sorted by their numeric suffix), extra properties are ignored, and the SH degree is
derived from the number of ``f_rest_`` properties. It additionally accepts ascii / big-endian files and non-float
property types (converted to float32); list properties are rejected.
"""
from __future__ import annotations

import os
from typing import Dict

import numpy as np

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
    "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
    "double": "f8", "float64": "f8",
}


def attribute_names(n_rest: int) -> list:
    """Property order of the file for K = n_rest // 3 higher-order coefficients (gaussian_model.py:396-409)."""
    names = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(n_rest)]
    return names + ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]


def write_gaussian_ply(path: str, xyz, features_dc, features_rest, opacity, scaling, rotation) -> None:
    """Write raw parameters: xyz [P,3], features_dc [P,1,3], features_rest [P,K,3], opacity [P,1] (logits),
    scaling [P,3] (log), rotation [P,4]. Arrays may be numpy or torch (any device)."""
    def np32(a):
        if hasattr(a, "detach"):
            a = a.detach().cpu().numpy()
        return np.ascontiguousarray(a, dtype=np.float32)

    xyz, features_dc, features_rest = np32(xyz), np32(features_dc), np32(features_rest)
    opacity, scaling, rotation = np32(opacity), np32(scaling), np32(rotation)
    P = xyz.shape[0]
    if features_dc.shape != (P, 1, 3) or features_rest.ndim != 3 or features_rest.shape[0] != P or features_rest.shape[2] != 3:
        raise ValueError("features_dc must be [P,1,3] and features_rest [P,K,3]")
    K = features_rest.shape[1]
    cols = [xyz, np.zeros_like(xyz), features_dc.transpose(0, 2, 1).reshape(P, 3),
            features_rest.transpose(0, 2, 1).reshape(P, 3 * K), opacity.reshape(P, 1), scaling.reshape(P, 3),
            rotation.reshape(P, 4)]
    table = np.concatenate(cols, axis=1).astype("<f4")
    names = attribute_names(3 * K)
    assert table.shape[1] == len(names)
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {P}\n" + \
             "".join(f"property float {n}\n" for n in names) + "end_header\n"
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(path, "wb") as fh:
        fh.write(header.encode("ascii"))
        fh.write(table.tobytes())


def _read_header(fh):
    if fh.readline().strip() != b"ply":
        raise ValueError("not a PLY file")
    fmt = None
    elements = []  # [name, count, [(prop name, numpy type)]]
    while True:
        line = fh.readline()
        if not line:
            raise ValueError("unexpected end of PLY header")
        tok = line.decode("ascii", "replace").split()
        if not tok or tok[0] in ("comment", "obj_info"):
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append([tok[1], int(tok[2]), []])
        elif tok[0] == "property":
            if tok[1] == "list":
                raise ValueError("list properties are not part of the Gaussian PLY format")
            if tok[1] not in _PLY_TYPES:
                raise ValueError(f"unknown PLY property type {tok[1]}")
            elements[-1][2].append((tok[2], _PLY_TYPES[tok[1]]))
        elif tok[0] == "end_header":
            break
    if fmt not in ("binary_little_endian", "binary_big_endian", "ascii"):
        raise ValueError(f"unsupported PLY format {fmt}")
    return fmt, elements


def read_vertex_table(path: str) -> Dict[str, np.ndarray]:
    """All scalar properties of the first (vertex) element as float32 columns, by name."""
    with open(path, "rb") as fh:
        fmt, elements = _read_header(fh)
        if not elements:
            raise ValueError("PLY file has no elements")
        name, count, props = elements[0]
        if fmt == "ascii":
            rows = np.loadtxt(fh, dtype=np.float64, max_rows=count, ndmin=2) if count else np.zeros((0, len(props)))
            if rows.shape != (count, len(props)):
                raise ValueError("ascii PLY body does not match its header")
            return {n: rows[:, i].astype(np.float32) for i, (n, _) in enumerate(props)}
        end = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(n, end + t) for n, t in props])
        raw = fh.read(dt.itemsize * count)
        if len(raw) != dt.itemsize * count:
            raise ValueError("PLY body is shorter than its header says")
        rec = np.frombuffer(raw, dtype=dt, count=count)
        return {n: rec[n].astype(np.float32) for n, _ in props}


def _numbered(cols: Dict[str, np.ndarray], prefix: str) -> list:
    names = [n for n in cols if n.startswith(prefix)]
    return sorted(names, key=lambda n: int(n.split("_")[-1]))


def read_gaussian_ply(path: str) -> Dict[str, np.ndarray]:
    """Raw parameters in the in-memory layout of the scene model: xyz [P,3], features_dc [P,1,3],
    features_rest [P,K,3], opacity [P,1], scaling [P,3], rotation [P,4], plus ``max_sh_degree``."""
    cols = read_vertex_table(path)
    for need in ("x", "y", "z", "opacity", "f_dc_0", "f_dc_1", "f_dc_2"):
        if need not in cols:
            raise ValueError(f"PLY file lacks property {need}")
    P = cols["x"].shape[0]
    xyz = np.stack([cols["x"], cols["y"], cols["z"]], axis=1)
    rest_names = _numbered(cols, "f_rest_")
    deg = int(((len(rest_names) + 3) / 3) ** 0.5 - 1)       # gaussian_model.py:479-480
    K = (deg + 1) ** 2 - 1
    if len(rest_names) != 3 * K:
        raise ValueError(f"{len(rest_names)} f_rest_ properties do not form a full SH degree")
    dc = np.stack([cols["f_dc_0"], cols["f_dc_1"], cols["f_dc_2"]], axis=1).reshape(P, 3, 1)
    rest = (np.stack([cols[n] for n in rest_names], axis=1) if rest_names else np.zeros((P, 0), np.float32)).reshape(P, 3, K)
    scale_names, rot_names = _numbered(cols, "scale_"), _numbered(cols, "rot")
    if len(scale_names) != 3 or len(rot_names) != 4:
        raise ValueError("expected scale_0..2 and rot_0..3")
    c = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(xyz=c(xyz), features_dc=c(dc.transpose(0, 2, 1)), features_rest=c(rest.transpose(0, 2, 1)),
                opacity=c(cols["opacity"].reshape(P, 1)), scaling=c(np.stack([cols[n] for n in scale_names], axis=1)),
                rotation=c(np.stack([cols[n] for n in rot_names], axis=1)), max_sh_degree=deg)


def activate(params: Dict[str, np.ndarray]):
    """Raw parameters -> what GaussianRasterizer.forward takes (scene/gaussian_model.py:221-258: exp, sigmoid,
    normalize, cat), as a ``synth.Cloud``."""
    from .synth import Cloud
    rot = params["rotation"].astype(np.float32)
    rot = rot / np.maximum(np.linalg.norm(rot, axis=1, keepdims=True), 1e-12).astype(np.float32)
    op = (1.0 / (1.0 + np.exp(-params["opacity"].astype(np.float64)))).astype(np.float32)
    shs = np.concatenate([params["features_dc"], params["features_rest"]], axis=1).astype(np.float32)
    return Cloud(means3D=params["xyz"].astype(np.float32), scales=np.exp(params["scaling"]).astype(np.float32),
                 rotations=rot.astype(np.float32), opacities=op, shs=np.ascontiguousarray(shs),
                 sh_degree=int(params["max_sh_degree"]))
