"""the PLY wire format and Adam-state surgery either side of the hot path (CPU only)."""
import struct

import numpy as np
import pytest
import torch

from gaussianeditor_b200 import optim_surgery as OS, ply_io


def _params(P=7, deg=3, seed=0):
    rng = np.random.default_rng(seed)
    K = (deg + 1) ** 2 - 1
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    return dict(xyz=f(P, 3), features_dc=f(P, 1, 3), features_rest=f(P, K, 3), opacity=f(P, 1), scaling=f(P, 3),
                rotation=f(P, 4))


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_ply_round_trip_is_bit_exact(tmp_path, deg):
    p = _params(deg=deg)
    path = str(tmp_path / "pc" / "point_cloud.ply")
    ply_io.write_gaussian_ply(path, **p)
    q = ply_io.read_gaussian_ply(path)
    assert q["max_sh_degree"] == deg
    for k, v in p.items():
        assert q[k].dtype == np.float32 and q[k].shape == v.shape and np.array_equal(q[k], v), k


def test_ply_bytes_follow_the_reference_layout(tmp_path):
    """Known-answer: header text, property order, f4 little-endian rows, channel-major SH."""
    p = _params(P=2, deg=1)
    path = str(tmp_path / "a.ply")
    ply_io.write_gaussian_ply(path, **p)
    raw = open(path, "rb").read()
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(9)] + \
            ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex 2\n" + "".join(f"property float {n}\n" for n in names)
              + "end_header\n").encode()
    assert raw.startswith(header) and len(raw) == len(header) + 2 * 4 * len(names)
    row0 = struct.unpack("<" + "f" * len(names), raw[len(header):len(header) + 4 * len(names)])
    want = list(p["xyz"][0]) + [0, 0, 0] + list(p["features_dc"][0, 0]) + \
        [p["features_rest"][0, k, c] for c in range(3) for k in range(3)] + \
        list(p["opacity"][0]) + list(p["scaling"][0]) + list(p["rotation"][0])
    assert np.array_equal(np.float32(row0), np.float32(want))


def test_ply_reader_is_name_based_and_tolerant(tmp_path):
    """Extra properties, shuffled order, double-typed columns, ascii and big-endian encodings."""
    p = _params(P=3, deg=1, seed=4)
    ref_path = str(tmp_path / "ref.ply")
    ply_io.write_gaussian_ply(ref_path, **p)
    cols = ply_io.read_vertex_table(ref_path)
    order = list(reversed(list(cols))) + ["extra_thing"]
    cols["extra_thing"] = np.arange(3, dtype=np.float32)
    for fmt in ("ascii", "binary_big_endian"):
        path = str(tmp_path / f"{fmt}.ply")
        with open(path, "wb") as fh:
            fh.write(f"ply\nformat {fmt} 1.0\ncomment made by a test\nelement vertex 3\n".encode())
            for n in order:
                fh.write(f"property {'double' if n == 'x' else 'float'} {n}\n".encode())
            fh.write(b"end_header\n")
            for i in range(3):
                if fmt == "ascii":
                    fh.write((" ".join(repr(float(cols[n][i])) for n in order) + "\n").encode())
                else:
                    fh.write(b"".join(struct.pack(">d" if n == "x" else ">f", float(cols[n][i])) for n in order))
        q = ply_io.read_gaussian_ply(path)
        for k, v in p.items():
            assert np.array_equal(q[k], v), (fmt, k)
    with open(str(tmp_path / "bad.ply"), "wb") as fh:
        fh.write(b"ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty list uchar int vertex_indices\nend_header\n")
    with pytest.raises(ValueError):
        ply_io.read_gaussian_ply(str(tmp_path / "bad.ply"))


def test_activate_matches_scene_model_getters():
    p = _params(P=50, deg=2, seed=2)
    p["max_sh_degree"] = 2
    c = ply_io.activate(p)
    t = {k: torch.from_numpy(v) for k, v in p.items() if k != "max_sh_degree"}
    assert np.allclose(c.opacities, torch.sigmoid(t["opacity"]).numpy(), rtol=1e-6, atol=1e-7)
    assert np.allclose(c.scales, torch.exp(t["scaling"]).numpy(), rtol=1e-6)
    assert np.allclose(c.rotations, torch.nn.functional.normalize(t["rotation"]).numpy(), rtol=1e-6, atol=1e-7)
    assert np.array_equal(c.shs, torch.cat((t["features_dc"], t["features_rest"]), dim=1).numpy()) and c.shs.shape == (50, 9, 3)
    assert c.sh_degree == 2


def _model(P=10):
    g = torch.Generator().manual_seed(0)
    names = {"xyz": (P, 3), "f_dc": (P, 1, 3), "f_rest": (P, 15, 3), "opacity": (P, 1), "scaling": (P, 3), "rotation": (P, 4)}
    params = {n: torch.nn.Parameter(torch.randn(*s, generator=g)) for n, s in names.items()}
    opt = torch.optim.Adam([{"params": [p], "lr": 0.01, "name": n} for n, p in params.items()], lr=0.0, eps=1e-15)
    for _ in range(3):
        loss = sum((p ** 2).sum() for p in params.values())
        opt.zero_grad(); loss.backward(); opt.step()
    return params, opt


def test_prune_keeps_rows_and_their_adam_moments():
    params, opt = _model()
    keep = torch.tensor([True, False, True, True, False, True, True, True, False, True])
    before = {n: (p.detach().clone(), opt.state[p]["exp_avg"].clone(), opt.state[p]["exp_avg_sq"].clone(),
                  opt.state[p]["step"].clone()) for n, p in params.items()}
    new = OS.prune_optimizer(opt, keep)
    assert set(new) == set(params)
    for g in opt.param_groups:
        p = g["params"][0]
        val, m, v, step = before[g["name"]]
        assert p is new[g["name"]] and p.requires_grad and p.shape[0] == int(keep.sum())
        assert torch.equal(p.detach(), val[keep]) and torch.equal(opt.state[p]["exp_avg"], m[keep])
        assert torch.equal(opt.state[p]["exp_avg_sq"], v[keep]) and torch.equal(opt.state[p]["step"], step)
    assert len(opt.state) == len(params)  # the old parameter objects left the state dict
    loss = sum((g["params"][0] ** 2).sum() for g in opt.param_groups)
    opt.zero_grad(); loss.backward(); opt.step()  # still steps


def test_cat_appends_rows_with_zero_moments_and_replace_resets():
    params, opt = _model()
    ext = {n: torch.ones(4, *p.shape[1:]) for n, p in params.items()}
    before = {n: (p.detach().clone(), opt.state[p]["exp_avg"].clone()) for n, p in params.items()}
    new = OS.cat_tensors_to_optimizer(opt, ext)
    for n, p in new.items():
        val, m = before[n]
        assert p.shape[0] == 14 and torch.equal(p.detach()[:10], val) and torch.equal(p.detach()[10:], ext[n])
        assert torch.equal(opt.state[p]["exp_avg"][:10], m) and float(opt.state[p]["exp_avg"][10:].abs().sum()) == 0.0
        assert float(opt.state[p]["exp_avg_sq"][10:].abs().sum()) == 0.0
    rep = OS.replace_tensor_to_optimizer(opt, torch.full((14, 1), -4.0), "opacity")
    p = rep["opacity"]
    assert list(rep) == ["opacity"] and torch.equal(p.detach(), torch.full((14, 1), -4.0))
    assert float(opt.state[p]["exp_avg"].abs().sum()) == 0.0 and float(opt.state[p]["exp_avg_sq"].abs().sum()) == 0.0
    with pytest.raises(ValueError):
        OS.cat_tensors_to_optimizer(opt, {n: torch.ones(2, 5) for n in params})
