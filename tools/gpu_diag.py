"""First-contact GPU diagnostics: ours vs the compiled reference (oracle/_ref) -- mismatch statistics instead of
asserts, then a timing sweep over kernel variants on BASELINE config 3. Output: gpurun_out/diag.json + stdout."""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

from gaussianeditor_b200 import synth, _lib
from oracle import ref_cuda
from util import run_ours, rel_l2
import test_parity_gpu as T

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
report = {}


def compare(name, cloud, cam, bg, fv, bv):
    _lib.set_option("render_fwd_variant", fv)
    _lib.set_option("render_bwd_variant", bv)
    dL = np.random.default_rng(7).uniform(size=(3, cam.image_height, cam.image_width)).astype(np.float32)
    ours = run_ours(cloud, cam, bg, dL=dL)
    ref = T._ref_run(cloud, cam, bg, dL=dL)
    ref2 = T._ref_run(cloud, cam, bg, dL=dL)
    v, s = ours["views"], ref["state"]
    r = dict(R=(ours["R"], ref["R"]))
    r["radii_mismatch"] = int((ours["radii"] != ref["radii"]).sum())
    r["tiles_mismatch"] = int((v["tiles_touched"] != s["tiles_touched"]).sum())
    if ours["R"] == ref["R"]:
        r["ranges_mismatch"] = int((v["ranges"] != s["ranges"]).sum())
        r["point_list_mismatch"] = int((v["point_list"] != s["point_list"]).sum())
    r["n_contrib_mismatch"] = int((v["n_contrib"] != s["n_contrib"]).sum())
    r["final_T_maxdiff"] = float((v["final_T"] - s["final_T"]).abs().max())
    r["color_maxdiff"] = float((ours["color"] - ref["color"]).abs().max())
    r["color_neq"] = int((ours["color"] != ref["color"]).sum())
    r["depth_maxdiff"] = float((ours["depth"] - ref["depth"]).abs().max())
    vis = (ours["radii"] > 0) & (ref["radii"] > 0)
    rec = v["records"][vis]
    r["means2D_neq"] = int((rec[:, 0:2] != s["means2D"][vis]).sum())
    r["conic_neq"] = int((rec[:, [2, 3, 4, 5]] != s["conic_opacity"][vis]).sum())
    r["conic_maxrel"] = float(((rec[:, [2, 3, 4]] - s["conic_opacity"][vis][:, :3]).abs() /
                               (s["conic_opacity"][vis][:, :3].abs() + 1e-30)).max()) if vis.any() else 0.0
    r["depth_neq"] = int((rec[:, 6] != s["depths"][vis]).sum())
    r["rgb_neq"] = int((rec[:, 8:11] != s["rgb"][vis]).sum())
    r["rgb_maxdiff"] = float((rec[:, 8:11] - s["rgb"][vis]).abs().max()) if vis.any() else 0.0
    for a, b in [("dmean3D", "dL_dmeans3D"), ("dmean2D", "dL_dmeans2D"), ("dopacity", "dL_dopacity"),
                 ("dscale", "dL_dscales"), ("drot", "dL_drotations"), ("dsh", "dL_dsh")]:
        g, rr = ours["grads"][a].cpu().numpy(), ref["grads"][b].cpu().numpy()
        r["grad_" + a] = dict(rel=rel_l2(g, rr), noise=rel_l2(ref2["grads"][b].cpu().numpy(), rr),
                              maxabs=float(np.abs(g - rr).max()), refmax=float(np.abs(rr).max()))
    return r


try:
    for name, cloud, cam, bg in T._small_cases():
        for fv, bv in [(0, 0), (2, 2), (3, 3)]:
            key = f"{name}_f{fv}_b{bv}"
            try:
                report[key] = compare(name, cloud, cam, bg, fv, bv)
            except Exception as ex:
                report[key] = dict(error=repr(ex), tb=traceback.format_exc())
            print(key, json.dumps(report[key])[:1500], flush=True)
except Exception:
    traceback.print_exc()

# ---- timing sweep on config 3 ---------------------------------------------------------------------------
try:
    import bench
    dev = torch.device("cuda", 0)
    wl = bench.Workload("c3", dev)
    timing = {}

    def time_runner(runner, n=16, warm=4):
        for i in range(warm):
            runner.step(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            runner.step(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    ours = bench.OursRunner(wl)
    for fv, bv in [(1, 1), (2, 2), (3, 3)]:
        if True:
            _lib.set_option("render_fwd_variant", fv)
            _lib.set_option("render_bwd_variant", bv)
            ms = time_runner(ours)
            _lib.set_option("profile", 1); _lib.profile_read()
            for i in range(8):
                ours.step(i)
            prof = _lib.profile_read(); _lib.set_option("profile", 0)
            timing[f"ours_f{fv}_b{bv}"] = dict(ms_per_step=ms, stages={k: v[0] / max(v[1], 1) for k, v in prof.items() if v[1]})
            print(f"ours_f{fv}_b{bv}", json.dumps(timing[f"ours_f{fv}_b{bv}"]), flush=True)
    _lib.set_option("render_fwd_variant", 2); _lib.set_option("render_bwd_variant", 2)
    for pv in [32, 16]:
        _lib.set_option("tile_key_bits", pv)
        _lib.set_option("profile", 1); _lib.profile_read()
        for i in range(8):
            ours.step(i)
        prof = _lib.profile_read(); _lib.set_option("profile", 0)
        timing[f"ours_pre{pv}"] = {k: v[0] / max(v[1], 1) for k, v in prof.items() if v[1]}
        print(f"ours_pre{pv}", json.dumps(timing[f"ours_pre{pv}"]), flush=True)
    _lib.set_option("tile_key_bits", 16)
    timing["workload"] = ours.describe()
    if ref_cuda.available():
        refr = bench.ReferenceCudaRunner(wl)
        timing["reference_ms_per_step"] = time_runner(refr, n=8, warm=2)
        # forward-only / backward split of the reference
        torch.cuda.synchronize()
        print("reference", timing["reference_ms_per_step"], flush=True)
    report["timing"] = timing
except Exception:
    traceback.print_exc()
    report["timing_error"] = traceback.format_exc()

with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
    json.dump(report, f, indent=1)
print("DIAG DONE")
