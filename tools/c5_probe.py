"""Diagnostics: per-stage device times of the hot path at BASELINE config 5's shape (500k Gaussians, 512x512) for both
binning variants, and the edit-loop step time with each."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from gaussianeditor_b200 import _lib, edit_loop
from gaussianeditor_b200.rasterizer import GaussianRasterizer
import gaussianeditor_b200.rasterizer as RZ

out = {}
dev = torch.device("cuda", 0)
wl = bench.Workload("c5", dev)
r = bench.OursRunner(wl)
for variant in (0, 1):
    _lib.set_option("binning_variant", variant)
    for i in range(10):
        r.step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(48):
        r.step(i)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 48 * 1e3
    _lib.set_option("profile", 1); _lib.profile_read()
    for i in range(24):
        r.step(i)
    prof = _lib.profile_read(); _lib.set_option("profile", 0)
    out[f"variant{variant}"] = {"ms_per_step_wall": wall, "stages_ms": {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if v[1]},
                                "desc": r.describe(), "spec": dict(RZ.SPEC_STATS)}
    e = edit_loop.run_edit_loop(GaussianRasterizer, steps=60, densification_interval=30)
    out[f"variant{variant}"]["edit_loop"] = {k: e[k] for k in ("ms_per_step", "render_ms", "render_fraction")}
RZ.SPECULATIVE = False
_lib.set_option("binning_variant", 1)
e = edit_loop.run_edit_loop(GaussianRasterizer, steps=60, densification_interval=30)
out["variant1_no_speculation"] = {k: e[k] for k in ("ms_per_step", "render_ms", "render_fraction")}
print(json.dumps(out))
