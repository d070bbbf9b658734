"""Timing sweep over kernel variants on config 3 (per-stage CUDA-event times)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from gaussianeditor_b200 import _lib
fw = [int(x) for x in os.environ.get("FWD", "3").split(",")]
bw = [int(x) for x in os.environ.get("BWD", "2,4,8,9").split(",")]
wl = bench.Workload("c3", torch.device("cuda", 0))
r = bench.OursRunner(wl)
res = {}
def stage_times(n=12):
    for i in range(4): r.step(i)
    _lib.set_option("profile", 1); _lib.profile_read()
    for i in range(n): r.step(i)
    p = _lib.profile_read(); _lib.set_option("profile", 0)
    return {k: round(v[0] / max(v[1], 1), 4) for k, v in p.items() if v[1]}
for f in fw:
    _lib.set_option("render_fwd_variant", f)
    res[f"fwd{f}"] = stage_times()["render_fwd"]
_lib.set_option("render_fwd_variant", 3)
for b in bw:
    _lib.set_option("render_bwd_variant", b)
    res[f"bwd{b}"] = stage_times()["render_bwd"]
_lib.set_option("render_bwd_variant", 4)
res["all"] = stage_times()
print(json.dumps(res))
open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w").write(json.dumps(res, indent=1))
