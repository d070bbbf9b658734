"""Instrumented forward pass on config 3: how much work the culling removes (counters of render_fwd variant 2)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from gaussianeditor_b200 import _lib
wl = bench.Workload("c3", torch.device("cuda", 0))
r = bench.OursRunner(wl)
_lib.set_option("render_fwd_variant", 2)
r.step(0)
_lib.set_option("stats", 1)
r.step(0)
torch.cuda.synchronize()
L = _lib.load()
names = ["staged_part_entries", "kept_part_entries", "subblock_evals", "subblock_evals_with_hit", "hit_lanes"]
st = {n: int(L.gsr_get_option(f"stat{i}".encode())) for i, n in enumerate(names)}
_lib.set_option("stats", 0)
d = r.describe()
st["R"] = d["R"]; st["mean_n_contrib"] = d["mean_n_contrib"]
st["lane_efficiency_of_evals"] = st["hit_lanes"] / max(1, 32 * st["subblock_evals"])
st["evals_per_kept"] = st["subblock_evals"] / max(1, st["kept_part_entries"])
print(json.dumps(st))
open(os.path.join(ROOT, "gpurun_out", "stats.json"), "w").write(json.dumps(st, indent=1))
