"""Run a few fwd+bwd steps of BASELINE config 3 through the public API (target of ncu captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from gaussianeditor_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    _lib.set_option(k, int(v))
wl = bench.Workload(os.environ.get("GSR_CONFIG", "c3"), torch.device("cuda", 0))
r = bench.OursRunner(wl) if os.environ.get("GSR_IMPL", "ours") == "ours" else bench.ReferenceCudaRunner(wl)
for i in range(n):
    r.step(i)
torch.cuda.synchronize()
print("done", r.describe())
