"""Host-side overhead of one rasterizer call: tiny scene (GPU time negligible), wall-clock per forward / forward+backward,
ours vs the reference adapter. Also a cProfile of our path."""
import cProfile, io, json, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from gaussianeditor_b200 import synth
from gaussianeditor_b200.rasterizer import GaussianRasterizer
from util import settings_from, cloud_tensors
cloud, _ = synth.make_config("c3", P=2000)
cam = synth.ring_cameras(8, 4.5, 15.0, 64, 48, 61.0)[0]
dev = "cuda"
ct = cloud_tensors(cloud, dev, requires_grad=True)
rs = settings_from(cam, (0, 0, 0), 3, dev)
m2 = torch.zeros_like(ct["means3D"], requires_grad=True)
G = torch.rand(3, 48, 64, device=dev)
def step(cls, bwd):
    r = cls(rs)
    c, radii, d = r(means3D=ct["means3D"], means2D=m2, opacities=ct["opacities"], shs=ct["shs"], scales=ct["scales"], rotations=ct["rotations"])
    if bwd:
        (c * G).sum().backward()
def timeit(cls, bwd, n=300):
    for _ in range(20): step(cls, bwd)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): step(cls, bwd)
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e6
out = {"ours_fwd_us": timeit(GaussianRasterizer, False), "ours_fwd_bwd_us": timeit(GaussianRasterizer, True)}
try:
    from oracle import ref_torch
    out["ref_fwd_us"] = timeit(ref_torch.RefGaussianRasterizer, False); out["ref_fwd_bwd_us"] = timeit(ref_torch.RefGaussianRasterizer, True)
except Exception as ex:
    out["ref_error"] = repr(ex)
print(json.dumps(out))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step(GaussianRasterizer, True)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:4500])
