"""cProfile of the host side of one small fwd+bwd step (development tool): where does the Python/C-ABI overhead go?"""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from casualhdrsplat_b200 import rasterize  # noqa: E402
from casualhdrsplat_b200.scene import make_config  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c1"
sc = make_config(name).to("cuda:0")
leaves = {k: getattr(sc, k).clone().requires_grad_(True) for k in ["means", "quats", "scales", "opacities", "colors", "knots", "exposure_times",
                                                                    "frame_times"]}
crf = sc.crf_params.clone().requires_grad_(True) if sc.crf_params is not None else None
sp = dict(knots=leaves["knots"], knot_t0=sc.knot_t0, knot_dt=sc.knot_dt, frame_times=leaves["frame_times"], kind=sc.spline_kind)


def step():
    ldr, alpha, meta = rasterize(leaves["means"], leaves["quats"], leaves["scales"], leaves["opacities"], leaves["colors"], None, sc.Ks,
                                 sc.width, sc.height, leaves["exposure_times"], sc.n_virtual, sc.crf_kind, crf, spline=sp)
    (ldr * sc.v_ldr).sum().backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(50):
    step()
torch.cuda.synchronize()
print("ms per step", (time.perf_counter() - t0) / 50 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
