"""Per-stage CUDA-event timing of one fwd+bwd step through the C ABI (development tool).
Usage: python tools/stage_times.py [config] [sort_mode] [iters]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from casualhdrsplat_b200 import _lib, rasterize  # noqa: E402
from casualhdrsplat_b200.scene import make_config  # noqa: E402

STAGES = ["chs_spline_fwd", "chs_project_fwd", "chs_bin_count", "chs_bin_sort", "chs_blend_fwd", "chs_crf_bwd", "chs_blend_bwd",
          "chs_project_bwd", "chs_spline_bwd"]


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    sort_mode = sys.argv[2] if len(sys.argv) > 2 else "presort"
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    over = {}
    tight = False
    tuning = {}
    fused = False
    for a in sys.argv[4:]:  # e.g. n_frames=1 n_gauss=2000000 tight=1 tune_bin=1 tune_blend_bwd=40 pose_fused=1
        k, v = a.split("=")
        if k == "tight":
            tight = bool(int(v))
        elif k == "pose_fused":
            fused = bool(int(v))
        elif k.startswith("tune_"):
            tuning[k[5:]] = int(v)
        else:
            over[k] = int(v)
    t0 = time.time()
    sc = make_config(name, **over).to("cuda:0")
    print(f"scene {name} built in {time.time() - t0:.1f}s", flush=True)
    L = _lib.lib()
    records = {k: [] for k in STAGES}
    orig = {k: getattr(L, k) for k in STAGES}

    def wrap(k):
        f = orig[k]

        def g(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = f(*a)
            e1.record()
            records[k].append((e0, e1))
            return r
        return g

    for k in STAGES:
        setattr(L, k, wrap(k))
    leaves = {k: getattr(sc, k).clone().requires_grad_(True) for k in ["means", "quats", "scales", "opacities", "colors", "knots",
                                                                        "exposure_times", "frame_times"]}
    crf = sc.crf_params.clone().requires_grad_(True) if sc.crf_params is not None else None
    sp = dict(knots=leaves["knots"], knot_t0=sc.knot_t0, knot_dt=sc.knot_dt, frame_times=leaves["frame_times"], kind=sc.spline_kind)
    total = []
    M = 0
    for it in range(iters + 2):
        for k in STAGES:
            records[k].clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        ldr, alpha, meta = rasterize(leaves["means"], leaves["quats"], leaves["scales"], leaves["opacities"], leaves["colors"], None, sc.Ks,
                                     sc.width, sc.height, leaves["exposure_times"], sc.n_virtual, sc.crf_kind, crf, spline=sp,
                                     sort_mode=sort_mode, tight_bounds=tight, tuning=tuning, pose_fused=fused)
        (ldr * sc.v_ldr).sum().backward()
        e1.record()
        torch.cuda.synchronize()
        M = meta["n_isect"]
        if it >= 2:
            total.append(e0.elapsed_time(e1))
            row = {k: sum(a.elapsed_time(b) for a, b in records[k]) for k in STAGES}
            print(json.dumps({"iter": it, "total_ms": round(total[-1], 3), **{k[4:]: round(v, 3) for k, v in row.items()}}), flush=True)
        del ldr, alpha, meta
    st = sorted(total)
    tt = {"config": name, "sort_mode": sort_mode, "tight_bounds": tight, "tuning": tuning, "pose_fused": fused, "M": M, "median_ms": st[len(st) // 2], "min_ms": st[0],
          "frames_per_s": sc.n_frames / (st[len(st) // 2] / 1e3), "mem_GB": torch.cuda.max_memory_allocated() / 1e9}
    print(json.dumps(tt), flush=True)


if __name__ == "__main__":
    main()
