"""Eager sync-free step versus the same step replayed from a CUDA graph (development tool).
Usage: python tools/graph_times.py [config] [steps]   ->  one JSON line (ms per step of all the config's frames)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from casualhdrsplat_b200.parallel import GraphedStep, StepState, formation_step  # noqa: E402
from casualhdrsplat_b200.scene import make_config  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    sc = make_config(name)
    dev = torch.device("cuda", 0)
    names = ["means", "quats", "scales", "opacities", "colors", "knots", "frame_times", "exposure_times", "Ks", "crf_params"]
    P = {k: getattr(sc, k).to(dev).contiguous() for k in names if getattr(sc, k) is not None}
    v = sc.v_ldr.to(dev)
    idx = {i: torch.tensor([i], device=dev) for i in range(sc.n_frames)}
    up = lambda f, ldr: torch.cat([v.index_select(0, idx[i]) for i in f])  # noqa: E731
    meta = {"knot_t0": sc.knot_t0, "knot_dt": sc.knot_dt, "kind": sc.spline_kind}
    args = (P, meta, sc.width, sc.height, sc.n_virtual, sc.crf_kind, list(range(sc.n_frames)), up)
    kw = dict(micro_batch=1, tight_bounds=True)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    st = StepState()
    for _ in range(3):
        formation_step(*args, state=st, **kw)
    st.verify()
    eager = timed(lambda: formation_step(*args, state=st, **kw))
    st.verify()
    gs = GraphedStep(*args, **kw)
    graph = timed(gs.replay)
    gs.verify()
    print(json.dumps({"config": name, "frames": sc.n_frames, "eager_ms_per_step": round(eager, 4), "graph_ms_per_step": round(graph, 4),
                      "eager_frames_per_s": round(sc.n_frames / eager * 1e3, 1), "graph_frames_per_s": round(sc.n_frames / graph * 1e3, 1)}))


if __name__ == "__main__":
    main()
