"""Full (not extrapolated) timing of the float64 CPU oracle on one BASELINE.json configuration, forward + backward.
Usage: python tests/tools/cpu_oracle_full.py [config=c2] [threads=all]   (result recorded in BASELINE.md)"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from casualhdrsplat_b200.scene import make_config  # noqa: E402
from tests.util import oracle_run  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    sc = make_config(name)
    t0 = time.perf_counter()
    ldr, alpha, meta, grads = oracle_run(sc, with_grad=False)
    t_fwd = time.perf_counter() - t0
    t0 = time.perf_counter()
    ldr, alpha, meta, grads = oracle_run(sc)
    t_all = time.perf_counter() - t0
    print(json.dumps({"config": name, "threads": threads, "n_isect": int(meta["n_isect"]), "fwd_only_s": round(t_fwd, 2),
                      "fwd_bwd_s": round(t_all, 2), "frames_per_s_fwd_bwd": sc.n_frames / t_all,
                      "what": "oracle.rasterize + autograd over every tile, float64, torch CPU"}), flush=True)


if __name__ == "__main__":
    main()
