"""Generates the committed golden vectors from the float64 oracle (run from the repo root:
``python tests/golden/make_golden.py``).  The reference ships no fixtures, so these pin the oracle
itself against drift between rounds and give the GPU tests a target that
does not need the oracle at full size.

  c1_oracle.pt   BASELINE.json configs[0]: 10k Gaussians, 256x256, 1 camera, 1 pose, exposure 1,
                 identity CRF: LDR image (fp32), M, gradients on a fixed 512-row subset.
  tiny_oracle.pt the 'tiny' test scene: LDR, alpha, sorted keys / vals / tile offsets, all gradients.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from casualhdrsplat_b200.scene import make_config  # noqa: E402
from tests.util import oracle_run  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sc = make_config("c1")
    ldr, alpha, meta, grads = oracle_run(sc)
    idx = torch.randperm(sc.means.shape[0], generator=torch.Generator().manual_seed(123))[:512].sort().values
    torch.save({"ldr": ldr.float(), "alpha_sum": float(alpha.sum()), "n_isect": meta["n_isect"], "grad_index": idx,
                "grads": {k: grads[k][idx].float() for k in ["means", "quats", "scales", "opacities", "colors"]},
                "grad_norms": {k: float(v.norm()) for k, v in grads.items()}}, os.path.join(HERE, "c1_oracle.pt"))
    sc = make_config("tiny")
    ldr, alpha, meta, grads = oracle_run(sc)
    b = meta["bins"]
    torch.save({"ldr": ldr.float(), "alpha": alpha.float(), "n_isect": meta["n_isect"], "keys_sorted": b["keys_sorted"],
                "vals_sorted": b["vals_sorted"], "tile_offsets": b["tile_offsets"],
                "grads": {k: v.float() for k, v in grads.items()}}, os.path.join(HERE, "tiny_oracle.pt"))
    print("wrote goldens")


if __name__ == "__main__":
    main()
