"""T6: the oracle must keep reproducing its committed goldens (guards drift of the parity definition)."""
import os

import torch

from casualhdrsplat_b200.scene import make_config
from tests.util import oracle_run, rel

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_reproduces_tiny_golden():
    gold = torch.load(os.path.join(HERE, "golden", "tiny_oracle.pt"))
    ldr, alpha, meta, grads = oracle_run(make_config("tiny"))
    assert meta["n_isect"] == gold["n_isect"]
    b = meta["bins"]
    assert torch.equal(b["keys_sorted"], gold["keys_sorted"]) and torch.equal(b["vals_sorted"], gold["vals_sorted"])
    assert torch.equal(b["tile_offsets"], gold["tile_offsets"])
    assert rel(ldr, gold["ldr"]) < 1e-6 and rel(alpha, gold["alpha"]) < 1e-6
    for k, v in gold["grads"].items():
        assert rel(grads[k], v) < 1e-5, k


def test_oracle_reproduces_c1_golden():
    gold = torch.load(os.path.join(HERE, "golden", "c1_oracle.pt"))
    ldr, alpha, meta, grads = oracle_run(make_config("c1"))
    assert meta["n_isect"] == gold["n_isect"]
    assert rel(ldr, gold["ldr"]) < 1e-6
    for k, v in gold["grads"].items():
        assert rel(grads[k][gold["grad_index"]], v) < 1e-5, k
