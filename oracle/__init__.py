"""CPU oracle (float64 PyTorch) of the CasualHDRSplat image-formation hot path.

TEST INFRASTRUCTURE ONLY: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package.  The product
(``casualhdrsplat_b200``) never does, and fails loudly if its CUDA library is missing.

PARITY UNPINNED: the reference ships no code, tests or golden vectors (see oracle/formation.py).
"""
from . import se3  # noqa: F401
from .formation import (  # noqa: F401
    CRF_IDENTITY, CRF_LUT, CRF_MLP, TIGHT_MARGIN, bin_tiles, bin_tiles_fused, blend, blend_pixel_loop, crf_apply, formation, key_bits, project,
    rasterize, tile_bounds, tile_grid,
)
