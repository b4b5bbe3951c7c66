"""casualhdrsplat_b200 — B200-native image-formation hot path of CasualHDRSplat.

Public surface: :func:`rasterize` (gsplat-style, extended with exposure times, virtual-pose count
and CRF parameters), backed by hand-written sm_100a CUDA kernels behind a C-ABI shared library
(``include/chs.h``).  There is no CPU fallback: calling :func:`rasterize` without the built
library or without a CUDA device raises.
"""
from .scene import CRF_IDENTITY, CRF_LUT, CRF_MLP, SPLINE_CUBIC, SPLINE_LINEAR  # noqa: F401

__version__ = "0.2.0"


def __getattr__(name):
    if name == "rasterize":
        from . import api
        return getattr(api, name)
    raise AttributeError(name)
