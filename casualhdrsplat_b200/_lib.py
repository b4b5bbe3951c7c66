"""ctypes binding of libchs.so (the C ABI declared in include/chs.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C casualhdrsplat_b200/csrc``.
There is no fallback: if the shared library is missing, :func:`lib` raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int32, c_int64, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libchs.so")

CHS_CRF_IDENTITY, CHS_CRF_MLP, CHS_CRF_LUT = 0, 1, 2
CHS_SPLINE_LINEAR, CHS_SPLINE_CUBIC = 0, 1
CHS_SORT_KEY64, CHS_SORT_DEPTH_PRESORT = 0, 1


class ChsConfig(ctypes.Structure):
    _fields_ = [
        ("n_gauss", c_int32), ("n_frames", c_int32), ("n_virtual", c_int32), ("width", c_int32), ("height", c_int32),
        ("tile_size", c_int32), ("near_plane", c_float), ("far_plane", c_float), ("eps2d", c_float),
        ("crf_kind", c_int32), ("crf_hidden", c_int32), ("crf_before_average", c_int32), ("ks_per_camera", c_int32),
        ("sort_mode", c_int32), ("background", c_float * 3), ("rgbo_per_camera", c_int32), ("tight_bounds", c_int32), ("pose_fused", c_int32),
        ("tune_blend_fwd", c_int32), ("tune_blend_bwd", c_int32), ("tune_crf_bwd", c_int32), ("tune_bin", c_int32),
        ("tune_bin_chunk", c_int32), ("tune_project_bwd", c_int32), ("reserved", c_int32 * 1),
    ]


class ChsTensors(ctypes.Structure):
    """chs_tensors of include/chs.h (one-shot entry points)."""
    _fields_ = ([(n, c_void_p) for n in ["means", "quats", "scales", "opacities", "colors", "Ks", "exposure", "crf_params", "viewmats"]]
                + [("spline_kind", c_int32), ("n_knots", c_int32), ("knots", c_void_p), ("frame_times", c_void_p),
                   ("knot_t0", c_double), ("knot_dt", c_double)]
                + [(n, c_void_p) for n in ["geom", "conic_c", "depths", "rgbo", "radii", "tiles_touched", "order", "vals_sorted", "last_id",
                                           "isect_offsets", "tile_offsets", "n_isect_dev"]]
                + [("isect_capacity", c_int64)]
                + [(n, c_void_p) for n in ["ldr", "alpha", "hdr_mean", "final_T", "v_ldr", "v_alpha", "v_hdr", "v_geom", "v_cogr", "v_blue",
                                           "grads_flat", "v_viewmats", "v_crf_params", "v_exposure", "v_knots", "v_frame_times",
                                           "v_exposure_window", "workspace"]]
                + [("workspace_bytes", c_uint64)])


class ChsWorkspaceSizes(ctypes.Structure):
    _fields_ = [("bin_count_bytes", c_uint64), ("bin_sort_bytes", c_uint64), ("reduce_bytes", c_uint64)]


# name -> (restype, argtypes); every symbol include/chs.h declares
P = c_void_p
CFG = POINTER(ChsConfig)

SIGNATURES = {
    "chs_version": (ctypes.c_int, []),
    "chs_last_error": (c_char_p, []),
    "chs_launch_count": (c_uint64, []),
    "chs_sizeof": (c_uint64, [c_int32]),
    "chs_workspace_query": (ctypes.c_int, [CFG, c_int64, c_int32, POINTER(ChsWorkspaceSizes)]),
    "chs_spline_fwd": (ctypes.c_int, [c_int32, P, c_int32, c_double, c_double, P, P, c_int32, c_int32, P, P]),
    "chs_spline_bwd": (ctypes.c_int, [c_int32, P, c_int32, c_double, c_double, P, P, c_int32, c_int32, P, P, P, P, P, c_uint64, P]),
    "chs_project_fwd": (ctypes.c_int, [CFG] + [P] * 14),
    "chs_project_bwd": (ctypes.c_int, [CFG] + [P] * 12 + [c_uint64, P]),
    "chs_bin_count": (ctypes.c_int, [CFG, P, P, P, P, P, POINTER(c_int64), P, c_uint64, P]),
    "chs_bin_sort": (ctypes.c_int, [CFG, c_int64] + [P] * 9 + [c_uint64, P]),
    "chs_bin_sort_dev": (ctypes.c_int, [CFG, c_int64] + [P] * 10 + [c_uint64, P]),
    "chs_radix_sort_pairs": (ctypes.c_int, [P, c_int32, ctypes.c_uint32, c_int32, P, P, P, c_uint64, P]),
    "chs_bin_emit_keys": (ctypes.c_int, [CFG, c_int64] + [P] * 8),
    "chs_blend_fwd": (ctypes.c_int, [CFG] + [P] * 13),
    "chs_crf_bwd": (ctypes.c_int, [CFG] + [P] * 8 + [c_uint64, P]),
    "chs_blend_bwd": (ctypes.c_int, [CFG] + [P] * 13),
    "chs_comm_unique_id": (ctypes.c_int, [P]),
    "chs_comm_init": (ctypes.c_int, [P, c_int32, c_int32, POINTER(c_void_p)]),
    "chs_allreduce_grads": (ctypes.c_int, [P, P, c_uint64, P]),
    "chs_comm_destroy": (ctypes.c_int, [P]),
    "chs_nvls_allreduce": (ctypes.c_int, [P, c_uint64, c_int32, c_int32, P]),
    "chs_nvls_broadcast": (ctypes.c_int, [P, P, c_uint64, c_uint64, P]),
    "chs_sh_fwd": (ctypes.c_int, [CFG, c_int32, P, P, P, P, P, P]),
    "chs_sh_bwd": (ctypes.c_int, [CFG, c_int32] + [P] * 10 + [c_uint64, P]),
    "chs_loss": (ctypes.c_int, [c_int32, P, P, c_uint64, c_float, P, P, P]),
    "chs_ssim_loss": (ctypes.c_int, [P, P, c_int32, c_int32, c_int32, c_float, c_float, P, P, P, c_uint64, P]),
    "chs_adam_step": (ctypes.c_int, [P, P, P, P, c_uint64, c_float, c_float, c_float, c_float, c_int32, c_float, P]),
    "chs_rasterize_fwd": (ctypes.c_int, [CFG, POINTER(ChsTensors), POINTER(c_int64), P]),
    "chs_rasterize_bwd": (ctypes.c_int, [CFG, POINTER(ChsTensors), c_int64, P]),
}

_LIB = None


def lib() -> ctypes.CDLL:
    """Load libchs.so (once). Raises if it has not been built — there is no CPU path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C casualhdrsplat_b200/csrc`. casualhdrsplat_b200 has no CPU fallback.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        for which, struct in enumerate([ChsConfig, ChsWorkspaceSizes, ChsTensors]):
            if l.chs_sizeof(which) != ctypes.sizeof(struct):
                raise RuntimeError(f"{LIB_PATH}: stale build ({struct.__name__} is {l.chs_sizeof(which)} bytes in the library, "
                                   f"{ctypes.sizeof(struct)} in the binding); rebuild with `make -C casualhdrsplat_b200/csrc`")
        _LIB = l
    return _LIB


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().chs_last_error()
        raise RuntimeError(f"libchs {what} failed with status {status}: {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a (contiguous) tensor, or NULL."""
    if t is None:
        return None
    assert t.is_contiguous(), "libchs needs contiguous tensors"
    return c_void_p(t.data_ptr())


def make_config(n_gauss, n_frames, n_virtual, width, height, *, near=0.01, far=1e10, eps2d=0.3, tile_size=16,
                crf_kind=CHS_CRF_IDENTITY, crf_hidden=0, crf_before_average=False, ks_per_camera=False,
                sort_mode=CHS_SORT_DEPTH_PRESORT, background=None, rgbo_per_camera=False, tight_bounds=False, pose_fused=False,
                tuning=None) -> ChsConfig:
    """``tuning``: optional dict of development knobs {blend_fwd, blend_bwd, crf_bwd, bin, bin_chunk} (chs_config.tune_*)."""
    cfg = ChsConfig()
    cfg.n_gauss, cfg.n_frames, cfg.n_virtual = int(n_gauss), int(n_frames), int(n_virtual)
    cfg.width, cfg.height, cfg.tile_size = int(width), int(height), int(tile_size)
    cfg.near_plane, cfg.far_plane, cfg.eps2d = float(near), float(far), float(eps2d)
    cfg.crf_kind, cfg.crf_hidden = int(crf_kind), int(crf_hidden)
    cfg.crf_before_average = int(bool(crf_before_average))
    cfg.ks_per_camera = int(bool(ks_per_camera))
    cfg.sort_mode = int(sort_mode)
    cfg.rgbo_per_camera = int(bool(rgbo_per_camera))
    cfg.tight_bounds = int(bool(tight_bounds))
    cfg.pose_fused = int(bool(pose_fused))
    for k, v in (tuning or {}).items():
        if not hasattr(cfg, "tune_" + k):
            raise RuntimeError(f"make_config: unknown tuning knob `{k}`")
        setattr(cfg, "tune_" + k, int(v))
    bg = (0.0, 0.0, 0.0) if background is None else tuple(float(v) for v in background)
    cfg.background[0], cfg.background[1], cfg.background[2] = bg
    return cfg


def workspace_sizes(cfg: ChsConfig, n_isect: int = 0, n_knots: int = 0) -> ChsWorkspaceSizes:
    out = ChsWorkspaceSizes()
    check(lib().chs_workspace_query(byref(cfg), int(n_isect), int(n_knots), byref(out)), "chs_workspace_query")
    return out


def crf_size(crf_kind: int, crf_params) -> int:
    """chs_config.crf_hidden for a parameter tensor: Hd of the MLP ([3, 3*Hd+1]) or the knot count of the LUT ([3, L+2])."""
    if crf_kind == CHS_CRF_IDENTITY or crf_params is None:
        return 0
    if crf_params.dim() != 2 or crf_params.shape[0] != 3:
        raise RuntimeError("crf_params must be [3, P]")
    if crf_kind == CHS_CRF_MLP:
        if (crf_params.shape[1] - 1) % 3 != 0 or crf_params.shape[1] < 4:
            raise RuntimeError("crf_params must be [3, 3*Hd+1] for the MLP CRF")
        return (crf_params.shape[1] - 1) // 3
    if crf_kind == CHS_CRF_LUT:
        if crf_params.shape[1] < 4:
            raise RuntimeError("crf_params must be [3, L+2] with L >= 2 for the LUT CRF")
        return crf_params.shape[1] - 2
    raise RuntimeError(f"unknown crf_kind {crf_kind}")
