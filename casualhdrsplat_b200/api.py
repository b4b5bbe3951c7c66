"""Host-side operator: ``rasterize(...)`` — the gsplat-style entry point BASELINE.json's north_star
names, extended with exposure times, the virtual-pose count and CRF parameters, differentiable
w.r.t. Gaussian, pose (view matrices or spline knots / frame times), exposure and CRF parameters.

The reference repository defines no operator interface for this path (it ships no code:
``/root/reference/Readme.md:57``); the signature is decision D1 of SURVEY.md section 8(b).  Every
stage runs in libchs.so (hand-written sm_100a CUDA behind the C ABI of ``include/chs.h``); PyTorch
only owns device memory, streams and autograd bookkeeping.  No CPU path exists: CPU tensors or a
missing library raise.

Stages per call:
    K0 spline -> K1 project -> K2 count (one D2H of M) -> K3-K5 keys/sort/offsets -> K6 blend+epilogue
backward:
    K7 crf_bwd -> K8 blend_bwd -> K9 project_bwd -> K0 spline_bwd
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_int64
from typing import Optional

import torch

from . import _lib
from ._lib import check, ptr

_SORT_MODES = {"key64": _lib.CHS_SORT_KEY64, "presort": _lib.CHS_SORT_DEPTH_PRESORT}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"rasterize: `{name}` must be a CUDA tensor (casualhdrsplat_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"rasterize: `{name}` must be float32, got {t.dtype}")
    return t.contiguous()


def _empty(shape, dtype, device):
    return torch.empty(shape, dtype=dtype, device=device)


class _State:
    """Everything one forward produced that the backward (and tests) need."""
    __slots__ = ("cfg", "spline", "viewmats", "Ks", "geom", "conic_c", "depths", "radii", "tiles_touched", "rgbo",
                 "isect_offsets", "order", "_n_isect", "_n_pinned", "_n_event", "isect_capacity", "keys_sorted", "vals_sorted",
                 "tile_offsets", "ldr", "alpha", "hdr_mean", "final_T", "last_id", "n_knots", "sh", "sh_degree", "rgbo_c")

    @property
    def n_isect(self) -> int:
        """M.  In the sync-free mode (``isect_capacity`` given) the count travels to the host asynchronously and the first
        read waits for that copy only."""
        if self._n_isect is None:
            self._n_event.synchronize()
            self._n_isect = int(self._n_pinned.item())
        return self._n_isect

    @property
    def overflowed(self) -> bool:
        """Sync-free mode: did the frame need more intersections than the buffers held (its lists were then truncated)?"""
        return self.isect_capacity is not None and self.n_isect > self.isect_capacity


class BufferPool:
    """Stage buffers allocated once per (name, shape, dtype) and reused by later calls ON THE SAME STREAM (a training step
    runs forward and backward of one frame batch before the next one starts, so stream order makes the reuse safe)."""

    def __init__(self):
        self._bufs = {}

    def get(self, name, shape, dtype, device):
        key = (name, tuple(shape), dtype, str(device))
        t = self._bufs.get(key)
        if t is None:
            t = self._bufs[key] = torch.empty(tuple(shape), dtype=dtype, device=device)
        return t

    def get_pinned(self, name, shape, dtype):
        """A pinned host buffer, allocated once per name (CUDA-graph capture cannot allocate pinned memory)."""
        key = ("pinned", name, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = self._bufs[key] = torch.empty(tuple(shape), dtype=dtype).pin_memory()
        return t

    def bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._bufs.values() if t.is_cuda)


def forward_stages(means, quats, scales, opacities, colors, viewmats, Ks, exposure, crf_params, cfg, spline=None,
                   want_keys=False, sh=None, sh_degree=0, isect_capacity=None, pool=None, pinned_name=None) -> _State:
    """Run K0-K6 through the C ABI. All tensors CUDA fp32 contiguous. Returns the stage buffers.
    With ``sh`` [N,K,3] the colours are view dependent (chs_sh_fwd, per-camera records) and ``colors`` is ignored.
    ``isect_capacity``: size the intersection buffers for that many entries and never synchronise with the host (K2's count
    stays on the device, chs_bin_sort_dev reads it there; ``st.n_isect`` / ``st.overflowed`` resolve it lazily).
    ``pool``: a BufferPool to take the stage buffers from instead of allocating them.  ``pinned_name``: take the pinned
    landing buffer of M from the pool under that name (required inside a CUDA-graph capture, see parallel.GraphedStep)."""
    L = _lib.lib()
    dev = means.device

    def _empty(shape, dtype, device, name=None, _n=[0]):  # noqa: B006  (call counter names anonymous buffers in call order)
        _n[0] += 1
        if pool is None:
            return torch.empty(shape, dtype=dtype, device=device)
        return pool.get(name or f"fwd{_n[0]}", shape, dtype, device)

    st = _State()
    st.isect_capacity = None if isect_capacity is None else int(isect_capacity)
    st.cfg, st.spline, st.Ks = cfg, spline, Ks
    N, B, n = cfg.n_gauss, cfg.n_frames, cfg.n_virtual
    C = B * n
    Cb = B if cfg.pose_fused else C  # cameras of the binning stage: with pose_fused the tile lists are per frame
    W, H = cfg.width, cfg.height
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    s = _stream()
    st.n_knots = 0
    if spline is not None:
        knots, knot_t0, knot_dt, frame_times, kind = spline
        st.n_knots = knots.shape[0]
        viewmats = _empty((C, 4, 4), torch.float32, dev)
        check(L.chs_spline_fwd(kind, ptr(knots), knots.shape[0], c_double(knot_t0), c_double(knot_dt), ptr(frame_times),
                               ptr(exposure), B, n, ptr(viewmats), s), "chs_spline_fwd")
    st.viewmats = viewmats
    st.sh, st.sh_degree, st.rgbo_c = sh, int(sh_degree), None
    if sh is not None:
        cfg.rgbo_per_camera = 1
        st.rgbo_c = _empty((C, N, 4), torch.float32, dev)
        check(L.chs_sh_fwd(byref(cfg), st.sh_degree, ptr(sh), ptr(means), ptr(opacities), ptr(viewmats), ptr(st.rgbo_c), s), "chs_sh_fwd")
        if colors is None:
            colors = torch.zeros((N, 3), dtype=torch.float32, device=dev)
    # K1
    st.geom = _empty((C, N, 4), torch.float32, dev)
    st.conic_c = _empty((C, N), torch.float32, dev)
    st.depths = _empty((C, N), torch.float32, dev)
    st.radii = _empty((C, N), torch.int32, dev)
    st.tiles_touched = _empty((Cb, N), torch.int32, dev)
    st.rgbo = _empty((N, 4), torch.float32, dev)
    check(L.chs_project_fwd(byref(cfg), ptr(means), ptr(quats), ptr(scales), ptr(opacities), ptr(colors), ptr(viewmats), ptr(Ks),
                            ptr(st.geom), ptr(st.conic_c), ptr(st.depths), ptr(st.radii), ptr(st.tiles_touched), ptr(st.rgbo), s),
          "chs_project_fwd")
    # K2 (the one host sync of the step: M sizes the intersection buffers)
    ws = _lib.workspace_sizes(cfg, 0, st.n_knots)
    work = _empty((max(int(ws.bin_count_bytes), 256),), torch.uint8, dev)
    st.isect_offsets = _empty((Cb * N,), torch.int32, dev)
    st.order = _empty((Cb * N,), torch.int32, dev) if cfg.sort_mode == _lib.CHS_SORT_DEPTH_PRESORT else None
    n_dev = _empty((1,), torch.int64, dev)
    st._n_pinned = st._n_event = None
    if st.isect_capacity is None:
        n_host = c_int64(0)
        check(L.chs_bin_count(byref(cfg), ptr(st.tiles_touched), ptr(st.depths), ptr(st.isect_offsets), ptr(st.order), ptr(n_dev),
                              byref(n_host), ptr(work), work.numel(), s), "chs_bin_count")
        M = int(n_host.value)
        st._n_isect = M
    else:  # sync-free: M stays on the device; a pinned copy follows asynchronously for the caller's overflow check
        check(L.chs_bin_count(byref(cfg), ptr(st.tiles_touched), ptr(st.depths), ptr(st.isect_offsets), ptr(st.order), ptr(n_dev),
                              None, ptr(work), work.numel(), s), "chs_bin_count")
        M = st.isect_capacity
        st._n_isect = None
        capturing = torch.cuda.is_current_stream_capturing()
        # (under CUDA-graph capture the copy becomes a node of the graph and lands in a buffer that outlives it)
        st._n_pinned = pool.get_pinned(pinned_name or "n_isect", (1,), torch.int64) if (pool is not None and (capturing or pinned_name)) \
            else torch.empty((1,), dtype=torch.int64).pin_memory()
        st._n_pinned.copy_(n_dev, non_blocking=True)
        if not capturing:
            st._n_event = torch.cuda.Event()
            st._n_event.record()
    # K3-K5
    ws = _lib.workspace_sizes(cfg, M, st.n_knots)
    work = _empty((max(int(ws.bin_sort_bytes), 256),), torch.uint8, dev)
    st.keys_sorted = _empty((M,), torch.int64, dev) if want_keys else None
    st.vals_sorted = _empty((max(M, 1),), torch.int32, dev)
    st.tile_offsets = _empty((Cb * tiles + 1,), torch.int32, dev)
    if st.isect_capacity is None:
        check(L.chs_bin_sort(byref(cfg), M, ptr(st.geom), ptr(st.radii), ptr(st.depths), ptr(st.isect_offsets), ptr(st.order),
                             ptr(st.keys_sorted), ptr(st.vals_sorted), ptr(st.tile_offsets), ptr(work), work.numel(), s), "chs_bin_sort")
    else:
        check(L.chs_bin_sort_dev(byref(cfg), M, ptr(n_dev), ptr(st.geom), ptr(st.radii), ptr(st.depths), ptr(st.isect_offsets),
                                 ptr(st.order), ptr(st.keys_sorted), ptr(st.vals_sorted), ptr(st.tile_offsets), ptr(work), work.numel(), s),
              "chs_bin_sort_dev")
    del work
    # K6
    st.ldr = _empty((B, H, W, 3), torch.float32, dev)
    st.alpha = _empty((B, H, W), torch.float32, dev)
    # default order: pose-averaged HDR per frame; crf_before_average (figure order): one HDR image per virtual pose
    st.hdr_mean = _empty((C if cfg.crf_before_average else B, H, W, 3), torch.float32, dev)
    st.final_T = _empty((C, H, W), torch.float32, dev)
    st.last_id = _empty((C, H, W), torch.int32, dev)
    check(L.chs_blend_fwd(byref(cfg), ptr(st.geom), ptr(st.conic_c), ptr(st.rgbo_c if st.rgbo_c is not None else st.rgbo),
                          ptr(st.vals_sorted), ptr(st.tile_offsets),
                          ptr(exposure), ptr(crf_params), ptr(st.ldr), ptr(st.alpha), ptr(st.hdr_mean), ptr(st.final_T),
                          ptr(st.last_id), s), "chs_blend_fwd")
    return st


def backward_stages(st: _State, means, quats, scales, exposure, crf_params, v_ldr, v_alpha, v_hdr_out=None, grads_out=None, pool=None):
    """Run K7-K9 (+K0 bwd). Returns dict of gradients; ``grads_flat`` is the [14N] buffer that multi-GPU runs all-reduce.
    ``pool``: a BufferPool for the backward's stage buffers (the returned gradient tensors then alias pool memory)."""
    L = _lib.lib()
    cfg = st.cfg
    dev = means.device

    def _empty(shape, dtype, device, _n=[0]):  # noqa: B006
        _n[0] += 1
        if pool is None:
            return torch.empty(shape, dtype=dtype, device=device)
        return pool.get(f"bwd{_n[0]}", shape, dtype, device)

    N, B, n = cfg.n_gauss, cfg.n_frames, cfg.n_virtual
    C = B * n
    s = _stream()
    ws = _lib.workspace_sizes(cfg, 0, st.n_knots)
    red = _empty((int(ws.reduce_bytes),), torch.uint8, dev)
    # K7
    v_hdr = _empty((C if cfg.crf_before_average else B, cfg.height, cfg.width, 3), torch.float32, dev)
    v_crf = _empty(tuple(crf_params.shape), torch.float32, dev) if crf_params is not None else None  # overwritten by chs_crf_bwd
    v_exposure = _empty((B,), torch.float32, dev)
    check(L.chs_crf_bwd(byref(cfg), ptr(st.hdr_mean), ptr(exposure), ptr(crf_params), ptr(v_ldr), ptr(v_hdr), ptr(v_crf),
                        ptr(v_exposure), ptr(red), red.numel(), s), "chs_crf_bwd")
    if v_hdr_out is not None:  # gradient arriving at the returned pose-averaged HDR image (return_hdr=True)
        extra = v_hdr_out / float(n)
        v_hdr = v_hdr + (extra.repeat_interleave(n, dim=0) if cfg.crf_before_average else extra)
    # K8
    v_geom = _empty((C, N, 4), torch.float32, dev)
    v_cogr = _empty((C, N, 4), torch.float32, dev)
    v_blue = _empty((C, N), torch.float32, dev)
    check(L.chs_blend_bwd(byref(cfg), ptr(st.geom), ptr(st.conic_c), ptr(st.rgbo_c if st.rgbo_c is not None else st.rgbo),
                          ptr(st.vals_sorted), ptr(st.tile_offsets),
                          ptr(st.final_T), ptr(st.last_id), ptr(v_hdr), ptr(v_alpha), ptr(v_geom), ptr(v_cogr), ptr(v_blue), s),
          "chs_blend_bwd")
    # K9
    grads_flat = grads_out if grads_out is not None else _empty((14 * N,), torch.float32, dev)  # grads_out: caller's [14N] slice
    v_viewmats = _empty((C, 4, 4), torch.float32, dev)
    check(L.chs_project_bwd(byref(cfg), ptr(means), ptr(quats), ptr(scales), ptr(st.viewmats), ptr(st.Ks), ptr(st.radii), ptr(v_geom),
                            ptr(v_cogr), ptr(v_blue), ptr(grads_flat), ptr(v_viewmats), ptr(red), red.numel(), s), "chs_project_bwd")
    v_sh = None
    if st.sh is not None:  # view-dependent colour: coefficient gradients + the view-direction path to means / poses
        v_sh = _empty(tuple(st.sh.shape), torch.float32, dev)
        v_means_sh = _empty((N, 3), torch.float32, dev)
        v_vm_sh = _empty((C, 4, 4), torch.float32, dev)
        check(L.chs_sh_bwd(byref(cfg), st.sh_degree, ptr(st.sh), ptr(means), ptr(st.viewmats), ptr(st.radii), ptr(v_cogr), ptr(v_blue),
                           ptr(v_sh), ptr(v_means_sh), ptr(v_vm_sh), ptr(red), red.numel(), s), "chs_sh_bwd")
        grads_flat[:3 * N].add_(v_means_sh.view(-1))
        v_viewmats.add_(v_vm_sh)
    out = {"grads_flat": grads_flat, "v_viewmats": v_viewmats, "v_crf": v_crf, "v_exposure": v_exposure, "v_sh": v_sh,
           "v_knots": None, "v_frame_times": None, "v_geom": v_geom, "v_cogr": v_cogr, "v_blue": v_blue, "v_hdr": v_hdr}
    if st.spline is not None:
        knots, knot_t0, knot_dt, frame_times, kind = st.spline
        v_knots = _empty(tuple(knots.shape), torch.float32, dev)
        v_ft = _empty((B,), torch.float32, dev)
        v_ex_win = _empty((B,), torch.float32, dev)
        check(L.chs_spline_bwd(kind, ptr(knots), knots.shape[0], c_double(knot_t0), c_double(knot_dt), ptr(frame_times), ptr(exposure),
                               B, n, ptr(v_viewmats), ptr(v_knots), ptr(v_ft), ptr(v_ex_win), ptr(red), red.numel(), s),
              "chs_spline_bwd")
        out["v_knots"], out["v_frame_times"] = v_knots, v_ft
        out["v_exposure"] = v_exposure + v_ex_win  # brightness path + sampling-window path
    return out


def split_flat_grads(grads_flat: torch.Tensor, N: int):
    """Views of the flat [14N] gradient buffer: means [N,3], quats [N,4], scales [N,3], opacities [N], colors [N,3]."""
    return (grads_flat[0:3 * N].view(N, 3), grads_flat[3 * N:7 * N].view(N, 4), grads_flat[7 * N:10 * N].view(N, 3),
            grads_flat[10 * N:11 * N], grads_flat[11 * N:14 * N].view(N, 3))


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, quats, scales, opacities, colors, pose, frame_times, Ks, exposure, crf_params, sh, opts):
        cfg = opts["cfg"]
        spline = None
        viewmats = pose
        if opts["spline_kind"] is not None:
            spline = (pose, opts["knot_t0"], opts["knot_dt"], frame_times, opts["spline_kind"])
            viewmats = None
        with torch.cuda.device(means.device):  # the C ABI launches on the current device
            st = forward_stages(means, quats, scales, opacities, colors, viewmats, Ks, exposure, crf_params, cfg, spline,
                                want_keys=opts["want_keys"], sh=sh, sh_degree=opts["sh_degree"])
        ctx.st = st
        ctx.opts = opts
        ctx.save_for_backward(means, quats, scales, exposure, crf_params if crf_params is not None else torch.empty(0))
        ctx.has_crf = crf_params is not None
        opts["state_out"].append(st)
        ctx.set_materialize_grads(False)
        hdr = st.hdr_mean
        if cfg.crf_before_average:
            hdr = hdr.view(cfg.n_frames, cfg.n_virtual, cfg.height, cfg.width, 3).mean(dim=1)
        return st.ldr, st.alpha.unsqueeze(-1), hdr

    @staticmethod
    def backward(ctx, v_ldr, v_alpha, v_hdr_out):
        means, quats, scales, exposure, crf_params = ctx.saved_tensors
        if not ctx.has_crf:
            crf_params = None
        st = ctx.st
        N = st.cfg.n_gauss
        v_ldr = v_ldr.contiguous() if v_ldr is not None else torch.zeros_like(st.ldr)
        v_alpha = v_alpha.contiguous().view(st.alpha.shape) if v_alpha is not None else None
        with torch.cuda.device(means.device):
            g = backward_stages(st, means, quats, scales, exposure, crf_params, v_ldr, v_alpha, v_hdr_out)
        hook = ctx.opts.get("grad_hook")
        if hook is not None:  # multi-GPU: all-reduce the flat Gaussian gradient buffer (+ small tails) in place
            hook(g)
        vm, vq, vs, vo, vc = split_flat_grads(g["grads_flat"], N)
        if st.spline is not None:
            v_pose, v_ft = g["v_knots"], g["v_frame_times"]
        else:
            v_pose, v_ft = g["v_viewmats"], None
        if st.sh is not None:
            vc = None  # the per-Gaussian colour input is unused with SH
        return vm, vq, vs, vo, vc, v_pose, v_ft, None, g["v_exposure"], g["v_crf"], g["v_sh"], None


def rasterize(means, quats, scales, opacities, colors, viewmats=None, Ks=None, width=0, height=0, exposure_times=None,
              n_virtual=1, crf_kind=_lib.CHS_CRF_IDENTITY, crf_params=None, *, spline=None, background=None, near=0.01,
              far=1e10, eps2d=0.3, tile_size=16, crf_before_average=False, return_hdr=False, sort_mode="presort", tight_bounds=False,
              debug_keys=False, grad_hook=None, sh_coeffs=None, sh_degree=None, pose_fused=False, tuning=None):
    """Render the blurred LDR frames ``B_i = F_theta(dt_i * mean_k H_{i,k})`` and make them differentiable.

    Args (all tensors CUDA float32):
        means [N,3], quats [N,4] (wxyz, un-normalised ok), scales [N,3] (>0), opacities [N] in (0,1),
        colors [N,3] linear HDR radiance.
        viewmats [C,4,4] world-to-camera with C = B*n_virtual and camera c = i*n_virtual + k, **or**
        spline = dict(knots [K,7] camera-to-world (t, q wxyz), knot_t0, knot_dt, frame_times [B], kind 0|1):
        the virtual poses are then sampled at t_i + (k/(n-1) - 1/2) * exposure_times[i].
        Ks [B,3,3] or [C,3,3]; width, height; exposure_times [B]; n_virtual;
        crf_kind 0 = identity, 1 = MLP with crf_params [3, 3*Hd+1] = [w1|b1|w2|b2] per channel, 2 = piecewise-linear
        log-exposure table with crf_params [3, L+2] = [z_min|z_max|v_0..v_{L-1}].
        tight_bounds: bin with opacity-aware per-axis bounds (the box of the alpha >= 1/255 ellipse inside the classic
        3-sigma square) — identical images and gradients, about a third fewer intersections; meta["state"].radii then
        holds packed rx | ry << 16.
        pose_fused: the n virtual poses of a frame share one tile list per (frame, tile) (a flagged
        variant of the model, see chs_config.pose_fused): n-fold less binning work, per-pose projection and alpha tests kept.
        tuning: dict of development knobs (chs_config.tune_*: blend_fwd, blend_bwd, crf_bwd, bin, bin_chunk).
        sh_coeffs [N,K,3] (+ sh_degree <= 3, K >= (deg+1)^2 read as the first coefficients): view-dependent HDR colour
        max(0, 0.5 + sum_k sh_k Y_k(view direction)) evaluated per virtual camera; `colors` may then be None.
    Returns:
        ldr [B,H,W,3], alpha [B,H,W,1], meta (dict: n_isect, viewmats, hdr (if return_hdr), state).
    """
    means = _f32(means, "means")
    dev = means.device
    quats, scales = _f32(quats, "quats"), _f32(scales, "scales")
    opacities = _f32(opacities, "opacities")
    sh = None
    if sh_coeffs is not None:
        sh = _f32(sh_coeffs, "sh_coeffs")
        deg = int(sh_degree) if sh_degree is not None else int(round(sh.shape[1] ** 0.5)) - 1
        if sh.dim() != 3 or sh.shape[0] != means.shape[0] or sh.shape[2] != 3 or not (0 <= deg <= 3) or sh.shape[1] != (deg + 1) ** 2:
            raise RuntimeError("rasterize: sh_coeffs must be [N, (sh_degree+1)^2, 3] with sh_degree in 0..3")
        sh_degree = deg
        colors = torch.zeros((means.shape[0], 3), dtype=torch.float32, device=dev)
    colors = _f32(colors, "colors")
    exposure = _f32(exposure_times, "exposure_times")
    Ks = _f32(Ks, "Ks")
    B = exposure.shape[0]
    C = B * int(n_virtual)
    N = means.shape[0]
    if quats.shape != (N, 4) or scales.shape != (N, 3) or opacities.shape != (N,) or colors.shape != (N, 3):
        raise RuntimeError("rasterize: inconsistent Gaussian tensor shapes")
    if Ks.shape[0] not in (B, C) or tuple(Ks.shape[1:]) != (3, 3):
        raise RuntimeError(f"rasterize: Ks must be [B,3,3] or [C,3,3], got {tuple(Ks.shape)}")
    ks_per_camera = Ks.shape[0] == C and C != B
    if crf_kind in (_lib.CHS_CRF_MLP, _lib.CHS_CRF_LUT):
        crf_params = _f32(crf_params, "crf_params")
    elif crf_kind == _lib.CHS_CRF_IDENTITY:
        crf_params = None
    else:
        raise RuntimeError(f"rasterize: unknown crf_kind {crf_kind}")
    crf_hidden = _lib.crf_size(crf_kind, crf_params)
    if sort_mode not in _SORT_MODES:
        raise RuntimeError(f"rasterize: sort_mode must be one of {sorted(_SORT_MODES)}")
    cfg = _lib.make_config(N, B, n_virtual, width, height, near=near, far=far, eps2d=eps2d, tile_size=tile_size,
                           crf_kind=crf_kind, crf_hidden=crf_hidden, crf_before_average=crf_before_average,
                           ks_per_camera=ks_per_camera, sort_mode=_SORT_MODES[sort_mode], background=background,
                           tight_bounds=tight_bounds, pose_fused=pose_fused, tuning=tuning)
    opts = {"cfg": cfg, "spline_kind": None, "knot_t0": 0.0, "knot_dt": 1.0, "want_keys": bool(debug_keys), "state_out": [],
            "grad_hook": grad_hook, "sh_degree": int(sh_degree) if sh is not None else 0}
    if spline is not None:
        if viewmats is not None:
            raise RuntimeError("rasterize: pass either viewmats or spline, not both")
        pose = _f32(spline["knots"], "spline.knots")
        frame_times = _f32(spline["frame_times"], "spline.frame_times")
        if pose.dim() != 2 or pose.shape[1] != 7 or frame_times.shape != (B,):
            raise RuntimeError("rasterize: spline knots must be [K,7] and frame_times [B]")
        opts["spline_kind"] = int(spline["kind"])
        opts["knot_t0"], opts["knot_dt"] = float(spline["knot_t0"]), float(spline["knot_dt"])
    else:
        pose = _f32(viewmats, "viewmats")
        if tuple(pose.shape) != (C, 4, 4):
            raise RuntimeError(f"rasterize: viewmats must be [B*n_virtual,4,4] = [{C},4,4], got {tuple(pose.shape)}")
        frame_times = None
    ldr, alpha, hdr_mean = _Rasterize.apply(means, quats, scales, opacities, colors, pose, frame_times, Ks, exposure, crf_params, sh,
                                            opts)
    st = opts["state_out"][0]
    meta = {"n_isect": st.n_isect, "viewmats": st.viewmats, "state": st}
    if return_hdr:
        meta["hdr"] = hdr_mean
    return ldr, alpha, meta
