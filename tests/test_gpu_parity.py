"""T3 CUDA-vs-oracle parity, through the public rasterize() -> C ABI path.

Bars (BASELINE.json north_star):
  * binning, keys, sort, tile offsets: bit-exact (integer equality) when the oracle's binning is fed
    the kernel's own fp32 (means2d, radii, depths);
  * forward LDR: ||B - B*|| / ||B*|| <= 1e-4;   every gradient tensor: <= 1e-3  (b = float64 oracle
    evaluated on the same fp32 inputs).
"""
import math
import os

import pytest
import torch

import oracle
from casualhdrsplat_b200.scene import SPLINE_CUBIC, SPLINE_LINEAR, make_config, make_scene
from tests.util import cuda_projection, cuda_run, oracle_run, rel, robust_grad_report

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-4
GRAD_TOL = 1e-3


def _u32(t):
    return t.cpu().to(torch.int64) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------
# K0 spline
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["tiny", "c2"])
def test_spline_viewmats(name):
    sc = make_config(name, n_gauss=64)
    _, _, meta, _ = cuda_run(sc, with_grad=False)
    want = oracle.se3.spline_viewmats(sc.knots.double(), sc.knot_t0, sc.knot_dt, sc.frame_times.double(),
                                      sc.exposure_times.double(), sc.n_virtual, sc.spline_kind)
    assert (meta["viewmats"].cpu().double() - want).abs().max() < 2e-7


# ------------------------------------------------------------------------------------------------
# K1 projection
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["small", "c1"])
def test_projection_parity(name):
    sc = make_config(name)
    _, _, meta, _ = cuda_run(sc, with_grad=False)
    proj = cuda_projection(meta)
    ref = oracle.project(sc.means, sc.quats, sc.scales, meta["viewmats"].cpu(), sc.Ks.repeat_interleave(sc.n_virtual, 0),
                         sc.width, sc.height)
    r_cuda, r_ref = proj["radii"], ref["radii"]
    flips = int((r_cuda != r_ref).sum())
    assert flips <= max(2, int(1e-4 * r_ref.numel())), f"{flips} radius flips"
    both = (r_cuda > 0) & (r_ref > 0)
    assert both.sum() > 100
    assert rel(proj["means2d"][both], ref["means2d"][both]) < 1e-6
    assert rel(proj["conics"][both], ref["conics"][both]) < 1e-4
    assert rel(proj["depths"], ref["depths"]) < 1e-6


# ------------------------------------------------------------------------------------------------
# K2-K5 binning: bit exact, both sort strategies
# ------------------------------------------------------------------------------------------------
# sort routes: the literal 64-bit key sort (hand-written LSD passes); depth presort + banded placement (default); depth presort +
# hand-written two-pass radix multisplit (chs_config.tune_bin = 3); the CUB baselines (tune_bin = 1); the round-1 counting
# placement (tune_bin = 2, small chunks so even tiny scenes have many)
BIN_ROUTES = {"key64": ("key64", None), "presort": ("presort", None), "presort-radix": ("presort", {"bin": 3}),
              "presort-cub": ("presort", {"bin": 1}), "presort-place": ("presort", {"bin": 2, "bin_chunk": 96}),
              "key64-cub": ("key64", {"bin": 1})}


@pytest.mark.parametrize("route", list(BIN_ROUTES))
@pytest.mark.parametrize("name", ["tiny", "small", "c1"])
def test_binning_bit_exact(name, route):
    sc = make_config(name)
    sort_mode, tuning = BIN_ROUTES[route]
    _, _, meta, _ = cuda_run(sc, with_grad=False, sort_mode=sort_mode, debug_keys=True, tuning=tuning)
    st = meta["state"]
    proj = cuda_projection(meta)
    b = oracle.bin_tiles(proj["means2d"], proj["radii"], proj["depths"], sc.width, sc.height)
    assert torch.equal(st.tiles_touched.cpu(), b["tiles_touched"])
    assert st.n_isect == b["n_isect"] and st.n_isect > 0
    assert torch.equal(st.keys_sorted.cpu(), b["keys_sorted"])
    assert torch.equal(st.vals_sorted.cpu()[: st.n_isect], b["vals_sorted"])
    assert torch.equal(_u32(st.tile_offsets), b["tile_offsets"])
    if sort_mode == "key64":
        assert torch.equal(_u32(st.isect_offsets), b["offsets"])


def test_binning_depth_ties_and_sort_modes_agree():
    # many exactly equal depths: a plane of Gaussians facing an axis-aligned camera
    sc = make_scene(3000, 160, 96, n_frames=1, n_virtual=1, spline_kind=SPLINE_LINEAR, static_camera=True, scale_mult=6.0,
                    crf_kind=0, unit_exposure=True)
    vm = oracle.se3.spline_viewmats(sc.knots.double(), sc.knot_t0, sc.knot_dt, sc.frame_times.double(), sc.exposure_times.double(), 1, 0)
    R, t = vm[0, :3, :3], vm[0, :3, 3]
    p = sc.means.double() @ R.T + t
    p[:, 2] = torch.where(torch.arange(3000) % 3 == 0, torch.tensor(4.0, dtype=torch.float64), p[:, 2])
    sc.means = ((p - t) @ R).float()
    outs = {}
    for mode in BIN_ROUTES:
        sort_mode, tuning = BIN_ROUTES[mode]
        _, _, meta, _ = cuda_run(sc, with_grad=False, sort_mode=sort_mode, debug_keys=True, tuning=tuning)
        st = meta["state"]
        outs[mode] = (st.keys_sorted.cpu(), st.vals_sorted.cpu()[: st.n_isect], st.tile_offsets.cpu())
        proj = cuda_projection(meta)
        b = oracle.bin_tiles(proj["means2d"], proj["radii"], proj["depths"], sc.width, sc.height)
        ks = b["keys_sorted"]
        assert int((ks[1:] == ks[:-1]).sum()) > 50, "test scene should contain depth ties inside tiles"
        assert torch.equal(outs[mode][0], ks) and torch.equal(outs[mode][1], b["vals_sorted"])
    for other in BIN_ROUTES:
        for a, bb in zip(outs["key64"], outs[other]):
            assert torch.equal(a, bb)


# ------------------------------------------------------------------------------------------------
# K6 forward
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["tiny", "small", "c1"])
def test_forward_parity_same_lists(name):
    """Blend + formation epilogue in isolation: the oracle gets the CUDA projection, so tile lists are identical."""
    sc = make_config(name)
    bg = [0.05, 0.1, 0.2]
    ldr, alpha, meta, _ = cuda_run(sc, with_grad=False, background=bg, return_hdr=True)
    o_ldr, o_alpha, o_meta, _ = oracle_run(sc, with_grad=False, projection_override=cuda_projection(meta), background=bg)
    assert meta["n_isect"] == o_meta["n_isect"]
    assert rel(meta["hdr"], o_meta["hdr_mean"]) <= FWD_TOL
    assert rel(ldr, o_ldr) <= FWD_TOL
    assert rel(alpha, o_alpha) <= FWD_TOL
    st = meta["state"]
    C = st.final_T.shape[0]
    # last_id: CUDA stores 1-based index inside the tile list; oracle the sorted index (-1 = none)
    tiles_w = (sc.width + 15) // 16
    to = _u32(st.tile_offsets)
    yy, xx = torch.meshgrid(torch.arange(sc.height), torch.arange(sc.width), indexing="ij")
    tid = (yy // 16) * tiles_w + (xx // 16)
    n_tiles = tiles_w * ((sc.height + 15) // 16)
    flips = 0
    for c in range(C):
        start = to[c * n_tiles + tid]
        got = st.last_id[c].cpu().long()
        got = torch.where(got > 0, got - 1 + start, torch.full_like(got, -1))
        flips += int((got != o_meta["last_id"][c]).sum())
    assert flips <= 1e-3 * C * sc.width * sc.height, f"{flips} last_id flips"
    assert rel(1 - st.final_T, o_meta["alpha_cams"]) <= FWD_TOL


@pytest.mark.parametrize("name", ["tiny", "small", "c1"])
def test_forward_parity_end_to_end(name):
    sc = make_config(name)
    ldr, alpha, meta, _ = cuda_run(sc, with_grad=False)
    o_ldr, o_alpha, o_meta, _ = oracle_run(sc, with_grad=False)
    assert abs(meta["n_isect"] - o_meta["n_isect"]) <= max(4, 1e-4 * o_meta["n_isect"])
    assert rel(ldr, o_ldr) <= FWD_TOL
    assert rel(alpha, o_alpha) <= FWD_TOL


# ------------------------------------------------------------------------------------------------
# K7-K9 + K0 backward
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["tiny", "small", "c1"])
def test_gradient_parity_end_to_end(name):
    sc = make_config(name)
    g = torch.Generator().manual_seed(11)
    v_alpha = torch.randn(sc.n_frames, sc.height, sc.width, 1, generator=g, dtype=torch.float32)
    _, _, meta, grads = cuda_run(sc, v_alpha=v_alpha)
    _, _, _, o_grads = oracle_run(sc, v_alpha=v_alpha, projection_override=cuda_projection(meta), straight_through=True)
    errs = {k: rel(grads[k], o_grads[k]) for k in grads if float(o_grads[k].norm()) > 0}
    bad = {k: e for k, e in errs.items() if not e <= GRAD_TOL}
    assert not bad, f"gradient rel errors above {GRAD_TOL}: {bad} (all: {errs})"


def test_blend_backward_stage_parity():
    """K8 in isolation: gradients w.r.t. the projected quantities (means2d, conics) and per-Gaussian colour/opacity."""
    sc = make_config("small")
    bg = [0.3, 0.2, 0.1]
    _, _, meta, _ = cuda_run(sc, with_grad=False, background=bg)
    st = meta["state"]
    proj = cuda_projection(meta)
    m2d = proj["means2d"].double().requires_grad_(True)
    con = proj["conics"].double().requires_grad_(True)
    op = sc.opacities.double().requires_grad_(True)
    col = sc.colors.double().requires_grad_(True)
    b = oracle.bin_tiles(proj["means2d"], proj["radii"], proj["depths"], sc.width, sc.height)
    hdr, alpha_c, _ = oracle.blend(m2d, con, op, col, b["vals_sorted"], b["tile_offsets"], sc.means.shape[0], sc.width, sc.height, bg)
    g = torch.Generator().manual_seed(5)
    B, n = sc.n_frames, sc.n_virtual
    v_hdr = torch.randn(B, sc.height, sc.width, 3, generator=g, dtype=torch.float32)
    v_al = torch.randn(B, sc.height, sc.width, generator=g, dtype=torch.float32)
    loss = (hdr.reshape(B, n, sc.height, sc.width, 3) * v_hdr[:, None].double()).sum() \
        + (alpha_c.reshape(B, n, sc.height, sc.width).mean(1) * v_al.double()).sum()
    gm, gc, go, gcol = torch.autograd.grad(loss, [m2d, con, op, col])
    # run K8 directly through the C ABI with these upstream gradients
    from ctypes import byref
    from casualhdrsplat_b200 import _lib, api

    dev = st.geom.device
    C, N = st.radii.shape
    f32 = torch.float32
    v_geom = torch.empty(C, N, 4, device=dev, dtype=f32); v_cogr = torch.empty(C, N, 4, device=dev, dtype=f32); v_blue = torch.empty(C, N, device=dev, dtype=f32)
    _lib.check(_lib.lib().chs_blend_bwd(byref(st.cfg), _lib.ptr(st.geom), _lib.ptr(st.conic_c), _lib.ptr(st.rgbo), _lib.ptr(st.vals_sorted),
                                        _lib.ptr(st.tile_offsets), _lib.ptr(st.final_T), _lib.ptr(st.last_id),
                                        _lib.ptr(v_hdr.to(dev).contiguous()), _lib.ptr(v_al.to(dev).contiguous()), _lib.ptr(v_geom),
                                        _lib.ptr(v_cogr), _lib.ptr(v_blue), api._stream()))
    torch.cuda.synchronize()
    assert rel(v_geom[..., :2], gm) <= GRAD_TOL
    assert rel(torch.cat([v_geom[..., 2:], v_cogr[..., :1]], -1), gc) <= GRAD_TOL
    assert rel(v_cogr[..., 1].sum(0), go) <= GRAD_TOL
    assert rel(torch.cat([v_cogr[..., 2:], v_blue[..., None]], -1).sum(0), gcol) <= GRAD_TOL


# ------------------------------------------------------------------------------------------------
# edge cases
# ------------------------------------------------------------------------------------------------
def test_explicit_viewmats_and_per_camera_Ks():
    from casualhdrsplat_b200 import rasterize

    sc = make_config("tiny")
    dev = torch.device("cuda:0")
    vm = oracle.se3.spline_viewmats(sc.knots.double(), sc.knot_t0, sc.knot_dt, sc.frame_times.double(), sc.exposure_times.double(),
                                    sc.n_virtual, sc.spline_kind).float()
    Kc = sc.Ks.repeat_interleave(sc.n_virtual, 0).clone()
    Kc[:, 0, 2] += torch.arange(Kc.shape[0]) * 0.25  # distinct principal points per virtual camera
    vmd = vm.to(dev).requires_grad_(True)
    ldr, alpha, meta = rasterize(sc.means.to(dev), sc.quats.to(dev), sc.scales.to(dev), sc.opacities.to(dev), sc.colors.to(dev),
                                 vmd, Kc.to(dev), sc.width, sc.height, sc.exposure_times.to(dev), sc.n_virtual, sc.crf_kind,
                                 sc.crf_params.to(dev))
    (gv,) = torch.autograd.grad((ldr * sc.v_ldr.to(dev)).sum(), [vmd])
    vmo = vm.double().requires_grad_(True)
    o_ldr, _, _ = oracle.rasterize(sc.means, sc.quats, sc.scales, sc.opacities, sc.colors, vmo, Kc, sc.width, sc.height,
                                   sc.exposure_times, sc.n_virtual, sc.crf_kind, sc.crf_params)
    (go,) = torch.autograd.grad((o_ldr * sc.v_ldr.double()).sum(), [vmo])
    assert rel(ldr, o_ldr) <= FWD_TOL
    assert rel(gv[:, :3, :], go[:, :3, :]) <= GRAD_TOL
    assert float(gv[:, 3, :].abs().max()) == 0.0


def test_empty_and_fully_culled_scene():
    from casualhdrsplat_b200 import rasterize

    sc = make_config("tiny")
    dev = torch.device("cuda:0")
    means = sc.means.clone()
    means[:, 2] -= 1000.0  # everything behind the camera
    m = means.to(dev).requires_grad_(True)
    ldr, alpha, meta = rasterize(m, sc.quats.to(dev), sc.scales.to(dev), sc.opacities.to(dev), sc.colors.to(dev), None,
                                 sc.Ks.to(dev), sc.width, sc.height, sc.exposure_times.to(dev), sc.n_virtual, 0, None,
                                 spline={k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in sc.spline().items()},
                                 background=[0.5, 0.25, 0.125])
    assert meta["n_isect"] == 0
    want = sc.exposure_times.to(dev)[:, None, None, None] * torch.tensor([0.5, 0.25, 0.125], device=dev, dtype=torch.float32)
    assert torch.allclose(ldr, want.expand_as(ldr), rtol=1e-6)
    assert float(alpha.detach().abs().max()) == 0.0
    (g,) = torch.autograd.grad(ldr.sum(), [m])
    assert float(g.abs().max()) == 0.0


def test_ragged_image_and_single_pose():
    sc = make_scene(1500, 71, 37, n_frames=3, n_virtual=1, spline_kind=SPLINE_CUBIC, crf_hidden=32, scale_mult=10.0)
    ldr, alpha, meta, grads = cuda_run(sc)
    o_ldr, o_alpha, o_meta, o_grads = oracle_run(sc)
    assert rel(ldr, o_ldr) <= FWD_TOL
    errs = {k: rel(grads[k], o_grads[k]) for k in grads if float(o_grads[k].norm()) > 0}
    assert all(e <= GRAD_TOL for e in errs.values()), errs
    # n_virtual = 1: exposure has only the brightness path
    assert float(o_grads["exposure_times"].norm()) > 0


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_figure_order_crf_before_average(name):
    """B = mean_k F(dt * H_k), the order assets/pipeline.png draws."""
    sc = make_config(name)
    g = torch.Generator().manual_seed(12)
    v_alpha = torch.randn(sc.n_frames, sc.height, sc.width, 1, generator=g, dtype=torch.float32)
    ldr, alpha, meta, grads = cuda_run(sc, v_alpha=v_alpha, crf_before_average=True, return_hdr=True)
    o_ldr, o_alpha, o_meta, o_grads = oracle_run(sc, v_alpha=v_alpha, crf_before_average=True,
                                                 projection_override=cuda_projection(meta), straight_through=True)
    d_ldr, _, _, _ = oracle_run(sc, with_grad=False)
    assert rel(o_ldr, d_ldr) > 1e-5, "the two orders must differ for a non-linear CRF"
    assert rel(ldr, o_ldr) <= FWD_TOL and rel(alpha, o_alpha) <= FWD_TOL
    assert rel(meta["hdr"], o_meta["hdr_mean"]) <= FWD_TOL
    errs = {k: rel(grads[k], o_grads[k]) for k in grads if float(o_grads[k].norm()) > 0}
    assert all(e <= GRAD_TOL for e in errs.values()), errs


@pytest.mark.parametrize("figure_order", [False, True])
@pytest.mark.parametrize("name,knots", [("tiny", 24), ("small", 256)])
def test_lut_crf_parity(name, knots, figure_order):
    """the piecewise-linear log-exposure table as the learned CRF (crf_kind = CHS_CRF_LUT)."""
    from casualhdrsplat_b200.scene import CRF_LUT

    sc = make_config(name, crf_kind=CRF_LUT, crf_hidden=knots)
    assert sc.crf_params.shape == (3, knots + 2)
    ldr, alpha, meta, grads = cuda_run(sc, crf_before_average=figure_order)
    o_ldr, o_alpha, o_meta, o_grads = oracle_run(sc, crf_before_average=figure_order, projection_override=cuda_projection(meta),
                                                 straight_through=True)
    assert 0.02 < float(o_ldr.mean()) < 0.98, "the table should be exercised in its sloped part"
    assert rel(ldr, o_ldr) <= FWD_TOL and rel(alpha, o_alpha) <= FWD_TOL
    errs = {k: rel(grads[k], o_grads[k]) for k in grads if float(o_grads[k].norm()) > 0}
    assert all(e <= GRAD_TOL for e in errs.values()), errs
    assert float(o_grads["crf_params"][:, 2:].norm()) > 0
    assert float(grads["crf_params"][:, :2].abs().max()) == 0.0  # the table range is a fixed calibration


@pytest.mark.parametrize("name", ["tiny", "small", "c1"])
def test_tight_bounds_same_images_shorter_lists(name):
    """chs_config.tight_bounds (opacity-aware per-axis bounds): binning stays bit-exact against the oracle's definition,
    the rendered frames are bit-identical to the classic square bounds, and gradients agree with the oracle."""
    sc = make_config(name)
    ldr_sq, alpha_sq, meta_sq, grads_sq = cuda_run(sc)
    ldr, alpha, meta, grads = cuda_run(sc, tight_bounds=True, debug_keys=True)
    st = meta["state"]
    assert st.n_isect < 0.9 * meta_sq["n_isect"]
    # (a) binning bit-exact on the kernel's own fp32 projection (packed radii)
    proj = cuda_projection(meta)
    b = oracle.bin_tiles(proj["means2d"], proj["radii"], proj["depths"], sc.width, sc.height, tight=True)
    assert torch.equal(st.tiles_touched.cpu(), b["tiles_touched"]) and st.n_isect == b["n_isect"]
    assert torch.equal(st.keys_sorted.cpu(), b["keys_sorted"])
    assert torch.equal(st.vals_sorted.cpu()[: st.n_isect], b["vals_sorted"])
    assert torch.equal(_u32(st.tile_offsets), b["tile_offsets"])
    # (b) only dead entries were dropped: the frames are the same bits
    assert torch.equal(ldr, ldr_sq) and torch.equal(alpha, alpha_sq)
    for k in grads:
        assert rel(grads[k], grads_sq[k]) <= 1e-5, k  # atomics order differs, nothing else
    # (c) oracle parity in tight mode
    o_ldr, o_alpha, o_meta, o_grads = oracle_run(sc, projection_override=proj, straight_through=True, tight_bounds=True)
    assert rel(ldr, o_ldr) <= FWD_TOL and rel(alpha, o_alpha) <= FWD_TOL
    errs = {k: rel(grads[k], o_grads[k]) for k in grads if float(o_grads[k].norm()) > 0}
    assert all(e <= GRAD_TOL for e in errs.values()), errs
    # (d) the kernel's packed radii agree with the oracle's own float64 ones up to a pixel at ceil() boundaries
    own = oracle_run(sc, with_grad=False, tight_bounds=True)[2]["proj"]["radii"]
    mine = proj["radii"]
    both = (own != 0) & (mine != 0)
    assert int(((own != 0) != (mine != 0)).sum()) <= 2
    assert int(((own & 0xFFFF) - (mine & 0xFFFF)).abs()[both].max()) <= 1 and int((((own >> 16) & 0xFFFF) - ((mine >> 16) & 0xFFFF)).abs()[both].max()) <= 1
    assert float(((own != mine) & both).float().mean()) < 1e-3


def test_golden_config1_forward():
    """Committed golden of BASELINE.json configs[0] (oracle-generated, tests/golden/make_golden.py)."""
    import os

    path = os.path.join(os.path.dirname(__file__), "golden", "c1_oracle.pt")
    gold = torch.load(path)
    sc = make_config("c1")
    ldr, alpha, meta, grads = cuda_run(sc)
    assert meta["n_isect"] == gold["n_isect"]
    assert rel(ldr, gold["ldr"]) <= FWD_TOL
    for k in ["means", "quats", "scales", "opacities", "colors"]:
        assert rel(grads[k].cpu()[gold["grad_index"]], gold["grads"][k]) <= 2 * GRAD_TOL, k


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[1] at full size, and size-independent properties at configs[2] (1M, 1080p, 8 poses)
# ------------------------------------------------------------------------------------------------
def test_config2_full_parity():
    """100k Gaussians, 800x800, 4 virtual poses on a linear SE(3) trajectory, exposure + learned CRF."""
    sc = make_config("c2")
    ldr, alpha, meta, grads = cuda_run(sc)
    # (1) parity as SURVEY.md A.8 defines it: all discrete decisions (tile lists, alpha >= 1/255, early stop) are taken
    #     on the kernel's own fp32 projection values; gradients flow through the oracle's float64 projection
    o_ldr, o_alpha, o_meta, o_grads = oracle_run(sc, projection_override=cuda_projection(meta), straight_through=True)
    assert meta["n_isect"] == o_meta["n_isect"]
    assert rel(ldr, o_ldr) <= FWD_TOL
    errs = {k: rel(grads[k], o_grads[k]) for k in grads if float(o_grads[k].norm()) > 0}
    assert all(e <= GRAD_TOL for e in errs.values()), errs
    # (2) fully independent oracle (its own fp32 cast of its own projection): a few tile lists may differ by a ceil()
    #     flip or a swap of two nearly equal depths, and boundary pixels of a Gaussian may fall on the other side of
    #     alpha = 1/255; count them, and check the gradients with statistics that a handful of flips cannot dominate
    i_ldr, _, i_meta, i_grads = oracle_run(sc)
    assert abs(meta["n_isect"] - i_meta["n_isect"]) <= 1e-4 * i_meta["n_isect"]
    assert rel(ldr, i_ldr) <= FWD_TOL
    n_rad = int((meta["state"].radii.cpu() != i_meta["proj"]["radii"]).sum())
    assert n_rad <= 1e-4 * i_meta["proj"]["radii"].numel()
    for k in ["means", "quats", "scales", "opacities", "colors"]:
        med, frac_bad = robust_grad_report(grads[k], i_grads[k])
        assert med <= 1e-4 and frac_bad <= 5e-3, (k, med, frac_bad)  # measured: med ~1e-5, frac_bad <= 1.8e-3


@pytest.fixture(scope="module")
def c3_scene():
    return make_config("c3")


def test_config3_binning_properties(c3_scene):
    """At 1M Gaussians the oracle is too slow for a direct comparison; check the invariants that pin the
    result: both sort strategies give identical lists, keys are sorted, offsets are consistent."""
    sc = c3_scene
    out = {}
    for mode in ["presort", "key64"]:
        _, _, meta, _ = cuda_run(sc, with_grad=False, sort_mode=mode, debug_keys=True)
        st = meta["state"]
        M = st.n_isect
        assert int(st.tiles_touched.sum(dtype=torch.int64)) == M
        keys = st.keys_sorted
        assert bool((keys[1:] >= keys[:-1]).all())
        to = st.tile_offsets.to(torch.int64) & 0xFFFFFFFF
        assert bool((to[1:] >= to[:-1]).all()) and int(to[-1]) == M and int(to[0]) == 0
        # every list entry's key names the (camera, tile) bucket the offsets put it in
        tiles = ((sc.width + 15) // 16) * ((sc.height + 15) // 16)
        tile_bits = tiles.bit_length()
        lin = (keys >> (32 + tile_bits)) * tiles + ((keys >> 32) & ((1 << tile_bits) - 1))
        probe = torch.randint(0, M, (200000,), device=keys.device)
        b = lin[probe]
        assert bool(((to[b] <= probe) & (probe < to[b + 1])).all())
        # the key's depth bits are the depth of the Gaussian the value names
        dbits = st.depths.reshape(-1).view(torch.int32)[st.vals_sorted[:M][probe].long()].to(torch.int64) & 0xFFFFFFFF
        assert torch.equal(keys[probe] & 0xFFFFFFFF, dbits)
        out[mode] = (keys.clone(), st.vals_sorted[:M].clone(), st.tile_offsets.clone())
        del meta, st
    for a, b in zip(out["presort"], out["key64"]):
        assert torch.equal(a, b)


def test_config3_tight_bounds_same_frames(c3_scene):
    """Full size (1M Gaussians, 1080p, 8 poses): the opacity-aware bounds the bench runs with render the same bits as the
    classic square bounds from two thirds of the intersections, their lists are sorted, and the gradients agree."""
    sc = c3_scene
    ldr_sq, alpha_sq, meta_sq, grads_sq = cuda_run(sc)
    m_sq = meta_sq["n_isect"]
    del meta_sq
    ldr, alpha, meta, grads = cuda_run(sc, tight_bounds=True, debug_keys=True)
    st = meta["state"]
    M = st.n_isect
    assert 0.5 * m_sq < M < 0.8 * m_sq
    assert torch.equal(ldr, ldr_sq) and torch.equal(alpha, alpha_sq)
    for k in grads:
        assert rel(grads[k], grads_sq[k]) <= 2e-5, (k, rel(grads[k], grads_sq[k]))
    keys = st.keys_sorted
    assert bool((keys[1:] >= keys[:-1]).all())
    to = st.tile_offsets.to(torch.int64) & 0xFFFFFFFF
    assert bool((to[1:] >= to[:-1]).all()) and int(to[-1]) == M and int(st.tiles_touched.sum(dtype=torch.int64)) == M
    # packed radii never exceed the classic radius: every tight rectangle lies inside the square one
    r = st.radii
    assert int((r & 0xFFFF).max()) < 65535 and bool(((r > 0) == (st.tiles_touched > 0))[st.tiles_touched > 0].all())


def test_config3_forward_is_deterministic_and_linear(c3_scene):
    """Identity CRF + explicit view matrices: B is exactly linear in exposure and in the colours, the forward has no
    atomics (bit-reproducible), and the analytic gradients of those two linear maps are known in closed form."""
    from casualhdrsplat_b200 import rasterize

    sc = c3_scene
    dev = torch.device("cuda:0")
    vm = oracle.se3.spline_viewmats(sc.knots.double(), sc.knot_t0, sc.knot_dt, sc.frame_times.double(), sc.exposure_times.double(),
                                    sc.n_virtual, sc.spline_kind).float().to(dev)
    args = [sc.means.to(dev), sc.quats.to(dev), sc.scales.to(dev), sc.opacities.to(dev)]
    col = sc.colors.to(dev).requires_grad_(True)
    dt = sc.exposure_times.to(dev).requires_grad_(True)
    Ks = sc.Ks.to(dev)

    def run(c, e):
        return rasterize(*args, c, vm, Ks, sc.width, sc.height, e, sc.n_virtual, 0, None, return_hdr=True)

    ldr, alpha, meta = run(col, dt)
    ldr2, _, _ = run(col, dt)
    assert torch.equal(ldr, ldr2)
    ldr3, _, _ = run(col.detach(), dt.detach() * 3)
    assert rel(ldr3, 3 * ldr.detach()) < 1e-6
    g = torch.Generator().manual_seed(3)
    v = sc.v_ldr.to(dev)
    g_col, g_dt = torch.autograd.grad((ldr * v).sum(), [col, dt])
    # d/d dt = <v, hdr_mean>
    want_dt = (v * meta["hdr"].detach()).sum(dim=(1, 2, 3))
    assert rel(g_dt, want_dt) < 1e-4
    # linear in colours: L(c + d) - L(c) = <g, d>
    d = torch.randn(col.shape, generator=g).to(dev) * sc.colors.to(dev)
    ldr4, _, _ = run(col.detach() + d, dt.detach())
    lhs = ((ldr4 - ldr.detach()).double() * v.double()).sum()
    rhs = (g_col.double() * d.double()).sum()
    assert abs(float(lhs - rhs)) <= 2e-3 * abs(float(rhs))


def test_one_shot_c_abi_matches_staged_path():
    """chs_rasterize_fwd / chs_rasterize_bwd (caller-owned buffers, capacity-checked M) against rasterize()."""
    import ctypes
    from ctypes import byref, c_int64

    from casualhdrsplat_b200 import _lib, api

    sc = make_config("small")
    ldr, alpha, meta, grads = cuda_run(sc)
    st = meta["state"]
    dev = torch.device("cuda:0")
    f32, i32 = torch.float32, torch.int32
    B, n, N, W, H = sc.n_frames, sc.n_virtual, sc.means.shape[0], sc.width, sc.height
    C, tiles, K = B * n, ((W + 15) // 16) * ((H + 15) // 16), sc.knots.shape[0]
    cfg = st.cfg
    cap = st.n_isect + 1000
    ws0, ws1 = _lib.workspace_sizes(cfg, 0, K), _lib.workspace_sizes(cfg, cap, K)
    wbytes = max(int(ws0.bin_count_bytes), int(ws1.bin_sort_bytes), int(ws0.reduce_bytes))
    bufs = {}

    def mk(name, shape, dtype=f32):
        bufs[name] = torch.empty(shape, dtype=dtype, device=dev)
        return ctypes.c_void_p(bufs[name].data_ptr())

    inp = {k: getattr(sc, k).to(dev).contiguous() for k in ["means", "quats", "scales", "opacities", "colors", "Ks", "exposure_times",
                                                              "crf_params", "knots", "frame_times", "v_ldr"]}
    t = _lib.ChsTensors()
    for k in ["means", "quats", "scales", "opacities", "colors", "Ks", "crf_params", "knots", "frame_times", "v_ldr"]:
        setattr(t, k, inp[k].data_ptr())
    t.exposure = inp["exposure_times"].data_ptr()
    t.spline_kind, t.n_knots, t.knot_t0, t.knot_dt = sc.spline_kind, K, sc.knot_t0, sc.knot_dt
    t.viewmats = mk("viewmats", (C, 4, 4))
    t.geom, t.conic_c, t.depths, t.rgbo = mk("geom", (C, N, 4)), mk("conic_c", (C, N)), mk("depths", (C, N)), mk("rgbo", (N, 4))
    t.radii, t.tiles_touched, t.order = mk("radii", (C, N), i32), mk("tiles", (C, N), i32), mk("order", (C * N,), i32)
    t.vals_sorted, t.last_id = mk("vals", (cap,), i32), mk("last_id", (C, H, W), i32)
    t.isect_offsets, t.tile_offsets = mk("offs", (C * N,), i32), mk("to", (C * tiles + 1,), i32)
    t.n_isect_dev, t.isect_capacity = mk("nd", (1,), torch.int64), cap
    t.ldr, t.alpha, t.hdr_mean, t.final_T = mk("ldr", (B, H, W, 3)), mk("alpha", (B, H, W)), mk("hdr", (B, H, W, 3)), mk("T", (C, H, W))
    t.v_alpha = None
    t.v_hdr, t.v_geom, t.v_cogr, t.v_blue = mk("v_hdr", (B, H, W, 3)), mk("v_geom", (C, N, 4)), mk("v_cogr", (C, N, 4)), mk("v_blue", (C, N))
    t.grads_flat, t.v_viewmats = mk("flat", (14 * N,)), mk("v_vm", (C, 4, 4))
    t.v_crf_params, t.v_exposure = mk("v_crf", tuple(sc.crf_params.shape)), mk("v_ex", (B,))
    t.v_knots, t.v_frame_times, t.v_exposure_window = mk("v_knots", (K, 7)), mk("v_ft", (B,)), mk("v_exw", (B,))
    t.workspace, t.workspace_bytes = mk("work", (wbytes,), torch.uint8), wbytes
    M = c_int64(0)
    L = _lib.lib()
    _lib.check(L.chs_rasterize_fwd(byref(cfg), byref(t), byref(M), api._stream()), "chs_rasterize_fwd")
    _lib.check(L.chs_rasterize_bwd(byref(cfg), byref(t), M.value, api._stream()), "chs_rasterize_bwd")
    torch.cuda.synchronize()
    assert M.value == st.n_isect
    assert torch.equal(bufs["ldr"], ldr) and torch.equal(bufs["alpha"], alpha[..., 0])
    vm, vq, vs, vo, vc = api.split_flat_grads(bufs["flat"], N)
    for got, name in [(vm, "means"), (vq, "quats"), (vs, "scales"), (vo, "opacities"), (vc, "colors"), (bufs["v_knots"], "knots"),
                      (bufs["v_ex"], "exposure_times"), (bufs["v_ft"], "frame_times"), (bufs["v_crf"], "crf_params")]:
        assert rel(got, grads[name]) < 1e-5, name
    # capacity check: too small a buffer is reported, with the required count
    t.isect_capacity = 10
    M2 = c_int64(0)
    assert L.chs_rasterize_fwd(byref(cfg), byref(t), byref(M2), api._stream()) == -4
    assert M2.value == st.n_isect and b"isect_capacity" in L.chs_last_error()


def _stress_scene(n, w, h, frames, n_virtual, seed):
    """Odd sizes, screen-filling and sub-pixel Gaussians, opacities below 1/255 and above the 0.999 clamp,
    exact duplicates (position and depth ties), Gaussians behind / beside the camera."""
    sc = make_scene(n, w, h, n_frames=frames, n_virtual=n_virtual, spline_kind=SPLINE_CUBIC, crf_hidden=16, scale_mult=10.0,
                    scene_seed=seed)
    g = torch.Generator().manual_seed(seed + 100)
    k = n // 8
    sc.scales[:k] *= 40.0                      # screen-filling
    sc.scales[k:2 * k] *= 0.02                 # far below one pixel
    sc.opacities[2 * k:2 * k + k // 2] = 0.003  # < 1/255: can never contribute
    sc.opacities[2 * k + k // 2:3 * k] = 0.9996  # above the clamp
    sc.means[3 * k:4 * k] = sc.means[3 * k:3 * k + 1]  # exact duplicates -> equal depths in every camera
    sc.means[4 * k:4 * k + k // 2, 2] -= 40.0   # behind the camera
    sc.means[4 * k + k // 2:5 * k, 0] += 500.0  # far off to the side
    sc.quats[5 * k:6 * k] *= 7.5               # un-normalised quaternions are allowed
    return sc


@pytest.mark.parametrize("n,w,h,frames,n_virtual", [(1237, 33, 17, 1, 2), (801, 5, 3, 2, 1), (2049, 130, 70, 1, 3)])
def test_stress_edge_cases(n, w, h, frames, n_virtual):
    sc = _stress_scene(n, w, h, frames, n_virtual, seed=n)
    for mode in ["presort", "key64"]:
        ldr, alpha, meta, grads = cuda_run(sc, sort_mode=mode, debug_keys=True, background=[0.2, 0.1, 0.05])
        st = meta["state"]
        proj = cuda_projection(meta)
        b = oracle.bin_tiles(proj["means2d"], proj["radii"], proj["depths"], w, h)
        assert st.n_isect == b["n_isect"]
        assert torch.equal(st.keys_sorted.cpu(), b["keys_sorted"]) and torch.equal(st.vals_sorted.cpu()[: st.n_isect], b["vals_sorted"])
        assert torch.equal(_u32(st.tile_offsets), b["tile_offsets"])
        o_ldr, o_alpha, o_meta, o_grads = oracle_run(sc, projection_override=proj, straight_through=True, background=[0.2, 0.1, 0.05])
        assert rel(ldr, o_ldr) <= FWD_TOL and rel(alpha, o_alpha) <= FWD_TOL
        errs = {k: rel(grads[k], o_grads[k]) for k in grads if float(o_grads[k].norm()) > 0}
        assert all(e <= GRAD_TOL for e in errs.values()), (mode, errs)


def test_fused_loss_and_adam_match_torch():
    """chs_loss (+ dL/dB in the same pass) and chs_adam_step against torch."""
    from casualhdrsplat_b200.train import LOSS_L1, LOSS_L2, FlatAdam, photometric_loss

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(21)
    ldr = torch.rand(2, 37, 53, 3, generator=g, dtype=torch.float32).to(dev)
    tgt = torch.rand(2, 37, 53, 3, generator=g, dtype=torch.float32).to(dev)
    for kind, fn in [(LOSS_L2, lambda d: 0.5 * 0.25 * (d * d).sum()), (LOSS_L1, lambda d: 0.25 * d.abs().sum())]:
        x = ldr.clone().double().requires_grad_(True)
        want = fn(x - tgt.double())
        (gx,) = torch.autograd.grad(want, x)
        v, acc = photometric_loss(ldr, tgt, kind, scale=0.25)
        assert abs(float(acc) - float(want.detach())) <= 1e-6 * abs(float(want.detach()))
        assert rel(v, gx) < 1e-6
    p0 = torch.randn(1001, 3, generator=g, dtype=torch.float32).to(dev)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=3e-3, betas=(0.9, 0.99), eps=1e-8)
    mine = {"w": p0.clone()}
    fa = FlatAdam(mine, lr=3e-3, betas=(0.9, 0.99), eps=1e-8)
    for it in range(5):
        gr = torch.randn(1001, 3, generator=g, dtype=torch.float32).to(dev)
        p_ref.grad = gr.clone() * 0.5
        opt.step()
        fa.step({"w": gr}, grad_scale=0.5)
    assert rel(mine["w"], p_ref.detach()) < 1e-6


@pytest.mark.parametrize("shape", [(2, 37, 53), (1, 16, 16), (3, 5, 70), (1, 270, 481)])
def test_ssim_loss_matches_oracle(shape):
    """the D-SSIM loss and its gradient (chs_ssim_loss) against oracle/ssim.py."""
    from casualhdrsplat_b200.train import ssim_loss
    from oracle import ssim as ossim

    n, h, w = shape
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(31 + h)
    tgt = torch.rand(n, h, w, 3, generator=g, dtype=torch.float32)
    # smooth structure + noise so that both the luminance and the contrast terms matter
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    tgt = (0.5 + 0.3 * torch.sin(0.31 * xx + 0.17 * yy)[None, :, :, None] + 0.15 * (tgt - 0.5)).clamp(0, 1)
    ldr = (tgt + 0.1 * torch.randn(n, h, w, 3, generator=g, dtype=torch.float32)).clamp(0, 1)
    for l1_w, ssim_w in [(0.8, 0.2), (0.0, 1.0)]:
        x = ldr.double().requires_grad_(True)
        want = ossim.ssim_loss(x, tgt.double(), l1_w, ssim_w)
        (gx,) = torch.autograd.grad(want, x)
        v, acc = ssim_loss(ldr.to(dev), tgt.to(dev), l1_w, ssim_w)
        assert abs(float(acc) - float(want.detach())) <= 2e-5 * abs(float(want.detach())), (float(acc), float(want.detach()))
        assert rel(v, gx) <= GRAD_TOL, rel(v, gx)
    # identical frames: SSIM = 1 everywhere, the loss vanishes
    v, acc = ssim_loss(tgt.to(dev), tgt.to(dev), 0.8, 0.2)
    assert abs(float(acc)) < 1e-6


@pytest.mark.parametrize("figure_order", [False, True])
def test_gradient_through_returned_hdr(figure_order):
    """return_hdr=True (README R9: 'render HDR and LDR images'): a loss on the pose-averaged HDR image back-propagates too."""
    from casualhdrsplat_b200 import rasterize

    sc = make_config("tiny")
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(31)
    w_h = torch.randn(sc.n_frames, sc.height, sc.width, 3, generator=g, dtype=torch.float32)
    col = sc.colors.to(dev).requires_grad_(True)
    op = sc.opacities.to(dev).requires_grad_(True)
    sp = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in sc.spline().items()}
    ldr, alpha, meta = rasterize(sc.means.to(dev), sc.quats.to(dev), sc.scales.to(dev), op, col, None, sc.Ks.to(dev), sc.width, sc.height,
                                 sc.exposure_times.to(dev), sc.n_virtual, sc.crf_kind, sc.crf_params.to(dev), spline=sp,
                                 return_hdr=True, crf_before_average=figure_order)
    gc, go = torch.autograd.grad((meta["hdr"] * w_h.to(dev)).sum() + (ldr * sc.v_ldr.to(dev)).sum(), [col, op])
    oc = sc.colors.double().requires_grad_(True)
    oo = sc.opacities.double().requires_grad_(True)
    o_ldr, _, o_meta = oracle.rasterize(sc.means, sc.quats, sc.scales, oo, oc, None, sc.Ks, sc.width, sc.height, sc.exposure_times,
                                        sc.n_virtual, sc.crf_kind, sc.crf_params, spline=sc.spline(), crf_before_average=figure_order,
                                        projection_override=cuda_projection(meta))
    ogc, ogo = torch.autograd.grad((o_meta["hdr_mean"] * w_h.double()).sum() + (o_ldr * sc.v_ldr.double()).sum(), [oc, oo])
    assert rel(meta["hdr"], o_meta["hdr_mean"]) <= FWD_TOL
    assert rel(gc, ogc) <= GRAD_TOL and rel(go, ogo) <= GRAD_TOL


@pytest.mark.parametrize("deg,name", [(0, "tiny"), (2, "tiny"), (3, "small")])
def test_sh_view_dependent_colour(deg, name):
    """colours from spherical harmonics per virtual camera (chs_sh_fwd / chs_sh_bwd)."""
    sc = make_config(name)
    g = torch.Generator().manual_seed(50 + deg)
    K = (deg + 1) ** 2
    sh = torch.randn(sc.means.shape[0], K, 3, generator=g, dtype=torch.float32) * 0.5
    sh[:, 0] += 1.5
    ldr, alpha, meta, grads = cuda_run(sc, sh=sh)
    o_ldr, o_alpha, o_meta, o_grads = oracle_run(sc, sh=sh, projection_override=cuda_projection(meta), straight_through=True)
    assert rel(ldr, o_ldr) <= FWD_TOL and rel(alpha, o_alpha) <= FWD_TOL
    assert float(o_grads["sh_coeffs"].norm()) > 0
    skip = {"colors"}  # unused with SH
    errs = {k: rel(grads[k], o_grads[k]) for k in grads if k not in skip and float(o_grads[k].norm()) > 0}
    assert all(e <= GRAD_TOL for e in errs.values()), errs
    assert float(grads["colors"].abs().max()) == 0.0


def test_tight_bounds_keep_splats_with_half_extent_above_32767():
    """a near-camera floater whose tight vertical half extent is >= 32768 px packs to an int32 with bit
    31 set.  Every consumer gates on != 0, so tight and square bounds render the same frame and the same gradients, and the
    binning stays bit-exact against the oracle."""
    from casualhdrsplat_b200 import rasterize
    from tests.util import huge_splat_scene

    kw = huge_splat_scene()
    dev = torch.device("cuda:0")
    names = ["means", "quats", "scales", "opacities", "colors"]

    def run(tight):
        args = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
        for k in names:
            args[k] = args[k].requires_grad_(True)
        ldr, alpha, meta = rasterize(**args, tight_bounds=tight, debug_keys=True)
        g = torch.autograd.grad((ldr * torch.linspace(0.5, 1.5, ldr.numel(), device=dev).view_as(ldr)).sum(), [args[k] for k in names])
        return ldr.detach(), alpha.detach(), meta, g

    ldr_sq, alpha_sq, meta_sq, g_sq = run(False)
    ldr, alpha, meta, g = run(True)
    st = meta["state"]
    packed = int(st.radii[0, 0])
    assert packed < 0 and (packed >> 16) & 0xFFFF >= 32768
    tiles = ((kw["width"] + 15) // 16) * ((kw["height"] + 15) // 16)
    assert int(st.tiles_touched[0, 0]) == tiles
    assert torch.equal(ldr, ldr_sq) and torch.equal(alpha, alpha_sq)
    for a, b in zip(g, g_sq):
        assert rel(a, b) <= 1e-5
    assert float(g[0][0].abs().sum()) > 0  # the floater receives a gradient in tight mode
    proj = cuda_projection(meta)
    b = oracle.bin_tiles(proj["means2d"], proj["radii"], proj["depths"], kw["width"], kw["height"], tight=True)
    assert st.n_isect == b["n_isect"] and torch.equal(st.vals_sorted.cpu()[: st.n_isect], b["vals_sorted"])
    assert torch.equal(_u32(st.tile_offsets), b["tile_offsets"])
    o_ldr, _, _ = oracle.rasterize(**kw, tight_bounds=True, projection_override=proj)
    assert rel(ldr, o_ldr) <= FWD_TOL


# ------------------------------------------------------------------------------------------------
# pose-fused binning: one tile list per (frame, tile)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tight", [False, True], ids=["square", "tight"])
@pytest.mark.parametrize("name", ["tiny", "small", "c2"])
def test_pose_fused_binning_and_parity(name, tight):
    """chs_config.pose_fused against the oracle's definition (oracle.bin_tiles_fused / rasterize(pose_fused=True)): the fused
    lists are bit-exact on the kernel's own fp32 projection, forward and every gradient inside the usual tolerances; and the
    flagged model stays close to the per-pose model (the difference is depth-order swaps between poses millimetres apart)."""
    sc = make_config(name)
    ldr, alpha, meta, grads = cuda_run(sc, pose_fused=True, tight_bounds=tight, debug_keys=True)
    st = meta["state"]
    proj = cuda_projection(meta)
    # culled (camera, Gaussian) pairs carry an "alpha = 0 everywhere" record; the oracle needs none (it skips them through `live`)
    b = oracle.bin_tiles_fused(proj["means2d"], proj["radii"], proj["depths"], sc.width, sc.height, sc.n_virtual, tight=tight)
    assert st.n_isect == b["n_isect"] and st.n_isect > 0
    assert torch.equal(st.tiles_touched.cpu(), b["tiles_touched"])
    assert torch.equal(st.vals_sorted.cpu()[: st.n_isect], b["vals_sorted"])
    assert torch.equal(_u32(st.tile_offsets), b["tile_offsets"])
    assert torch.equal(st.keys_sorted.cpu()[: st.n_isect], b["keys_sorted"])
    o_ldr, o_alpha, o_meta, o_grads = oracle_run(sc, projection_override=proj, straight_through=True, tight_bounds=tight, pose_fused=True)
    assert rel(ldr, o_ldr) <= FWD_TOL and rel(alpha, o_alpha) <= FWD_TOL
    errs = {k: rel(grads[k], o_grads[k]) for k in grads if float(o_grads[k].norm()) > 0}
    assert all(e <= GRAD_TOL for e in errs.values()), errs
    # against the per-pose model: n-fold fewer list entries, nearly the same frames
    ldr_pp, _, meta_pp, _ = cuda_run(sc, with_grad=False, tight_bounds=tight)
    n = sc.n_virtual
    assert meta_pp["n_isect"] / n <= st.n_isect <= 1.6 * meta_pp["n_isect"] / n
    assert rel(ldr, ldr_pp) < 5e-3, rel(ldr, ldr_pp)


def test_pose_fused_static_camera_is_the_per_pose_model():
    """Identical poses inside the exposure window: the fused lists are the per-pose lists, so the frames are the same bits."""
    sc = make_config("small", static_camera=True)
    a = cuda_run(sc)
    b = cuda_run(sc, pose_fused=True)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    for k in a[3]:
        assert rel(b[3][k], a[3][k]) <= 1e-5, k  # atomics order only


def test_graphed_step_matches_eager_and_survives_parameter_updates():
    """the sync-free step is capturable in a CUDA graph.  A replay equals the eager step (atomics order
    aside), follows in-place parameter updates, and its intersection counts are checked after the fact."""
    from casualhdrsplat_b200.parallel import GraphedStep, StepState, formation_step

    sc = make_scene(6000, 192, 128, n_frames=3, n_virtual=3, crf_hidden=32, scale_mult=5.0)
    dev = torch.device("cuda", 0)
    names = ["means", "quats", "scales", "opacities", "colors", "knots", "frame_times", "exposure_times", "Ks", "crf_params"]
    P = {k: getattr(sc, k).to(dev).contiguous() for k in names}
    v = sc.v_ldr.to(dev)
    idx = {i: torch.tensor([i], device=dev) for i in range(sc.n_frames)}
    up = lambda f, ldr: torch.cat([v.index_select(0, idx[i]) for i in f])  # noqa: E731  (device-only: capturable)
    meta = {"knot_t0": sc.knot_t0, "knot_dt": sc.knot_dt, "kind": sc.spline_kind}
    args = (P, meta, sc.width, sc.height, sc.n_virtual, sc.crf_kind, list(range(sc.n_frames)), up)
    kw = dict(micro_batch=1, tight_bounds=True)
    gs = GraphedStep(*args, **kw)
    for trial in range(2):
        lay, flat = gs.replay()
        gs.verify()
        got = flat.clone()
        _, ref = formation_step(*args, state=StepState(), **kw)
        torch.cuda.synchronize()
        for name, view in lay.views(got).items():
            want = lay.views(ref)[name]
            err = float((view - want).norm() / want.norm().clamp(min=1e-30))
            assert err < 1e-5, (trial, name, err)
        assert float(got.abs().sum()) > 0
        P["means"].add_(0.003 * torch.randn_like(P["means"]))  # an optimiser step: same storage, new values
        P["colors"].mul_(1.01)


KERNEL_VARIANTS = {
    "bwd_tensor_core": {"blend_bwd": 47},       # blend_bwd4_kernel: phase B as mma.sync over fp16 hi + lo tables
    "bwd_tensor_core_8": {"blend_bwd": 48},
    "bwd_16_rows": {"blend_bwd": 65},           # blend_bwd5_kernel with 16 table rows: two Gaussians per phase-B lane
    "bwd_batch_256": {"blend_bwd": 76},         # blend_bwd5_kernel with 256-entry batches (two list entries per thread)
    "bwd_unstaged": {"blend_bwd": 3},           # blend_bwd3_kernel: one Gaussian's chain after the other (the default until r3c)
    "fwd_batch_128": {"blend_fwd": 57},         # grouped forward with 128-entry batches (one list entry per thread)
    "fwd_ungrouped": {"blend_fwd": 27},         # the pair loop without the speculative group of four
    "scatter_warp_per_tile": {"bin_chunk": 1},  # P2 scatter with a warp per tile
    "crf_per_unit": {"crf_bwd": 14},            # crf_bwd_kernel: the MLP CRF backward unit by unit (the default walks intervals)
}


@pytest.mark.parametrize("variant", sorted(KERNEL_VARIANTS))
@pytest.mark.parametrize("name", ["tiny", "small"])
def test_opt_in_kernel_variants_match_the_oracle(name, variant):
    """Every kernel instantiation kept behind a chs_config.tune_* knob (measured alternatives of the defaults) meets the same
    bars: forward 1e-4, every gradient 1e-3, identical intersection count."""
    sc = make_config(name)
    g = torch.Generator().manual_seed(11)
    v_alpha = torch.randn(sc.n_frames, sc.height, sc.width, 1, generator=g, dtype=torch.float32)
    ldr, _, meta, grads = cuda_run(sc, v_alpha=v_alpha, tuning=KERNEL_VARIANTS[variant], tight_bounds=True)
    o_ldr, _, o_meta, o_grads = oracle_run(sc, v_alpha=v_alpha, projection_override=cuda_projection(meta), straight_through=True,
                                           tight_bounds=True)
    assert rel(ldr, o_ldr) <= FWD_TOL
    errs = {k: rel(grads[k], o_grads[k]) for k in grads if float(o_grads[k].norm()) > 0}
    bad = {k: e for k, e in errs.items() if not e <= GRAD_TOL}
    assert not bad, f"{variant}: gradient rel errors above {GRAD_TOL}: {bad} (all: {errs})"
    ldr0, _, meta0, grads0 = cuda_run(sc, v_alpha=v_alpha, tight_bounds=True)
    assert meta0["n_isect"] == meta["n_isect"]
    if variant != "bwd_tensor_core" and variant != "bwd_tensor_core_8":
        assert torch.equal(ldr0, ldr)  # forward and binning variants are bit-identical to the defaults
