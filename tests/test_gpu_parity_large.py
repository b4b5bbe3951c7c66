"""Oracle parity at the headline sizes (BASELINE.json configs[2] "c3": 1M Gaussians, 1080p, 8 poses; one frame of configs[4]
"c5": 3M Gaussians, 4K, 16 poses), where the full float64 oracle would take days.

  * binning: per camera, ``oracle.bin_tiles`` on the kernel's own fp32 projection over ALL Gaussians -> ``torch.equal`` on the
    camera's slice of vals_sorted / tile_offsets (both sort modes, square and tight bounds);
  * forward + every gradient: the tile-subset recipe validated on the CPU in tests/test_oracle_subset.py — choose random tiles,
    zero the upstream gradient outside them ON BOTH SIDES, run the oracle on exactly the Gaussians those tiles list.  Then the
    oracle's gradient equals the CUDA path's full backward: rows outside the subset must be exactly zero, the others within 1e-3,
    the chosen tiles' LDR pixels within 1e-4.
"""
import dataclasses

import pytest
import torch

import oracle
from casualhdrsplat_b200.scene import make_config
from tests.util import (cuda_projection, cuda_run, gaussians_of_tiles, oracle_run, rel, subset_projection, subset_scene,
                        tile_pixel_mask)

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-4
GRAD_TOL = 1e-3
GAUSS_KEYS = ("means", "quats", "scales", "opacities", "colors")


def _u32(t):
    return t.cpu().to(torch.int64) & 0xFFFFFFFF


def _check_camera_binning(st, proj, sc, cams, tight):
    """Bit-exact comparison of the listed cameras' tile lists with the oracle's binning of the kernel's fp32 projection."""
    N = sc.means.shape[0]
    tiles = ((sc.width + 15) // 16) * ((sc.height + 15) // 16)
    to = _u32(st.tile_offsets)
    vals = st.vals_sorted.cpu().to(torch.int64)
    touched = st.tiles_touched.cpu()
    for c in cams:
        b = oracle.bin_tiles(proj["means2d"][c:c + 1], proj["radii"][c:c + 1], proj["depths"][c:c + 1], sc.width, sc.height, tight=tight)
        lo, hi = int(to[c * tiles]), int(to[(c + 1) * tiles])
        assert hi - lo == b["n_isect"], (c, hi - lo, b["n_isect"])
        assert torch.equal(touched[c], b["tiles_touched"][0]), c
        assert torch.equal(to[c * tiles:(c + 1) * tiles + 1] - lo, b["tile_offsets"]), c
        assert torch.equal(vals[lo:hi] - c * N, b["vals_sorted"].to(torch.int64)), c
        del b


def _subset_parity(sc, n_tiles, tight, seed, check_cams, fused=False):
    """Forward + gradient parity on `n_tiles` random tiles of frame 0 (all its virtual poses); see the module docstring."""
    N = sc.means.shape[0]
    C = sc.n_frames * sc.n_virtual
    tile_w, tile_h = (sc.width + 15) // 16, (sc.height + 15) // 16
    tiles = tile_w * tile_h
    g = torch.Generator().manual_seed(seed)
    pick = sorted(torch.randperm(tiles, generator=g)[:n_tiles].tolist())
    mask = tile_pixel_mask(sc.width, sc.height, pick)
    sc_m = dataclasses.replace(sc, v_ldr=sc.v_ldr * mask[None, :, :, None])
    ldr, alpha, meta, grads = cuda_run(sc_m, tight_bounds=tight, pose_fused=fused)
    st = meta["state"]
    proj = cuda_projection(meta)
    _check_camera_binning(st, proj, sc, check_cams, tight)
    L = sc.n_frames if fused else C  # number of tile-list owners: frames with pose_fused, cameras otherwise
    g_idx = gaussians_of_tiles(st.vals_sorted[: st.n_isect], st.tile_offsets, N, L, tiles, pick)
    assert 0 < g_idx.numel() < N
    sub = subset_scene(sc_m, g_idx)
    o_ldr, o_alpha, o_meta, o_grads = oracle_run(sub, projection_override=subset_projection(proj, g_idx), straight_through=True,
                                                 tight_bounds=tight, pose_fused=fused, tile_subset=[(c, t) for c in range(C) for t in pick])
    # the oracle blended the same lists
    to_f, to_s = _u32(st.tile_offsets).tolist(), o_meta["bins"]["tile_offsets"].tolist()
    vals = st.vals_sorted.cpu().to(torch.int64)
    n_list = 0
    for c in range(L):
        for t in pick:
            a = vals[to_f[c * tiles + t]:to_f[c * tiles + t + 1]] - c * N
            b = o_meta["bins"]["vals_sorted"][to_s[c * tiles + t]:to_s[c * tiles + t + 1]].long() - c * g_idx.numel()
            assert torch.equal(a, g_idx[b]), (c, t)
            n_list += a.numel()
    e_fwd = rel(ldr.cpu()[:, mask], o_ldr[:, mask])
    e_alpha = rel(alpha.cpu()[:, mask], o_alpha[:, mask])
    assert e_fwd <= FWD_TOL and e_alpha <= FWD_TOL, (e_fwd, e_alpha)
    errs = {}
    rest = torch.ones(N, dtype=torch.bool)
    rest[g_idx] = False
    for k in grads:
        mine = grads[k].cpu()
        if k in GAUSS_KEYS:
            assert float(mine[rest].abs().max()) == 0.0, f"{k}: gradient outside the chosen tiles' Gaussians"
            mine = mine[g_idx]
        if float(o_grads[k].norm()) > 0:
            errs[k] = rel(mine, o_grads[k])
    assert set(GAUSS_KEYS) <= set(errs) and all(e <= GRAD_TOL for e in errs.values()), errs
    return {"tiles": len(pick), "gaussians": int(g_idx.numel()), "list_entries": n_list, "fwd": e_fwd, "alpha": e_alpha, "grads": errs,
            "n_isect": st.n_isect}


@pytest.fixture(scope="module")
def c3_scene():
    return make_config("c3")


@pytest.mark.parametrize("tight", [True, False], ids=["tight", "square"])
def test_config3_binning_bit_exact_against_oracle(c3_scene, tight):
    """Every camera of the headline configuration, all 1M Gaussians, both sort modes: the lists are the oracle's lists."""
    sc = c3_scene
    C = sc.n_frames * sc.n_virtual
    ref = None
    for mode in ["presort", "key64"]:
        _, _, meta, _ = cuda_run(sc, with_grad=False, sort_mode=mode, tight_bounds=tight)
        st = meta["state"]
        if ref is None:
            _check_camera_binning(st, cuda_projection(meta), sc, range(C), tight)
            ref = (st.vals_sorted[: st.n_isect].clone(), st.tile_offsets.clone())
        else:  # same projection kernel, so the second sort mode only has to reproduce the lists already checked
            assert torch.equal(st.vals_sorted[: st.n_isect], ref[0]) and torch.equal(st.tile_offsets, ref[1])
        del meta, st


@pytest.mark.parametrize("tight", [True, False], ids=["tight", "square"])
def test_config3_forward_and_gradients_on_tile_subset(c3_scene, tight):
    """64 random tiles x 8 poses of the headline configuration against oracle.blend(tile_subset) + formation + autograd."""
    r = _subset_parity(c3_scene, 64, tight, seed=7, check_cams=[])
    print("c3 subset parity", "tight" if tight else "square", r)


def test_config5_one_frame_against_oracle():
    """One frame of BASELINE.json configs[4] (3M Gaussians, 3840x2160, 16 virtual poses): binning of two cameras bit-exact over
    all 3M Gaussians, forward and gradients on 32 random tiles x 16 poses."""
    sc = make_config("c5", n_frames=1)
    r = _subset_parity(sc, 32, True, seed=9, check_cams=[0, 11])
    print("c5 subset parity", r)


def test_config3_pose_fused_against_oracle_and_per_pose_model(c3_scene):
    """chs_config.pose_fused at the headline size: 64-tile subset parity against the oracle's pose_fused definition, and the
    distance of the flagged model from the per-pose model on the whole 1080p frame."""
    r = _subset_parity(c3_scene, 64, True, seed=7, check_cams=[], fused=True)
    print("c3 pose_fused subset parity", r)
    ldr_f, _, meta_f, _ = cuda_run(c3_scene, with_grad=False, tight_bounds=True, pose_fused=True)
    m_f = meta_f["n_isect"]
    del meta_f
    ldr_p, _, meta_p, _ = cuda_run(c3_scene, with_grad=False, tight_bounds=True)
    e = rel(ldr_f, ldr_p)
    print("c3 pose_fused vs per-pose model: rel", e, "list entries", m_f, "vs", meta_p["n_isect"])
    assert e < 5e-3 and m_f < 0.2 * meta_p["n_isect"]
