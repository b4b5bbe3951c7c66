"""Known-answer tests of the pose-fused binning definition (oracle ``pose_fused=True``)."""
import math

import pytest
import torch

import oracle
from casualhdrsplat_b200.scene import make_config
from tests.util import oracle_run, rel


def test_static_camera_fused_equals_per_pose():
    """With identical poses inside the exposure window the union rectangle is each pose's rectangle and the mid-pose depth is
    each pose's depth: the fused lists are the per-pose lists and the frames are the same bits."""
    sc = make_config("tiny", static_camera=True)
    a = oracle_run(sc)
    b = oracle_run(sc, pose_fused=True)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    for k in a[3]:
        assert torch.equal(a[3][k], b[3][k]), k
    n, N = sc.n_virtual, sc.means.shape[0]
    tiles = ((sc.width + 15) // 16) * ((sc.height + 15) // 16)
    fa, fb = a[2]["bins"], b[2]["bins"]
    assert fb["n_isect"] * n == fa["n_isect"]
    for f in range(sc.n_frames):
        for t in range(tiles):
            lf = fb["vals_sorted"][fb["tile_offsets"][f * tiles + t]:fb["tile_offsets"][f * tiles + t + 1]].long() - f * N
            for k in range(n):
                c = f * n + k
                lc = fa["vals_sorted"][fa["tile_offsets"][c * tiles + t]:fa["tile_offsets"][c * tiles + t + 1]].long() - c * N
                assert torch.equal(lf, lc)


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_fused_lists_are_supersets_in_mid_pose_depth_order(name):
    sc = make_config(name)
    a = oracle_run(sc, with_grad=False)
    b = oracle_run(sc, with_grad=False, pose_fused=True)
    n, N = sc.n_virtual, sc.means.shape[0]
    tiles = ((sc.width + 15) // 16) * ((sc.height + 15) // 16)
    fa, fb = a[2]["bins"], b[2]["bins"]
    # n-fold fewer list entries (a little more than 1/n: union rectangles)
    assert fa["n_isect"] / n <= fb["n_isect"] <= 1.6 * fa["n_isect"] / n
    depth = a[2]["proj"]["depths"].float()
    for f in range(sc.n_frames):
        for t in range(0, tiles, 3):
            lf = fb["vals_sorted"][fb["tile_offsets"][f * tiles + t]:fb["tile_offsets"][f * tiles + t + 1]].long() - f * N
            d_mid = depth[f * n + n // 2, lf]
            assert bool((d_mid[1:] >= d_mid[:-1]).all())  # ordered by the depth at the middle pose
            for k in range(n):
                c = f * n + k
                lc = fa["vals_sorted"][fa["tile_offsets"][c * tiles + t]:fa["tile_offsets"][c * tiles + t + 1]].long() - c * N
                assert set(lc.tolist()) <= set(lf.tolist())  # every pose's list is contained in the frame's list
    # the model changes only through depth-order differences between the middle pose and pose k: a small image difference
    e = rel(b[0], a[0])
    assert e < 5e-3, e


def test_fused_matches_literal_pixel_loop():
    """The vectorised fused blend against a literal per-pixel walk of the frame's list with the pose's own projection."""
    sc = make_config("tiny")
    ldr, alpha, meta, _ = oracle_run(sc, with_grad=False, pose_fused=True)
    n, N = sc.n_virtual, sc.means.shape[0]
    tiles_w = (sc.width + 15) // 16
    tiles = tiles_w * ((sc.height + 15) // 16)
    bins, proj = meta["bins"], meta["proj"]
    g = torch.Generator().manual_seed(1)
    for _ in range(40):
        c = int(torch.randint(0, sc.n_frames * n, (1,), generator=g))
        i, j = int(torch.randint(0, sc.height, (1,), generator=g)), int(torch.randint(0, sc.width, (1,), generator=g))
        f = c // n
        tid = (i // 16) * tiles_w + j // 16
        s0, s1 = int(bins["tile_offsets"][f * tiles + tid]), int(bins["tile_offsets"][f * tiles + tid + 1])
        T, acc = 1.0, [0.0, 0.0, 0.0]
        for s in range(s0, s1):
            gid = int(bins["vals_sorted"][s]) - f * N
            if not bool(bins["live"][c, gid]):
                continue
            dx = float(proj["means2d"][c, gid, 0]) - (j + 0.5)
            dy = float(proj["means2d"][c, gid, 1]) - (i + 0.5)
            q = proj["conics"][c, gid]
            sigma = 0.5 * (float(q[0]) * dx * dx + float(q[2]) * dy * dy) + float(q[1]) * dx * dy
            if sigma < 0:
                continue
            a = min(0.999, float(sc.opacities[gid]) * math.exp(-sigma))
            if a < (1.0 / 255.0):
                continue
            if T * (1 - a) <= 1e-4:
                break
            for ch in range(3):
                acc[ch] += a * T * float(sc.colors[gid, ch])
            T *= 1 - a
        got = meta["hdr_cams"][c, i, j]
        for ch in range(3):
            assert abs(float(got[ch]) - acc[ch]) <= 1e-9 * max(1.0, abs(acc[ch]))
        assert abs(float(meta["alpha_cams"][c, i, j]) - (1 - T)) < 1e-12
