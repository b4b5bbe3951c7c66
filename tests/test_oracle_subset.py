"""The tile-subset recipe used for oracle parity at the headline sizes (tests/test_gpu_parity_large.py), validated on the CPU:
restricting the oracle to the Gaussians that appear in the chosen tiles' lists, with the upstream gradient zeroed outside those
tiles, reproduces the full oracle's pixels on the chosen tiles and its gradients exactly (rows outside the subset are zero)."""
import dataclasses

import torch

from casualhdrsplat_b200.scene import make_config
from tests.util import gaussians_of_tiles, oracle_run, rel, subset_scene, tile_pixel_mask


def test_subset_oracle_equals_full_oracle_on_chosen_tiles():
    sc = make_config("small")
    tile_w, tile_h = (sc.width + 15) // 16, (sc.height + 15) // 16
    tiles = tile_w * tile_h
    C = sc.n_frames * sc.n_virtual
    g = torch.Generator().manual_seed(4)
    pick = sorted(torch.randperm(tiles, generator=g)[:9].tolist())
    mask = tile_pixel_mask(sc.width, sc.height, pick)
    sc_m = dataclasses.replace(sc, v_ldr=sc.v_ldr * mask[None, :, :, None])
    ldr, alpha, meta, grads = oracle_run(sc_m)
    N = sc.means.shape[0]
    g_idx = gaussians_of_tiles(meta["bins"]["vals_sorted"], meta["bins"]["tile_offsets"], N, C, tiles, pick)
    assert 0 < g_idx.numel() < N
    sub = subset_scene(sc_m, g_idx)
    s_ldr, s_alpha, s_meta, s_grads = oracle_run(sub, tile_subset=[(c, t) for c in range(C) for t in pick])
    # the chosen tiles' lists are the same lists (ids renumbered)
    to_f, to_s = meta["bins"]["tile_offsets"].tolist(), s_meta["bins"]["tile_offsets"].tolist()
    for c in range(C):
        for t in pick:
            a = meta["bins"]["vals_sorted"][to_f[c * tiles + t]:to_f[c * tiles + t + 1]].long() - c * N
            b = s_meta["bins"]["vals_sorted"][to_s[c * tiles + t]:to_s[c * tiles + t + 1]].long() - c * g_idx.numel()
            assert torch.equal(a, g_idx[b])
    assert rel(s_ldr[:, mask], ldr[:, mask]) < 1e-12 and rel(s_alpha[:, mask], alpha[:, mask]) < 1e-12
    for k in grads:
        full = grads[k]
        if k in ("means", "quats", "scales", "opacities", "colors"):
            rest = torch.ones(N, dtype=torch.bool)
            rest[g_idx] = False
            assert float(full[rest].abs().max()) == 0.0, k
            full = full[g_idx]
        assert rel(s_grads[k], full) < 1e-10, (k, rel(s_grads[k], full))
