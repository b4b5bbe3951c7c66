"""Shared helpers for the parity tests: run the oracle / the CUDA path on a Scene, error norms."""
import torch

import oracle

LEAF_NAMES = ["means", "quats", "scales", "opacities", "colors", "knots", "exposure_times", "frame_times", "crf_params"]


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a - b||_2 / max(||b||_2, 1e-30) over the whole tensor, b = oracle."""
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def oracle_run(sc, with_grad=True, v_alpha=None, projection_override=None, binning_override=None, straight_through=False, sh=None,
               **kw):
    """float64 oracle on the scene's fp32 inputs (upcast). Returns (ldr, alpha, meta, grads dict).
    ``sh`` [N,K,3]: view-dependent colours from spherical harmonics (the scene's ``colors`` are then unused)."""
    leaves = {}
    if sh is not None:
        leaves["sh_coeffs"] = sh.detach().cpu().double().requires_grad_(with_grad)
        kw["sh_coeffs"] = leaves["sh_coeffs"]
        kw["sh_degree"] = int(round(sh.shape[1] ** 0.5)) - 1
    for k in LEAF_NAMES:
        v = getattr(sc, k)
        if v is None:
            continue
        leaves[k] = v.detach().cpu().double().requires_grad_(with_grad)
    sp = dict(knots=leaves["knots"], knot_t0=sc.knot_t0, knot_dt=sc.knot_dt, frame_times=leaves["frame_times"], kind=sc.spline_kind)
    ldr, alpha, meta = oracle.rasterize(leaves["means"], leaves["quats"], leaves["scales"], leaves["opacities"], leaves["colors"],
                                        None, sc.Ks.cpu(), sc.width, sc.height, leaves["exposure_times"], sc.n_virtual, sc.crf_kind,
                                        leaves.get("crf_params"), spline=sp, projection_override=projection_override,
                                        binning_override=binning_override, straight_through=straight_through, **kw)
    grads = {}
    if with_grad:
        loss = (ldr * sc.v_ldr.cpu().double()).sum()
        if v_alpha is not None:
            loss = loss + (alpha * v_alpha.cpu().double()).sum()
        names = [k for k in leaves]
        gs = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
        grads = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, gs)}
    return ldr.detach(), alpha.detach(), meta, grads


def cuda_run(sc, with_grad=True, v_alpha=None, sort_mode="presort", debug_keys=False, sh=None, **kw):
    """The product path on cuda:0. Returns (ldr, alpha, meta, grads dict)."""
    from casualhdrsplat_b200 import rasterize

    dev = torch.device("cuda:0")
    leaves = {}
    if sh is not None:
        leaves["sh_coeffs"] = sh.detach().to(dev).requires_grad_(with_grad)
        kw["sh_coeffs"] = leaves["sh_coeffs"]
    for k in LEAF_NAMES:
        v = getattr(sc, k)
        if v is None:
            continue
        leaves[k] = v.detach().to(dev).requires_grad_(with_grad)
    sp = dict(knots=leaves["knots"], knot_t0=sc.knot_t0, knot_dt=sc.knot_dt, frame_times=leaves["frame_times"], kind=sc.spline_kind)
    ldr, alpha, meta = rasterize(leaves["means"], leaves["quats"], leaves["scales"], leaves["opacities"], leaves["colors"], None,
                                 sc.Ks.to(dev), sc.width, sc.height, leaves["exposure_times"], sc.n_virtual, sc.crf_kind,
                                 leaves.get("crf_params"), spline=sp, sort_mode=sort_mode, debug_keys=debug_keys, **kw)
    grads = {}
    if with_grad:
        loss = (ldr * sc.v_ldr.to(dev)).sum()
        if v_alpha is not None:
            loss = loss + (alpha * v_alpha.to(dev)).sum()
        names = [k for k in leaves]
        gs = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
        grads = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, gs)}
    torch.cuda.synchronize()
    return ldr.detach(), alpha.detach(), meta, grads


def cuda_projection(meta):
    """Unpack the CUDA path's fp32 projection outputs as the oracle's projection_override."""
    st = meta["state"]
    geom = st.geom.cpu()
    return {"means2d": geom[..., :2].contiguous(), "conics": torch.cat([geom[..., 2:4], st.conic_c.cpu()[..., None]], -1),
            "depths": st.depths.cpu(), "radii": st.radii.cpu()}


def robust_grad_report(a: torch.Tensor, b: torch.Tensor):
    """Per-row relative errors of a gradient tensor against the oracle: (median, fraction of rows off by > 1e-2)."""
    a = a.detach().double().cpu().reshape(a.shape[0], -1)
    b = b.detach().double().cpu().reshape(b.shape[0], -1)
    nz = b.norm(dim=1) > 0
    e = (a - b).norm(dim=1)[nz] / b.norm(dim=1)[nz]
    return float(e.median()), float((e > 1e-2).double().mean())


def huge_splat_scene(n_back=300, width=64, height=48, seed=11):
    """Explicit-camera scene with one near-camera Gaussian whose tight half extents exceed 32767 px (packed radii entry with
    bit 31 set) in front of `n_back` ordinary ones.  Returns the keyword arguments of rasterize() as CPU fp32 tensors."""
    g = torch.Generator().manual_seed(seed)
    f = 50.0
    z = 2.0 + 4.0 * torch.rand(n_back, generator=g)
    xy = (torch.rand(n_back, 2, generator=g) - 0.5) * torch.tensor([width, height]) / f * z[:, None]
    means = torch.cat([torch.tensor([[0.0, 0.0, 0.012]]), torch.cat([xy, z[:, None]], 1)])
    scales = torch.cat([torch.full((1, 3), 3.0), 0.05 + 0.2 * torch.rand(n_back, 3, generator=g)])
    quats = torch.randn(n_back + 1, 4, generator=g)
    quats[0] = torch.tensor([1.0, 0.0, 0.0, 0.0])
    opac = torch.cat([torch.tensor([0.5]), 0.1 + 0.8 * torch.rand(n_back, generator=g)])
    colors = torch.rand(n_back + 1, 3, generator=g) * 2
    K = torch.tensor([[[f, 0, width / 2], [0, f, height / 2], [0, 0, 1]]])
    return dict(means=means, quats=quats, scales=scales, opacities=opac, colors=colors, viewmats=torch.eye(4)[None].clone(), Ks=K,
                width=width, height=height, exposure_times=torch.ones(1), n_virtual=1)


# ---- tile-subset parity at sizes the full oracle cannot finish --------------------------
def tile_pixel_mask(width, height, tile_ids, tile=16):
    """[H,W] bool mask of the pixels of the given tile ids (row-major tile numbering)."""
    tile_w = (width + tile - 1) // tile
    m = torch.zeros(height, width, dtype=torch.bool)
    for tid in tile_ids:
        ty, tx = divmod(int(tid), tile_w)
        m[ty * tile:(ty + 1) * tile, tx * tile:(tx + 1) * tile] = True
    return m


def gaussians_of_tiles(vals_sorted, tile_offsets, n_gauss, n_cams, tiles, tile_ids):
    """Sorted unique Gaussian ids appearing in the lists of (every camera, each chosen tile)."""
    to = (tile_offsets.cpu().to(torch.int64) & 0xFFFFFFFF).tolist()
    vals = vals_sorted.cpu().to(torch.int64)
    parts = [vals[to[c * tiles + t]:to[c * tiles + t + 1]] - c * n_gauss for c in range(n_cams) for t in tile_ids]
    return torch.unique(torch.cat(parts)) if parts else torch.zeros(0, dtype=torch.int64)


def subset_scene(sc, g_idx):
    """The scene restricted to Gaussians g_idx (ascending), everything else unchanged."""
    import dataclasses

    return dataclasses.replace(sc, means=sc.means[g_idx], quats=sc.quats[g_idx], scales=sc.scales[g_idx], opacities=sc.opacities[g_idx],
                               colors=sc.colors[g_idx])


def subset_projection(proj, g_idx):
    return {k: v[:, g_idx].contiguous() for k, v in proj.items()}
