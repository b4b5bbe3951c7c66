"""Stage-wise gradient error diagnosis against the oracle (test infrastructure: it runs the oracle, so it lives under tests/). Usage: python tests/tools/diag_grads.py [config]"""
import os
import sys
from ctypes import byref

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from casualhdrsplat_b200 import _lib, api  # noqa: E402
from casualhdrsplat_b200.scene import make_config  # noqa: E402
from tests.util import cuda_projection, cuda_run, oracle_run, rel  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    sc = make_config(name)
    ldr, alpha, meta, grads = cuda_run(sc)
    o_ldr, o_alpha, o_meta, o_grads = oracle_run(sc)
    print("M", meta["n_isect"], o_meta["n_isect"], "fwd rel", rel(ldr, o_ldr))
    for k in grads:
        print("  e2e grad", k, rel(grads[k], o_grads[k]))
    # outlier structure of the end-to-end error
    for k in ["means", "scales", "colors"]:
        a, b_ = grads[k].cpu().double(), o_grads[k]
        e = (a - b_).norm(dim=1)
        tot = float(e.norm())
        top = torch.topk(e, 10)
        print(f"  {k}: total err {tot:.3e}; top-10 gaussians carry {float(top.values.norm()) / tot:.3f} of it; ids {top.indices.tolist()[:5]}")
    vs_c, vs_o = meta["state"].vals_sorted.cpu()[: meta["n_isect"]], o_meta["bins"]["vals_sorted"]
    print("  list entries that differ between CUDA and oracle end-to-end:", int((vs_c != vs_o).sum()) if vs_c.numel() == vs_o.numel() else "length differs")
    rc, ro = meta["state"].radii.cpu(), o_meta["proj"]["radii"]
    print("  radius flips:", int((rc != ro).sum()))
    st = meta["state"]
    proj = cuda_projection(meta)
    B, n, H, W = sc.n_frames, sc.n_virtual, sc.height, sc.width
    # ---- stage A: blend + epilogue backward with identical lists ----
    m2d = proj["means2d"].double().requires_grad_(True)
    con = proj["conics"].double().requires_grad_(True)
    op = sc.opacities.double().requires_grad_(True)
    col = sc.colors.double().requires_grad_(True)
    b = oracle.bin_tiles(proj["means2d"], proj["radii"], proj["depths"], W, H)
    hdr, alpha_c, _ = oracle.blend(m2d, con, op, col, b["vals_sorted"], b["tile_offsets"], sc.means.shape[0], W, H)
    o2_ldr, _, _ = oracle.formation(hdr, alpha_c, sc.exposure_times, n, sc.crf_kind, sc.crf_params)
    gm, gc, go, gcol = torch.autograd.grad((o2_ldr * sc.v_ldr.double()).sum(), [m2d, con, op, col])
    print("same-list fwd rel", rel(ldr, o2_ldr))
    dev = st.geom.device
    g = api.backward_stages(st, sc.means.to(dev), sc.quats.to(dev), sc.scales.to(dev), sc.exposure_times.to(dev),
                            sc.crf_params.to(dev) if sc.crf_params is not None else None, sc.v_ldr.to(dev).contiguous(), None)
    torch.cuda.synchronize()
    v_geom, v_cogr, v_blue = g["v_geom"].cpu(), g["v_cogr"].cpu(), g["v_blue"].cpu()
    print("stage A  v_means2d", rel(v_geom[..., :2], gm), " v_conics", rel(torch.cat([v_geom[..., 2:], v_cogr[..., :1]], -1), gc),
          " v_opac", rel(v_cogr[..., 1].sum(0), go), " v_col", rel(torch.cat([v_cogr[..., 2:], v_blue[..., None]], -1).sum(0), gcol))
    # per size class
    rad = proj["radii"]
    for lo, hi in [(1, 3), (3, 6), (6, 12), (12, 1000)]:
        m = (rad >= lo) & (rad < hi)
        if m.sum() > 0:
            print(f"   radius [{lo},{hi}) n={int(m.sum())}  v_means2d {rel(v_geom[..., :2][m], gm[m]):.2e}  v_conic_A {rel(v_geom[..., 2][m], gc[..., 0][m]):.2e}")
    # ---- stage B: projection backward fed with the oracle's stage gradients ----
    leaves = [sc.means.double().requires_grad_(True), sc.quats.double().requires_grad_(True), sc.scales.double().requires_grad_(True)]
    vm = meta["viewmats"].cpu().double().requires_grad_(True)
    Ks = sc.Ks.repeat_interleave(n, 0).double()
    pr = oracle.project(leaves[0], leaves[1], leaves[2], vm, Ks, W, H)
    vis = (proj["radii"] > 0)
    og = torch.autograd.grad((pr["means2d"] * (gm * vis[..., None])).sum() + (pr["conics"] * (gc * vis[..., None])).sum(), leaves + [vm])
    C, N = rad.shape
    f32 = torch.float32
    vg = torch.cat([gm, gc[..., :2]], -1).to(f32).to(dev).contiguous()
    vc = torch.zeros(C, N, 4, dtype=f32); vc[..., 0] = gc[..., 2].to(f32)
    vc = vc.to(dev).contiguous()
    vb = torch.zeros(C, N, dtype=f32, device=dev)
    flat = torch.empty(14 * N, dtype=f32, device=dev); vvm = torch.empty(C, 4, 4, dtype=f32, device=dev)
    ws = _lib.workspace_sizes(st.cfg, 0, st.n_knots)
    red = torch.empty(int(ws.reduce_bytes), dtype=torch.uint8, device=dev)
    dm, dq, ds = sc.means.to(dev), sc.quats.to(dev), sc.scales.to(dev)
    _lib.check(_lib.lib().chs_project_bwd(byref(st.cfg), _lib.ptr(dm), _lib.ptr(dq), _lib.ptr(ds),
                                          _lib.ptr(st.viewmats), _lib.ptr(st.Ks), _lib.ptr(st.radii), _lib.ptr(vg), _lib.ptr(vc), _lib.ptr(vb),
                                          _lib.ptr(flat), _lib.ptr(vvm), _lib.ptr(red), red.numel(), api._stream()))
    torch.cuda.synchronize()
    a, q, s, _, _ = api.split_flat_grads(flat.cpu(), N)
    print("stage B  v_means", rel(a, og[0]), " v_quats", rel(q, og[1]), " v_scales", rel(s, og[2]), " v_viewmats", rel(vvm.cpu()[:, :3], og[3][:, :3]))


if __name__ == "__main__":
    main()
