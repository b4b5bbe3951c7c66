"""The kernels' per-element arithmetic (csrc/chs_math.cuh, chs_spline.cuh) compiled for the host by
tests/hostsim and compared with the float64 oracle.  CPU-only; validates every hand-derived
forward/backward formula before a GPU is involved."""
import ctypes
import math

import numpy as np
import pytest
import torch

import oracle
from casualhdrsplat_b200.scene import make_config, make_scene
from oracle import se3
from tests.hostsim.loader import load

@pytest.fixture(autouse=True)
def _float64_default():
    """These tests build float64 tensors implicitly; keep that local to the module's tests."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)
HS = load()


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _scene_cams(name="small", **kw):
    sc = make_config(name, **kw)
    vm = se3.spline_viewmats(sc.knots.double(), sc.knot_t0, sc.knot_dt, sc.frame_times.double(), sc.exposure_times.double(),
                             sc.n_virtual, sc.spline_kind)
    Ks = sc.Ks.double().repeat_interleave(sc.n_virtual, dim=0)
    return sc, vm, Ks


def _hs_project(sc, vm, Ks, dtype):
    N, C = sc.means.shape[0], vm.shape[0]
    np_t = np.float64 if dtype == "f64" else np.float32
    arr = lambda t: np.ascontiguousarray(t.detach().numpy().astype(np_t))
    m2d = np.zeros((C, N, 2), np_t); dep = np.zeros((C, N), np_t); con = np.zeros((C, N, 3), np_t)
    rad = np.zeros((C, N), np.int32); tou = np.zeros((C, N), np.int32)
    fn = getattr(HS, f"hs_project_fwd_{dtype}")
    ct = ctypes.c_double if dtype == "f64" else ctypes.c_float
    a = [arr(sc.means), arr(sc.quats), arr(sc.scales), arr(vm), arr(Ks)]
    fn(N, C, sc.width, sc.height, ct(0.01), ct(1e10), ct(0.3), *[_p(x) for x in a], _p(m2d), _p(dep), _p(con), _p(rad), _p(tou))
    return m2d, dep, con, rad, tou


def test_project_fwd_f64_matches_oracle():
    sc, vm, Ks = _scene_cams()
    m2d, dep, con, rad, tou = _hs_project(sc, vm, Ks, "f64")
    ref = oracle.project(sc.means, sc.quats, sc.scales, vm, Ks, sc.width, sc.height)
    vis = ref["radii"].numpy() > 0
    assert vis.sum() > 1000
    assert np.array_equal(rad, ref["radii"].numpy())
    assert np.allclose(m2d[vis], ref["means2d"].numpy()[vis], rtol=1e-11, atol=1e-9)
    assert np.allclose(con[vis], ref["conics"].numpy()[vis], rtol=1e-9, atol=1e-12)
    assert np.allclose(dep, ref["depths"].numpy(), rtol=1e-12)
    b = oracle.bin_tiles(torch.from_numpy(m2d).float(), torch.from_numpy(rad), torch.from_numpy(dep).float(), sc.width, sc.height)
    assert np.array_equal(tou, b["tiles_touched"].numpy())


def test_project_fwd_f32_close_and_binning_bit_exact():
    sc, vm, Ks = _scene_cams()
    m2d, dep, con, rad, tou = _hs_project(sc, vm.float(), Ks.float(), "f32")
    ref = oracle.project(sc.means, sc.quats, sc.scales, vm.float(), Ks.float(), sc.width, sc.height)
    rr = ref["radii"].numpy()
    both = (rad > 0) & (rr > 0)
    assert (rad != rr).mean() < 1e-3  # ceil() flips between fp32 and fp64 are a counted, tiny set
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert rel(m2d[both], ref["means2d"].numpy()[both]) < 1e-6
    assert rel(con[both], ref["conics"].numpy()[both]) < 1e-4
    # A.4 contract: tiles_touched is an integer function of the fp32 outputs
    b = oracle.bin_tiles(torch.from_numpy(m2d), torch.from_numpy(rad), torch.from_numpy(dep), sc.width, sc.height)
    assert np.array_equal(tou, b["tiles_touched"].numpy())


def test_tile_bounds_bit_exact_random():
    g = torch.Generator().manual_seed(0)
    n = 20000
    mx = ((torch.rand(n, generator=g) * 2400 - 200)).float()
    my = ((torch.rand(n, generator=g) * 1500 - 200)).float()
    mx[:100] = torch.arange(100).float() * 16.0  # exact tile boundaries
    r = torch.randint(0, 300, (n,), generator=g, dtype=torch.int32)
    rect = np.zeros((n, 4), np.int32)
    HS.hs_tile_bounds(n, _p(mx.numpy()), _p(my.numpy()), _p(r.numpy()), 120, 68, _p(rect))
    x0, y0, x1, y1, _ = oracle.tile_bounds(torch.stack([mx, my], -1), torch.clamp(r, min=1), 1920, 1080)
    x0r, y0r, x1r, y1r, _ = oracle.tile_bounds(torch.stack([mx, my], -1), r, 1920, 1080)
    assert np.array_equal(rect[:, 0], x0r.numpy()) and np.array_equal(rect[:, 2], x1r.numpy())
    assert np.array_equal(rect[:, 1], y0r.numpy()) and np.array_equal(rect[:, 3], y1r.numpy())


def test_tight_bounds_radii_and_rectangles():
    """chs_config.tight_bounds: packed per-axis radii (formula) and their tile rectangles (bit-exact vs the oracle)."""
    import math
    g = torch.Generator().manual_seed(6)
    n = 4000
    sxx = torch.exp(torch.randn(n, generator=g, dtype=torch.float64) * 1.5 + 2)
    syy = torch.exp(torch.randn(n, generator=g, dtype=torch.float64) * 1.5 + 2)
    op = torch.rand(n, generator=g, dtype=torch.float64) * 0.99 + 0.0005
    op[:50] = 1.0 / 255.0 * torch.linspace(0.5, 1.5, 50, dtype=torch.float64)  # around the visibility limit
    lam = torch.maximum(sxx, syy) * (1 + torch.rand(n, generator=g, dtype=torch.float64))
    radius = torch.ceil(3 * torch.sqrt(lam)).to(torch.int32)
    p64 = np.zeros(n, np.int32); p32 = np.zeros(n, np.int32)
    HS.hs_tight_radii(n, _p(sxx.numpy()), _p(syy.numpy()), _p(op.numpy()), _p(radius.numpy()), _p(p64), _p(p32))
    margin = oracle.TIGHT_MARGIN
    n_cull = 0
    for i in range(n):
        tau = 2 * (math.log(255 * float(op[i])) + margin)
        if tau <= 0:
            assert p64[i] == 0
            n_cull += 1
            continue
        cap = min(int(radius[i]), 65535)
        rx = max(1, min(cap, math.ceil(math.sqrt(tau * float(sxx[i])))))
        ry = max(1, min(cap, math.ceil(math.sqrt(tau * float(syy[i])))))
        assert p64[i] == (rx | (ry << 16))
        # the fp32 instantiation may differ by one pixel at a ceil() boundary, never more, and never exceeds the square radius
        rx32, ry32 = int(p32[i]) & 0xFFFF, (int(p32[i]) >> 16) & 0xFFFF
        assert abs(rx32 - rx) <= 1 and abs(ry32 - ry) <= 1 and rx32 <= cap and ry32 <= cap
    assert n_cull > 10
    # rectangles of the packed radii: bit-exact against the oracle's tile_bounds(tight=True)
    live = p32 != 0
    mx = (torch.rand(n, generator=g) * 2200 - 140).float()
    my = (torch.rand(n, generator=g) * 1300 - 110).float()
    rect = np.zeros((n, 4), np.int32)
    HS.hs_tile_bounds_packed(n, _p(mx.numpy()), _p(my.numpy()), _p(p32), 120, 68, _p(rect))
    x0, y0, x1, y1, touched = oracle.tile_bounds(torch.stack([mx, my], -1), torch.from_numpy(p32.copy()), 1920, 1080, tight=True)
    for k, t in enumerate([x0, y0, x1, y1]):
        assert np.array_equal(rect[live, k], t.numpy()[live])
    assert int(touched[~torch.from_numpy(live)].sum()) == 0
    # a splat wider than 65534 pixels whose mean is far off screen: the packed radius saturates at 65535 = unbounded, and the
    # rectangle must still cover the screen (the classic square of the true radius would)
    big = np.zeros(2, np.int32)
    b64 = np.zeros(2, np.int32)
    HS.hs_tight_radii(2, _p(np.array([4e9, 100.0])), _p(np.array([50.0, 6e9])), _p(np.array([0.9, 0.9])),
                      _p(np.array([200000, 250000], np.int32)), _p(b64), _p(big))
    assert (big[0] & 0xFFFF) == 0xFFFF and (big[1] >> 16) & 0xFFFF == 0xFFFF and np.array_equal(big, b64)
    rect2 = np.zeros((2, 4), np.int32)
    far = np.array([-100000.0, 500.0], np.float32), np.array([300.0, 150000.0], np.float32)
    HS.hs_tile_bounds_packed(2, _p(far[0]), _p(far[1]), _p(big), 120, 68, _p(rect2))
    assert rect2[0, 0] == 0 and rect2[0, 2] == 120 and rect2[1, 1] == 0 and rect2[1, 3] == 68
    o = oracle.tile_bounds(torch.stack([torch.from_numpy(far[0]), torch.from_numpy(far[1])], -1), torch.from_numpy(big.copy()), 1920, 1080,
                           tight=True)
    for k in range(4):
        assert np.array_equal(rect2[:, k], o[k].numpy())
    # a vertical half extent in [32768, 65534] sets bit 31 of the packed entry: it is an unsigned pair, still live (!= 0), and
    # decodes to the same rectangle on both sides (gating on `> 0` silently culled such splats)
    neg = np.zeros(2, np.int32)
    n64 = np.zeros(2, np.int32)
    HS.hs_tight_radii(2, _p(np.array([1.5e8, 2.0e8])), _p(np.array([2.0e8, 50.0])), _p(np.array([0.5, 0.5])),
                      _p(np.array([46639, 60000], np.int32)), _p(n64), _p(neg))
    assert neg[0] < 0 and (int(neg[0]) >> 16) & 0xFFFF >= 32768 and neg[1] > 0
    rect3 = np.zeros((2, 4), np.int32)
    ctr = np.array([900.0, 40.0], np.float32), np.array([500.0, 700.0], np.float32)
    HS.hs_tile_bounds_packed(2, _p(ctr[0]), _p(ctr[1]), _p(neg), 120, 68, _p(rect3))
    o3 = oracle.tile_bounds(torch.stack([torch.from_numpy(ctr[0]), torch.from_numpy(ctr[1])], -1), torch.from_numpy(neg.copy()), 1920, 1080,
                            tight=True)
    for k in range(4):
        assert np.array_equal(rect3[:, k], o3[k].numpy())
    assert int(o3[4][0]) == 120 * 68 and int(o3[4][1]) > 0


def test_project_bwd_f64_matches_autograd():
    sc, vm, Ks = _scene_cams("tiny")
    N, C = sc.means.shape[0], vm.shape[0]
    leaves = [sc.means.double().requires_grad_(True), sc.quats.double().mul(1.3).requires_grad_(True),
              sc.scales.double().requires_grad_(True), vm.clone().requires_grad_(True)]
    ref = oracle.project(leaves[0], leaves[1], leaves[2], leaves[3], Ks, sc.width, sc.height)
    g = torch.Generator().manual_seed(1)
    vis = (ref["radii"] > 0)
    vm2 = torch.randn(C, N, 2, generator=g) * vis[..., None]
    vc = torch.randn(C, N, 3, generator=g) * vis[..., None]
    loss = (ref["means2d"] * vm2).sum() + (ref["conics"] * vc).sum()
    gm, gq, gs, gv = torch.autograd.grad(loss, leaves)
    arr = lambda t: np.ascontiguousarray(t.detach().numpy().astype(np.float64))
    o_m = np.zeros((N, 3)); o_q = np.zeros((N, 4)); o_s = np.zeros((N, 3)); o_v = np.zeros((C, 12))
    a = [arr(leaves[0]), arr(leaves[1]), arr(leaves[2]), arr(vm), arr(Ks)]
    rad = np.ascontiguousarray(ref["radii"].numpy())
    HS.hs_project_bwd_f64(N, C, sc.width, sc.height, ctypes.c_double(0.3), *[_p(x) for x in a], _p(rad), _p(arr(vm2)), _p(arr(vc)),
                          _p(o_m), _p(o_q), _p(o_s), _p(o_v))
    assert np.allclose(o_m, gm.numpy(), rtol=1e-8, atol=1e-9 * np.abs(gm.numpy()).max())
    assert np.allclose(o_q, gq.numpy(), rtol=1e-8, atol=1e-9 * np.abs(gq.numpy()).max())
    assert np.allclose(o_s, gs.numpy(), rtol=1e-8, atol=1e-9 * np.abs(gs.numpy()).max())
    gvR = gv[:, :3, :3].reshape(C, 9).numpy()
    gvt = gv[:, :3, 3].numpy()
    assert np.allclose(o_v[:, :9], gvR, rtol=1e-8, atol=1e-9 * np.abs(gvR).max())
    assert np.allclose(o_v[:, 9:], gvt, rtol=1e-8, atol=1e-9 * np.abs(gvt).max())


def test_project_bwd_clamped_jacobian_branch():
    # Gaussians far off-axis exercise the frustum clamp of the EWA Jacobian (x~ = z * lim)
    W = H = 64
    K = torch.tensor([[30.0, 0, 32], [0, 30.0, 32], [0, 0, 1]])[None]
    vm = torch.eye(4)[None].clone()
    means = torch.tensor([[9.0, 0.2, 2.0], [-0.3, -8.0, 2.5], [0.1, 0.1, 3.0]])
    quats = torch.tensor([[1.0, 0.1, 0.2, 0.3], [0.5, 0.5, -0.5, 0.1], [1.0, 0, 0, 0]])
    scales = torch.tensor([[3.0, 2.5, 2.0], [2.0, 3.0, 2.2], [0.3, 0.2, 0.1]])
    leaves = [means.clone().requires_grad_(True), quats.clone().requires_grad_(True), scales.clone().requires_grad_(True),
              vm.clone().requires_grad_(True)]
    ref = oracle.project(*leaves, K, W, H)
    assert (ref["radii"] > 0).all()
    g = torch.Generator().manual_seed(2)
    vm2, vc = torch.randn(1, 3, 2, generator=g), torch.randn(1, 3, 3, generator=g)
    gm, gq, gs, gv = torch.autograd.grad((ref["means2d"] * vm2).sum() + (ref["conics"] * vc).sum(), leaves)
    arr = lambda t: np.ascontiguousarray(t.detach().numpy().astype(np.float64))
    o_m = np.zeros((3, 3)); o_q = np.zeros((3, 4)); o_s = np.zeros((3, 3)); o_v = np.zeros((1, 12))
    HS.hs_project_bwd_f64(3, 1, W, H, ctypes.c_double(0.3), _p(arr(means)), _p(arr(quats)), _p(arr(scales)), _p(arr(vm)), _p(arr(K)),
                          _p(np.ascontiguousarray(ref["radii"].numpy())), _p(arr(vm2)), _p(arr(vc)), _p(o_m), _p(o_q), _p(o_s), _p(o_v))
    assert np.allclose(o_m, gm.numpy(), rtol=1e-9) and np.allclose(o_q, gq.numpy(), rtol=1e-8, atol=1e-12)
    assert np.allclose(o_s, gs.numpy(), rtol=1e-9)
    assert np.allclose(o_v[:, :9], gv[:, :3, :3].reshape(1, 9).numpy(), rtol=1e-9)
    assert np.allclose(o_v[:, 9:], gv[:, :3, 3].numpy(), rtol=1e-9)


def _tile_case(seed=0, n_list=60, bg=(0.1, 0.3, 0.2)):
    g = torch.Generator().manual_seed(seed)
    m = torch.rand(n_list, 2, generator=g) * 24 - 4
    L = torch.randn(n_list, 2, 2, generator=g) * 0.35 + torch.eye(2) * 0.5
    con = L @ L.transpose(1, 2) * 0.2 + 0.01 * torch.eye(2)
    conic = torch.stack([con[:, 0, 0], con[:, 0, 1], con[:, 1, 1]], 1)
    o = torch.rand(n_list, generator=g) * 0.95 + 0.02
    o[::7] = 0.9995  # exercise the 0.999 clamp
    col = torch.exp(torch.randn(n_list, 3, generator=g))
    yy, xx = torch.meshgrid(torch.arange(16.0), torch.arange(16.0), indexing="ij")
    pix = torch.stack([xx.reshape(-1) + 0.5, yy.reshape(-1) + 0.5], 1)
    return m, conic, o, col, pix, torch.tensor(bg)


@pytest.mark.parametrize("form", ["direct", "tabled", "tabled_r"])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_blend_pair_math_matches_oracle(dtype, form):
    """form "direct": chs_pair_bwd (nine partials per pair); form "tabled": the per-pair scalars, moment sums and final
    scaling blend_bwd2_kernel uses (chs_pair_bwd_scalars / chs_pair_moments / chs_moments_to_grads); form "tabled_r": the
    division-free normalised-colour recurrence of blend_bwd3_kernel (chs_pair_bwd_scalars_r / chs_moments_to_grads_neg)."""
    m, conic, o, col, pix, bg = _tile_case()
    n_list, n_pix = m.shape[0], pix.shape[0]
    leaves = [t.clone().requires_grad_(True) for t in (m, conic, o, col)]
    vals = torch.arange(n_list, dtype=torch.int32)
    to = torch.tensor([0, n_list])
    hdr, alpha, last = oracle.blend(leaves[0][None], leaves[1][None], leaves[2], leaves[3], vals, to, n_list, 16, 16, background=bg)
    g = torch.Generator().manual_seed(3)
    vh, va = torch.randn(16, 16, 3, generator=g), torch.randn(16, 16, generator=g)
    grads = torch.autograd.grad((hdr[0] * vh).sum() + (alpha[0] * va).sum(), leaves)
    np_t = np.float64 if dtype == "f64" else np.float32
    arr = lambda t: np.ascontiguousarray(t.detach().numpy().astype(np_t))
    params = arr(torch.cat([m, conic, o[:, None], col], 1))
    o_h = np.zeros((n_pix, 3), np_t); o_a = np.zeros(n_pix, np_t); o_l = np.zeros(n_pix, np.int32); o_v = np.zeros((n_list, 9), np_t)
    getattr(HS, f"hs_blend_{dtype}" if form == "direct" else f"hs_blend_{form}_{dtype}")(n_list, _p(params), n_pix, _p(arr(pix)), _p(arr(bg)), _p(arr(vh.reshape(-1, 3))),
                                     _p(arr(va.reshape(-1))), _p(o_h), _p(o_a), _p(o_l), _p(o_v))
    assert (o_l > -1000000).all(), "sub-tile cull dropped a contributing pair"
    tol = 1e-11 if dtype == "f64" else 2e-5
    rel = lambda a, b: np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
    assert rel(o_h, hdr[0].reshape(-1, 3).detach().numpy()) < tol
    assert rel(o_a, alpha[0].reshape(-1).detach().numpy()) < tol
    want_last = (last[0].reshape(-1) + 1).numpy()
    assert (o_l != want_last).mean() <= (0 if dtype == "f64" else 0.01)
    ref = torch.cat([grads[0], grads[1], grads[2][:, None], grads[3]], 1).numpy()
    gtol = 1e-9 if dtype == "f64" else 1e-3
    for k in range(9):
        assert rel(o_v[:, k], ref[:, k]) < gtol, (k, rel(o_v[:, k], ref[:, k]))


def test_design_study_low_precision_table():
    """DESIGN.md "next levers" 1: would a 16-bit table of (vs, f) keep the gradients inside the 1e-3 bar?  fp16 does, with
    little margin (and needs range management for HDR magnitudes); bfloat16 does not.  Pins the numbers the decision rests on."""
    m, conic, o, col, pix, bg = _tile_case()
    n_list, n_pix = m.shape[0], pix.shape[0]
    leaves = [t.clone().requires_grad_(True) for t in (m, conic, o, col)]
    hdr, alpha, _ = oracle.blend(leaves[0][None], leaves[1][None], leaves[2], leaves[3], torch.arange(n_list, dtype=torch.int32),
                                 torch.tensor([0, n_list]), n_list, 16, 16, background=bg)
    g = torch.Generator().manual_seed(3)
    vh, va = torch.randn(16, 16, 3, generator=g), torch.randn(16, 16, generator=g)
    grads = torch.autograd.grad((hdr[0] * vh).sum() + (alpha[0] * va).sum(), leaves)
    ref = torch.cat([grads[0], grads[1], grads[2][:, None], grads[3]], 1).numpy()
    arr = lambda t: np.ascontiguousarray(t.detach().numpy().astype(np.float32))
    params = arr(torch.cat([m, conic, o[:, None], col], 1))
    worst = {}
    for mode, name in [(1, "fp32"), (2, "fp16"), (3, "bf16")]:
        o_h = np.zeros((n_pix, 3), np.float32); o_a = np.zeros(n_pix, np.float32); o_l = np.zeros(n_pix, np.int32)
        o_v = np.zeros((n_list, 9), np.float32)
        HS.hs_blend_tabled_lowp_f32(mode, n_list, _p(params), n_pix, _p(arr(pix)), _p(arr(bg)), _p(arr(vh.reshape(-1, 3))),
                                    _p(arr(va.reshape(-1))), _p(o_h), _p(o_a), _p(o_l), _p(o_v))
        worst[name] = max(np.linalg.norm(o_v[:, k] - ref[:, k]) / np.linalg.norm(ref[:, k]) for k in range(9))
    assert worst["fp32"] < 5e-6 and 5e-5 < worst["fp16"] < 1e-3 and worst["bf16"] > 1e-3, worst


@pytest.mark.parametrize("scale", [1.0, 1e-7, 3e4])
def test_tensor_core_phase_b_precision(scale):
    """blend_bwd4_kernel's phase B as arithmetic: fp16 hi + lo splits of the scaled dL/dsigma and alpha T tables and of the
    upstream gradient, sums against block-origin pixel monomials, shift to the mean (chs_shift_moments).  Must match the fp32
    table whatever the magnitude of the upstream gradient (power-of-two scaling per 8x8 block)."""
    m, conic, o, col, pix, bg = _tile_case()
    m = m + torch.tensor([40.0, -25.0]) * (torch.arange(m.shape[0]) % 3 == 0)[:, None]  # some means far outside the tile
    n_list, n_pix = m.shape[0], pix.shape[0]
    leaves = [t.clone().requires_grad_(True) for t in (m, conic, o, col)]
    hdr, alpha, _ = oracle.blend(leaves[0][None], leaves[1][None], leaves[2], leaves[3], torch.arange(n_list, dtype=torch.int32),
                                 torch.tensor([0, n_list]), n_list, 16, 16, background=bg)
    g = torch.Generator().manual_seed(3)
    vh, va = torch.randn(16, 16, 3, generator=g) * scale, torch.randn(16, 16, generator=g) * scale
    vh[:8, :8] *= 1e-3  # blocks of very different magnitude
    grads = torch.autograd.grad((hdr[0] * vh).sum() + (alpha[0] * va).sum(), leaves)
    ref = torch.cat([grads[0], grads[1], grads[2][:, None], grads[3]], 1).numpy()
    arr = lambda t: np.ascontiguousarray(t.detach().numpy().astype(np.float32))
    params = arr(torch.cat([m, conic, o[:, None], col], 1))
    errs = {}
    for name in ["tabled_r", "tabled_mma"]:
        o_h = np.zeros((n_pix, 3), np.float32); o_a = np.zeros(n_pix, np.float32); o_l = np.zeros(n_pix, np.int32)
        o_v = np.zeros((n_list, 9), np.float32)
        getattr(HS, f"hs_blend_{name}_f32")(n_list, _p(params), n_pix, _p(arr(pix)), _p(arr(bg)), _p(arr(vh.reshape(-1, 3))),
                                            _p(arr(va.reshape(-1))), _p(o_h), _p(o_a), _p(o_l), _p(o_v))
        errs[name] = [np.linalg.norm(o_v[:, k] - ref[:, k]) / np.linalg.norm(ref[:, k]) for k in range(9)]
    # as accurate as the fp32 table of blend_bwd3_kernel
    assert max(errs["tabled_mma"]) < 3 * max(max(errs["tabled_r"]), 1e-6), errs
    # what-if kept for the record: alpha T as a single fp16 would put the colour gradients at 4e-4, too close to the 1e-3 bar
    o_v = np.zeros((n_list, 9), np.float32)
    HS.hs_blend_tabled_mma_variant_f32(5, n_list, _p(params), n_pix, _p(arr(pix)), _p(arr(bg)), _p(arr(vh.reshape(-1, 3))),
                                       _p(arr(va.reshape(-1))), _p(o_h), _p(o_a), _p(o_l), _p(o_v))
    single = max(np.linalg.norm(o_v[:, k] - ref[:, k]) / np.linalg.norm(ref[:, k]) for k in range(6, 9))
    assert 5e-5 < single < 1e-3, single


def test_block_cull_bound_is_conservative_and_tight():
    """chs_block_max_power >= max over the block's pixel centres (never culls a live pair) and is tight."""
    g = torch.Generator().manual_seed(9)
    n = 200000
    th = torch.rand(n, generator=g) * math.pi
    l1 = torch.exp(torch.rand(n, generator=g) * 6 - 3)          # eigenvalues of the conic, 0.05 .. 20
    l2 = l1 * torch.exp(-torch.rand(n, generator=g) * 5)        # anisotropy up to e^5
    c, s_ = torch.cos(th), torch.sin(th)
    A = l1 * c * c + l2 * s_ * s_
    B = (l1 - l2) * c * s_
    C = l1 * s_ * s_ + l2 * c * c
    bx = torch.randint(0, 8, (n,), generator=g).double() * 8 + 0.5
    by = torch.randint(0, 8, (n,), generator=g).double() * 8 + 0.5
    mx = bx + torch.rand(n, generator=g) * 40 - 16
    my = by + torch.rand(n, generator=g) * 30 - 13
    o = torch.rand(n, generator=g) * 0.98 + 0.01
    params = np.ascontiguousarray(torch.stack([mx, my, A, B, C, o], 1).numpy().astype(np.float32))
    rect = np.ascontiguousarray(torch.stack([bx, bx + 7, by, by + 7], 1).numpy().astype(np.float32))
    bound = np.zeros(n, np.float32); brute = np.zeros(n, np.float32)
    HS.hs_block_bound_f32(n, _p(params), _p(rect), _p(bound), _p(brute))
    thr = math.log2(1 / 255)
    live = brute >= thr
    assert live.sum() > 1000
    # conservative: whenever some pixel of the block is live the bound (minus the kernel's margin) keeps it
    assert (bound[live] >= thr - 1e-3).all()
    assert (bound >= brute - 1e-3 * np.maximum(1, np.abs(brute))).all()
    # tight: blocks the bound keeps but that have no live pixel are a small fraction
    kept = bound >= thr - 1e-3
    assert (kept & ~live).sum() < 0.35 * kept.sum()


def test_crf_mlp_fwd_bwd():
    from casualhdrsplat_b200.scene import gamma_crf_params
    P = gamma_crf_params(32).double()
    g = torch.Generator().manual_seed(4)
    X = torch.exp(torch.randn(500, generator=g) * 2 - 3)
    for ch in range(3):
        p = P[ch].clone().requires_grad_(True)
        Xl = X.clone().requires_grad_(True)
        y = oracle.crf_apply(Xl[:, None].expand(-1, 3), oracle.CRF_MLP, torch.stack([p, p, p]))[:, 0]
        vy = torch.randn(500, generator=g)
        gx, gp = torch.autograd.grad((y * vy).sum(), [Xl, p])
        o_y = np.zeros(500); o_d = np.zeros(500); o_p = np.zeros(3 * 32 + 1)
        HS.hs_crf_f64(500, _p(X.numpy()), _p(np.ascontiguousarray(p.detach().numpy())), 32, _p(vy.numpy()), _p(o_y), _p(o_d), _p(o_p))
        assert np.allclose(o_y, y.detach().numpy(), rtol=1e-12)
        assert np.allclose(o_d * vy.numpy(), gx.numpy(), rtol=1e-10, atol=1e-14)
        assert np.allclose(o_p, gp.numpy(), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("hd", [8, 32, 64, 128])
def test_crf_mlp_interval_form_matches_the_per_unit_form(hd):
    """crf_bwd_interval_kernel's algorithm (ranked breakpoints, slope / offset per interval, two sums per interval) restated
    serially with the kernel's own helper functions: same y, dy/dX and parameter gradients as the per-unit form and as the
    oracle's autograd, also with zero weights, duplicated breakpoints and units that are on or off everywhere."""
    from casualhdrsplat_b200.scene import gamma_crf_params
    g = torch.Generator().manual_seed(40 + hd)
    P = gamma_crf_params(hd).double()
    p0 = P[1].clone()
    p0[0] = 0.0                      # w1 = 0 with b1 > 0: on everywhere
    p0[hd] = abs(float(p0[hd])) + 0.1
    p0[1] = 0.0                      # w1 = 0 with b1 < 0: never on
    p0[hd + 1] = -abs(float(p0[hd + 1])) - 0.1
    p0[3], p0[hd + 3] = p0[2], p0[hd + 2]          # duplicated breakpoint
    p0[4] = -abs(float(p0[4])) - 0.05               # a negative slope
    n = 4000
    X = torch.exp(torch.randn(n, generator=g, dtype=torch.float64) * 2.5 - 3)
    X[::97] = -0.01                 # below the clamp: zero gradient
    vy = torch.randn(n, generator=g, dtype=torch.float64)
    p = p0.clone().requires_grad_(True)
    Xl = X.clone().requires_grad_(True)
    y = oracle.crf_apply(Xl[:, None].expand(-1, 3), oracle.CRF_MLP, torch.stack([p, p, p]))[:, 0]
    gx, gp = torch.autograd.grad((y * vy).sum(), [Xl, p])
    outs = {}
    for name in ["hs_crf_f64", "hs_crf_interval_f64"]:
        o_y = np.zeros(n); o_d = np.zeros(n); o_p = np.zeros(3 * hd + 1)
        getattr(HS, name)(n, _p(X.numpy()), _p(np.ascontiguousarray(p0.numpy())), hd, _p(vy.numpy()), _p(o_y), _p(o_d), _p(o_p))
        outs[name] = (o_y, o_d, o_p)
        assert np.allclose(o_y, y.detach().numpy(), rtol=1e-11), name
        assert np.allclose(o_d * vy.numpy(), gx.numpy(), rtol=1e-9, atol=1e-13), name
        assert np.allclose(o_p, gp.numpy(), rtol=1e-9, atol=1e-10), name
    # single precision, as the kernel runs it: within 1e-5 of the float64 result
    f32 = lambda t: np.ascontiguousarray(np.asarray(t, dtype=np.float32))
    o_y = np.zeros(n, np.float32); o_d = np.zeros(n, np.float32); o_p = np.zeros(3 * hd + 1, np.float32)
    HS.hs_crf_interval_f32(n, _p(f32(X.numpy())), _p(f32(p0.numpy())), hd, _p(f32(vy.numpy())), _p(o_y), _p(o_d), _p(o_p))
    assert np.abs(o_y - y.detach().numpy()).max() < 2e-5
    ref = gp.numpy()
    assert np.linalg.norm(o_p - ref) / np.linalg.norm(ref) < 1e-4


def test_crf_lut_fwd_bwd():
    from casualhdrsplat_b200.scene import gamma_lut_params
    L = 48
    P = gamma_lut_params(L).double()
    g = torch.Generator().manual_seed(5)
    # exposures on both sides of the table range (z in [-10, 1.5]) as well as inside it
    X = torch.exp(torch.randn(800, generator=g) * 4 - 4)
    for ch in range(3):
        p = P[ch].clone().requires_grad_(True)
        Xl = X.clone().requires_grad_(True)
        y = oracle.crf_apply(Xl[:, None].expand(-1, 3), oracle.CRF_LUT, torch.stack([p, p, p]))[:, 0]
        vy = torch.randn(800, generator=g)
        gx, gp = torch.autograd.grad((y * vy).sum(), [Xl, p])
        assert float(gp[:2].abs().max()) == 0.0  # the range entries are a fixed calibration
        o_y = np.zeros(800); o_d = np.zeros(800); o_p = np.zeros(L + 2)
        HS.hs_crf_lut_f64(800, _p(X.numpy()), _p(np.ascontiguousarray(p.detach().numpy())), L, _p(vy.numpy()), _p(o_y), _p(o_d), _p(o_p))
        assert np.allclose(o_y, y.detach().numpy(), rtol=1e-12, atol=1e-15)
        assert np.allclose(o_d * vy.numpy(), gx.numpy(), rtol=1e-10, atol=1e-14)
        assert np.allclose(o_p, gp.numpy(), rtol=1e-10, atol=1e-12)
        # the fp32 instantiation the kernels use: the curve is continuous, so a knot decided differently costs nothing
        o32 = np.zeros(800, np.float32)
        HS.hs_crf_lut_f32(800, _p(X.numpy().astype(np.float32)), _p(np.ascontiguousarray(P[ch].numpy().astype(np.float32))), L, _p(o32))
        assert np.abs(o32 - y.detach().numpy()).max() < 2e-5
    # outside the table the response is constant
    lo = oracle.crf_apply(torch.full((1, 3), 1e-9, dtype=torch.float64), oracle.CRF_LUT, P)
    hi = oracle.crf_apply(torch.full((1, 3), 1e3, dtype=torch.float64), oracle.CRF_LUT, P)
    assert torch.allclose(lo[0], P[:, 2]) and torch.allclose(hi[0], P[:, -1])


@pytest.mark.parametrize("kind,name", [(se3.SPLINE_LINEAR, "c2"), (se3.SPLINE_CUBIC, "tiny")])
def test_spline_fwd_bwd_matches_oracle(kind, name):
    sc = make_config(name, n_gauss=4) if name == "c2" else make_config(name)
    B, n, K = sc.n_frames, sc.n_virtual, sc.knots.shape[0]
    knots = sc.knots.double().requires_grad_(True)
    ft = sc.frame_times.double().requires_grad_(True)
    ex = sc.exposure_times.double().requires_grad_(True)
    vm = se3.spline_viewmats(knots, sc.knot_t0, sc.knot_dt, ft, ex, n, kind)
    out = np.zeros((B * n, 16))
    f32 = lambda t: np.ascontiguousarray(t.detach().numpy().astype(np.float32))
    args = (kind, _p(f32(sc.knots)), K, ctypes.c_double(sc.knot_t0), ctypes.c_double(sc.knot_dt), _p(f32(sc.frame_times)),
            _p(f32(sc.exposure_times)), B, n)
    HS.hs_spline_fwd(*args, _p(out))
    assert np.allclose(out.reshape(-1, 4, 4), vm.detach().numpy(), rtol=0, atol=1e-12)
    g = torch.Generator().manual_seed(5)
    vv = torch.randn(B * n, 4, 4, generator=g)
    gk, gf, ge = torch.autograd.grad((vm * vv).sum(), [knots, ft, ex])
    o_k = np.zeros((K, 7)); o_f = np.zeros(B); o_e = np.zeros(B)
    HS.hs_spline_bwd(*args, _p(np.ascontiguousarray(vv.numpy())), _p(o_k), _p(o_f), _p(o_e))
    assert np.allclose(o_k, gk.numpy(), rtol=1e-8, atol=1e-9 * np.abs(gk.numpy()).max())
    assert np.allclose(o_f, gf.numpy(), rtol=1e-8, atol=1e-12)
    assert np.allclose(o_e, ge.numpy(), rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_sh_colour_fwd_bwd_matches_oracle(deg):
    from oracle.sh import sh_colors

    g = torch.Generator().manual_seed(40 + deg)
    N, C, K = 300, 3, (deg + 1) ** 2
    sh = (torch.randn(N, K, 3, generator=g) * 0.6).requires_grad_(True)
    means = (torch.randn(N, 3, generator=g) * 2 + torch.tensor([0.0, 0.0, 6.0])).requires_grad_(True)
    q = torch.randn(C, 4, generator=g)
    R = se3.quat_to_rotmat(q)
    vm = torch.eye(4).repeat(C, 1, 1)
    vm[:, :3, :3] = R
    vm[:, :3, 3] = torch.randn(C, 3, generator=g)
    vm = vm.requires_grad_(True)
    col = sh_colors(sh, means, vm, deg)
    assert (col > 0).any() and (deg == 0 or (col == 0).any()), "both sides of the relu should be exercised"
    v = torch.randn(C, N, 3, generator=g)
    gs, gm, gv = torch.autograd.grad((col * v).sum(), [sh, means, vm], allow_unused=True)
    gm = gm if gm is not None else torch.zeros_like(means)
    gv = gv if gv is not None else torch.zeros_like(vm)
    arr = lambda t: np.ascontiguousarray(t.detach().numpy().astype(np.float64))
    rgb = np.zeros((C, N, 3)); o_sh = np.zeros((N, K, 3)); o_m = np.zeros((N, 3)); o_v = np.zeros((C, 12))
    HS.hs_sh_fwd_bwd(N, C, deg, _p(arr(sh)), _p(arr(means)), _p(arr(vm)), _p(arr(v)), _p(rgb), _p(o_sh), _p(o_m), _p(o_v))
    assert np.allclose(rgb, col.detach().numpy(), rtol=1e-12, atol=1e-13)
    assert np.allclose(o_sh, gs.numpy(), rtol=1e-10, atol=1e-13)
    assert np.allclose(o_m, gm.numpy(), rtol=1e-9, atol=1e-12)
    assert np.allclose(o_v[:, :9], gv[:, :3, :3].reshape(C, 9).numpy(), rtol=1e-9, atol=1e-11)
    assert np.allclose(o_v[:, 9:], gv[:, :3, 3].numpy(), rtol=1e-9, atol=1e-11)
