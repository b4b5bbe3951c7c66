"""T1 known-answer tests for the float64 oracle of the formation model."""
import math

import pytest
import torch

import oracle
from casualhdrsplat_b200.scene import make_config, make_scene, gamma_crf_params
from oracle import se3

@pytest.fixture(autouse=True)
def _float64_default():
    """These tests build float64 tensors implicitly; keep that local to the module's tests."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)


def _run(sc, leaves=None, **kw):
    leaves = leaves or {}
    get = lambda k: leaves.get(k, getattr(sc, k))
    sp = dict(knots=get("knots"), knot_t0=sc.knot_t0, knot_dt=sc.knot_dt, frame_times=get("frame_times"), kind=sc.spline_kind)
    crf = leaves.get("crf_params", sc.crf_params)
    return oracle.rasterize(get("means"), get("quats"), get("scales"), get("opacities"), get("colors"), None, sc.Ks,
                            sc.width, sc.height, get("exposure_times"), kw.pop("n_virtual", sc.n_virtual),
                            kw.pop("crf_kind", sc.crf_kind), crf, spline=sp, **kw)


def test_single_isotropic_gaussian_closed_form():
    # One isotropic Gaussian on the optical axis: pixel = o * exp(-r^2 / (2 s2)) * c, s2 = (fx s / z)^2 + 0.3
    W = H = 32
    fx = 20.0
    z, s, o = 4.0, 0.5, 0.6
    col = torch.tensor([2.0, 0.5, 7.0])
    K = torch.tensor([[fx, 0, W / 2], [0, fx, H / 2], [0, 0, 1.0]])[None]
    vm = torch.eye(4)[None]
    ldr, alpha, meta = oracle.rasterize(torch.tensor([[0.0, 0.0, z]]), torch.tensor([[1.0, 0, 0, 0]]),
                                        torch.tensor([[s, s, s]]), torch.tensor([o]), col[None], vm, K, W, H,
                                        torch.tensor([1.0]), 1, oracle.CRF_IDENTITY)
    s2 = (fx * s / z) ** 2 + 0.3
    for (i, j) in [(16, 16), (15, 17), (10, 20), (3, 4)]:
        r2 = (j + 0.5 - W / 2) ** 2 + (i + 0.5 - H / 2) ** 2
        a = o * math.exp(-r2 / (2 * s2))
        a = a if a >= 1 / 255 else 0.0
        assert torch.allclose(ldr[0, i, j], a * col, atol=1e-12)
        assert abs(float(alpha[0, i, j, 0]) - a) < 1e-12
    assert int(meta["proj"]["radii"][0, 0]) == math.ceil(3 * math.sqrt(s2 + 0.1))  # lambda = m + sqrt(max(0.01, 0))


def test_static_camera_blur_equals_sharp():
    sc = make_config("tiny", static_camera=True)
    a, _, _ = _run(sc)
    b, _, _ = _run(sc, n_virtual=1)
    assert torch.allclose(a, b, atol=1e-12)


def test_identity_crf_linear_in_exposure():
    sc = make_config("tiny", crf_kind=oracle.CRF_IDENTITY, static_camera=True)
    a, _, _ = _run(sc)
    b, _, _ = _run(sc, leaves={"exposure_times": sc.exposure_times.double() * 3})
    assert torch.allclose(b, 3 * a, rtol=1e-12, atol=0)


def test_crf_order_flag_coincides_for_identity_and_differs_for_mlp():
    sc = make_config("tiny")
    a, _, _ = _run(sc, crf_kind=oracle.CRF_IDENTITY)
    b, _, _ = _run(sc, crf_kind=oracle.CRF_IDENTITY, crf_before_average=True)
    assert torch.allclose(a, b, atol=1e-13)
    c, _, _ = _run(sc)
    d, _, _ = _run(sc, crf_before_average=True)
    assert (c - d).abs().max() > 1e-6


def test_vectorised_blend_matches_literal_loop():
    sc = make_config("tiny", n_frames=1, n_virtual=2)
    _, _, meta = _run(sc, background=torch.tensor([0.1, 0.2, 0.3]))
    p, b = meta["proj"], meta["bins"]
    g = torch.Generator().manual_seed(0)
    for _ in range(40):
        c = int(torch.randint(0, 2, (1,), generator=g))
        i = int(torch.randint(0, sc.height, (1,), generator=g))
        j = int(torch.randint(0, sc.width, (1,), generator=g))
        pix, al, last = oracle.blend_pixel_loop(p["means2d"], p["conics"], sc.opacities.double(), sc.colors.double(),
                                                b["vals_sorted"], b["tile_offsets"], sc.means.shape[0], sc.width,
                                                sc.height, c, i, j, background=[0.1, 0.2, 0.3])
        assert torch.allclose(meta["hdr_cams"][c, i, j], torch.tensor(pix), atol=1e-13)
        assert abs(float(meta["alpha_cams"][c, i, j]) - al) < 1e-13
        assert int(meta["last_id"][c, i, j]) == last


def test_gamma_crf_is_gamma_like():
    P = gamma_crf_params(64)
    X = torch.tensor([0.001, 0.01, 0.1, 0.5])[:, None].expand(4, 3)
    y = oracle.crf_apply(X, oracle.CRF_MLP, P)
    assert ((y - X ** (1 / 2.2)).abs() < 0.05).all()
    assert (y[1:] > y[:-1]).all()


def _directional_check(sc, names, seed=0, eps=1e-6, rtol=2e-5, **kw):
    g = torch.Generator().manual_seed(seed)
    base = {k: getattr(sc, k).double() for k in names}
    leaves = {k: v.clone().requires_grad_(True) for k, v in base.items()}
    ldr, alpha, _ = _run(sc, leaves=leaves, **kw)
    wl = torch.randn(ldr.shape, generator=g)
    wa = torch.randn(alpha.shape, generator=g)
    loss = (ldr * wl).sum() + (alpha * wa).sum()
    grads = torch.autograd.grad(loss, [leaves[k] for k in names])
    for k, gk in zip(names, grads):
        d = torch.randn(base[k].shape, generator=g)
        d = d / d.norm()
        an = float((gk * d).sum())
        errs = []
        # depth-order swaps ("popping") remain genuine discontinuities; a step that crosses one is
        # an outlier, so the best of three step sizes is taken.
        for h in (eps, eps * 1e-1, eps * 1e-2):
            def L(sign):
                lv = dict(base)
                lv[k] = base[k] + sign * h * d
                l2, a2, _ = _run(sc, leaves=lv, **kw)
                return float((l2 * wl).sum() + (a2 * wa).sum())

            fd = (L(+1) - L(-1)) / (2 * h)
            errs.append(abs(fd - an) / max(abs(an), 1e-9))
        assert min(errs) <= rtol, (k, an, errs)


def test_full_oracle_directional_derivatives():
    # <=64 Gaussians, 32x32, n=3, MLP CRF (SURVEY.md T1 item v); central differences in random directions.
    # The model's skip/stop/radius decisions are discontinuities that carry no gradient by definition
    # (SURVEY.md A.6), so finite differences are taken on the discontinuity-free variant of the same code.
    sc = make_scene(64, 32, 32, n_frames=2, n_virtual=3, crf_hidden=16, scale_mult=8.0)
    _directional_check(sc, ["means", "quats", "scales", "opacities", "colors", "knots", "exposure_times",
                            "frame_times", "crf_params"], alpha_min=0.0, t_stop=0.0, radius_sigmas=9.0)
    # parameters that cannot move a decision are checked on the unmodified model as well
    _directional_check(sc, ["colors", "crf_params"])


def test_exposure_gradient_has_two_paths():
    # brightness path only: static camera, identity CRF -> d ldr / d dt = hdr_mean
    sc = make_config("tiny", static_camera=True, crf_kind=oracle.CRF_IDENTITY)
    ex = sc.exposure_times.double().requires_grad_(True)
    ldr, _, meta = _run(sc, leaves={"exposure_times": ex})
    (gr,) = torch.autograd.grad(ldr.sum(), ex)
    assert torch.allclose(gr, meta["hdr_mean"].sum(dim=(1, 2, 3)), rtol=1e-10)
    # window path only: moving camera, loss on alpha (independent of brightness) still depends on dt
    sc2 = make_config("tiny")
    ex2 = sc2.exposure_times.double().requires_grad_(True)
    _, alpha, _ = _run(sc2, leaves={"exposure_times": ex2})
    g = torch.Generator().manual_seed(5)
    (gr2,) = torch.autograd.grad((alpha * torch.randn(alpha.shape, generator=g)).sum(), ex2)
    assert gr2.abs().min() > 0


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_tight_bounds_change_lists_but_not_images(name):
    """Opacity-aware per-axis bounds only drop tiles in which every pixel fails alpha >= 1/255: images and gradients are
    bit-identical to the classic square bounds, with fewer intersections."""
    from tests.util import oracle_run

    sc = make_config(name)
    a = oracle_run(sc)
    b = oracle_run(sc, tight_bounds=True)
    assert b[2]["n_isect"] < 0.9 * a[2]["n_isect"]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    for k in a[3]:
        assert torch.equal(a[3][k], b[3][k]), k
    # every tight rectangle lies inside the square one
    sq = oracle.tile_bounds(a[2]["proj"]["means2d"].detach().float(), a[2]["proj"]["radii"], sc.width, sc.height)
    tg = oracle.tile_bounds(b[2]["proj"]["means2d"].detach().float(), b[2]["proj"]["radii"], sc.width, sc.height, tight=True)
    live = tg[4] > 0
    assert bool((tg[0][live] >= sq[0][live]).all() and (tg[2][live] <= sq[2][live]).all())
    assert bool((tg[1][live] >= sq[1][live]).all() and (tg[3][live] <= sq[3][live]).all())


def test_tight_bounds_keep_splats_with_half_extent_above_32767():
    """A near-camera floater whose opacity-aware vertical half extent is >= 32768 px packs to a NEGATIVE int32 entry; it must
    stay live (gate on != 0) so that tight and square bounds render the same frame."""
    from tests.util import huge_splat_scene

    kw = huge_splat_scene()
    a = oracle.rasterize(**kw)
    b = oracle.rasterize(**kw, tight_bounds=True)
    packed = b[2]["proj"]["radii"][0, 0]
    assert int(packed) < 0 and (int(packed) >> 16) & 0xFFFF >= 32768
    tiles = ((kw["width"] + 15) // 16) * ((kw["height"] + 15) // 16)
    tb = oracle.tile_bounds(b[2]["proj"]["means2d"].detach().float(), b[2]["proj"]["radii"], kw["width"], kw["height"], tight=True)
    assert int(tb[4][0, 0]) == tiles  # the floater covers every tile
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert float((a[0] - oracle.rasterize(**{**kw, "opacities": kw["opacities"] * torch.cat([torch.zeros(1), torch.ones(300)])})[0]).abs().max()) > 1e-3
