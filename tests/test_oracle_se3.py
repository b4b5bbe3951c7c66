"""T1 known-answer tests for the oracle's SE(3) spline (SURVEY.md section 4, items iv)."""
import math

import pytest
import torch

from oracle import se3

@pytest.fixture(autouse=True)
def _float64_default():
    """These tests build float64 tensors implicitly; keep that local to the module's tests."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)


def _rand_pose(g, scale=0.5):
    t = torch.randn(3, generator=g)
    q = torch.randn(4, generator=g)
    return t, q / q.norm()


def test_exp_log_roundtrip():
    g = torch.Generator().manual_seed(0)
    for mag in [0.0, 1e-7, 1e-5, 1e-3, 0.3, 2.5]:
        phi = torch.randn(3, generator=g)
        phi = phi / phi.norm() * mag
        rho = torch.randn(3, generator=g)
        R, t = se3.se3_exp(rho, phi)
        # build quaternion of R by exponentiating half angle
        half = phi / 2
        th = half.norm()
        q = torch.cat([torch.cos(th)[None], half * (torch.sin(th) / th if th > 0 else 1.0)])
        assert torch.allclose(se3.quat_to_rotmat(q), R, atol=1e-12)
        rho2, phi2 = se3.se3_rel_log(torch.zeros(3), torch.tensor([1.0, 0, 0, 0]), t, q)
        assert torch.allclose(phi2, phi, atol=1e-10)
        assert torch.allclose(rho2, rho, atol=1e-9)


def test_v_vinv_inverse():
    g = torch.Generator().manual_seed(1)
    for mag in [1e-6, 5e-5, 2e-4, 0.7]:
        phi = torch.randn(3, generator=g)
        phi = phi / phi.norm() * mag
        assert torch.allclose(se3.se3_V(phi) @ se3.se3_Vinv(phi), torch.eye(3), atol=1e-10)


def test_linear_spline_hits_knots():
    g = torch.Generator().manual_seed(2)
    poses = [_rand_pose(g) for _ in range(3)]
    knots = torch.stack([torch.cat(p) for p in poses])
    times = torch.tensor([0.0, 1.0, 1.0 - 1e-12, 2.0])
    R, t = se3.spline_c2w(knots, 0.0, 1.0, times, se3.SPLINE_LINEAR)
    for idx, k in [(0, 0), (1, 1), (3, 2)]:
        assert torch.allclose(R[idx], se3.quat_to_rotmat(knots[k, 3:]), atol=1e-9)
        assert torch.allclose(t[idx], knots[k, :3], atol=1e-9)


def test_cubic_equal_knots_is_constant():
    g = torch.Generator().manual_seed(3)
    t0, q0 = _rand_pose(g)
    knots = torch.cat([t0, q0])[None].repeat(5, 1)
    times = torch.tensor([1.0, 1.3, 1.99, 2.5])
    R, t = se3.spline_c2w(knots, 0.0, 1.0, times, se3.SPLINE_CUBIC)
    assert torch.allclose(R, se3.quat_to_rotmat(q0).expand(4, 3, 3), atol=1e-12)
    assert torch.allclose(t, t0.expand(4, 3), atol=1e-12)


def test_cubic_collinear_constant_velocity():
    # knots T_j = T_0 Exp(j xi): the cumulative B-spline must reproduce T_0 Exp((s-1 + b1+b2+b3) xi)
    # and b1+b2+b3 = u + 1 exactly, i.e. constant twist velocity.
    xi_rho = torch.tensor([0.3, -0.1, 0.2])
    xi_phi = torch.tensor([0.02, 0.05, -0.03])
    t0 = torch.tensor([0.1, 0.2, 0.3])
    q0 = torch.tensor([0.9, 0.1, -0.3, 0.2])
    q0 = q0 / q0.norm()
    R0 = se3.quat_to_rotmat(q0)
    knots = []
    for j in range(6):
        Rj, tj = se3.se3_exp(j * xi_rho, j * xi_phi)
        Rw, tw = se3.compose(R0, t0, Rj, tj)
        # rotation -> quaternion through the log/exp of the known twist
        half = j * xi_phi / 2
        th = half.norm()
        dq = torch.cat([torch.cos(th)[None], half * (torch.sin(th) / th if th > 0 else 1.0)])
        knots.append(torch.cat([tw, se3.quat_mul(q0, dq)]))
    knots = torch.stack(knots)
    times = torch.tensor([1.0, 1.25, 2.5, 3.75])
    R, t = se3.spline_c2w(knots, 0.0, 1.0, times, se3.SPLINE_CUBIC)
    for i, tm in enumerate(times.tolist()):
        Re, te = se3.se3_exp(tm * xi_rho, tm * xi_phi)
        Rw, tw = se3.compose(R0, t0, Re, te)
        assert torch.allclose(R[i], Rw, atol=1e-9)
        assert torch.allclose(t[i], tw, atol=1e-9)


def test_sample_times_window():
    ft = torch.tensor([1.0, 2.0])
    ex = torch.tensor([0.2, 0.4])
    ts = se3.sample_times(ft, ex, 5).reshape(2, 5)
    assert torch.allclose(ts[:, 0], ft - ex / 2) and torch.allclose(ts[:, -1], ft + ex / 2)
    assert torch.allclose(ts[:, 2], ft)
    assert torch.equal(se3.sample_times(ft, ex, 1), ft)


@pytest.mark.parametrize("kind", [se3.SPLINE_LINEAR, se3.SPLINE_CUBIC])
def test_spline_gradcheck(kind):
    g = torch.Generator().manual_seed(4)
    K = 6
    knots = torch.cat([torch.randn(K, 3, generator=g) * 0.2, torch.tensor([1.0, 0, 0, 0]) + 0.05 * torch.randn(K, 4, generator=g)], 1)
    knots.requires_grad_(True)
    ft = torch.tensor([1.6, 2.4], requires_grad=True)
    ex = torch.tensor([0.3, 0.5], requires_grad=True)

    def f(kn, a, b):
        return se3.spline_viewmats(kn, 0.0, 1.0, a, b, 3, kind)[:, :3, :]

    assert torch.autograd.gradcheck(f, (knots, ft, ex), eps=1e-6, atol=1e-6)
