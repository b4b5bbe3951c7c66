"""KATs of the D-SSIM oracle (oracle/ssim.py): closed-form cases and an explicit-loop cross-check of the convolution form."""
import pytest
import torch

from oracle import ssim


@pytest.fixture(autouse=True)
def _f64_default():
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)


def _pair(seed=0, n=2, h=19, w=23):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, h, w, 3, generator=g, dtype=torch.float64)
    y = (x + 0.2 * torch.randn(n, h, w, 3, generator=g, dtype=torch.float64)).clamp(0, 1)
    return x, y


def test_window_is_normalised_and_symmetric():
    g = ssim.gaussian_window()
    assert g.shape == (11,) and abs(float(g.sum()) - 1.0) < 1e-15
    assert torch.equal(g, g.flip(0)) and float(g[5]) == float(g.max())


def test_identical_images_have_ssim_one_and_zero_loss():
    x, _ = _pair()
    assert torch.allclose(ssim.ssim_map(x, x), torch.ones_like(x), atol=1e-12)
    assert abs(float(ssim.ssim_loss(x, x))) < 1e-12


def test_symmetric_in_its_arguments_and_bounded():
    x, y = _pair(1)
    a, b = ssim.ssim_map(x, y), ssim.ssim_map(y, x)
    assert torch.allclose(a, b, atol=1e-13)
    assert float(a.max()) <= 1.0 + 1e-12 and float(a.min()) >= -1.0


def test_convolution_form_matches_explicit_window_loop():
    x, y = _pair(2, n=2, h=14, w=17)
    S = ssim.ssim_map(x, y)
    for (n, py, px, ch) in [(0, 0, 0, 0), (1, 13, 16, 2), (0, 7, 8, 1), (1, 0, 9, 0), (0, 13, 3, 2), (1, 5, 0, 1)]:
        assert abs(float(S[n, py, px, ch]) - ssim.ssim_value_loop(x, y, n, py, px, ch)) < 1e-12


def test_constant_images_closed_form():
    # constant a vs constant b far from the border: sigma terms vanish, SSIM = (2ab + C1) / (a^2 + b^2 + C1)
    a, b = 0.3, 0.7
    x = torch.full((1, 31, 31, 3), a)
    y = torch.full((1, 31, 31, 3), b)
    S = ssim.ssim_map(x, y)
    want = (2 * a * b + ssim.C1) / (a * a + b * b + ssim.C1)
    assert abs(float(S[0, 15, 15, 0]) - want) < 1e-12


def test_loss_gradient_matches_finite_differences():
    x, y = _pair(3, n=1, h=13, w=12)
    x = x.requires_grad_(True)
    L = ssim.ssim_loss(x, y, 0.0, 1.0)
    (gx,) = torch.autograd.grad(L, x)
    g = torch.Generator().manual_seed(9)
    for _ in range(6):
        d = torch.randn(x.shape, generator=g)
        eps = 1e-6
        fd = (float(ssim.ssim_loss(x.detach() + eps * d, y, 0.0, 1.0)) - float(ssim.ssim_loss(x.detach() - eps * d, y, 0.0, 1.0))) / (2 * eps)
        assert abs(fd - float((gx * d).sum())) <= 1e-6 * max(1.0, abs(fd))
