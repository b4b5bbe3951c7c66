"""TEST INFRASTRUCTURE ONLY — float64 CPU oracle of the D-SSIM photometric loss.

PARITY UNPINNED: the reference repository ships no code (Readme.md:57), so this restates the published 3DGS training
loss the family uses — L = (1 - lambda) L1 + lambda (1 - SSIM), SSIM with an 11x11 Gaussian window (sigma 1.5) applied
with zero padding per channel, C1 = 0.01^2, C2 = 0.03^2, mean over all values — in plain torch float64 with autograd.
Nothing under casualhdrsplat_b200/ may import this module.
"""
import torch
import torch.nn.functional as F

WINDOW = 11
SIGMA = 1.5
C1 = 0.01 ** 2
C2 = 0.03 ** 2


def gaussian_window():
    k = torch.arange(WINDOW, dtype=torch.float64) - WINDOW // 2
    g = torch.exp(-(k * k) / (2 * SIGMA * SIGMA))
    return g / g.sum()


def ssim_map(x, y):
    """x, y [n, H, W, 3] float64 -> SSIM per value [n, H, W, 3]."""
    g = gaussian_window()
    w2 = (g[:, None] * g[None, :])[None, None].expand(3, 1, WINDOW, WINDOW)
    xc, yc = x.permute(0, 3, 1, 2), y.permute(0, 3, 1, 2)

    def blur(t):
        return F.conv2d(t, w2, padding=WINDOW // 2, groups=3)

    mu1, mu2 = blur(xc), blur(yc)
    s11 = blur(xc * xc) - mu1 * mu1
    s22 = blur(yc * yc) - mu2 * mu2
    s12 = blur(xc * yc) - mu1 * mu2
    S = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s11 + s22 + C2))
    return S.permute(0, 2, 3, 1)


def ssim_loss(x, y, l1_weight=0.8, ssim_weight=0.2):
    """L = l1_weight * mean|x - y| + ssim_weight * (1 - mean SSIM(x, y))."""
    x, y = x.double(), y.double()
    return l1_weight * (x - y).abs().mean() + ssim_weight * (1.0 - ssim_map(x, y).mean())


def ssim_value_loop(x, y, n, py, px, ch):
    """SSIM of ONE output value by explicit loops over the window (cross-check of the convolution form)."""
    g = gaussian_window()
    H, W = x.shape[1], x.shape[2]
    mu1 = mu2 = e11 = e22 = e12 = 0.0
    for dy in range(WINDOW):
        for dx in range(WINDOW):
            yy, xx = py + dy - WINDOW // 2, px + dx - WINDOW // 2
            if 0 <= yy < H and 0 <= xx < W:
                w = float(g[dy] * g[dx])
                a, b = float(x[n, yy, xx, ch]), float(y[n, yy, xx, ch])
                mu1 += w * a; mu2 += w * b; e11 += w * a * a; e22 += w * b * b; e12 += w * a * b
    s11, s22, s12 = e11 - mu1 * mu1, e22 - mu2 * mu2, e12 - mu1 * mu2
    return ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s11 + s22 + C2))
