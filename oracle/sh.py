"""Real spherical harmonics (degree 0..3) for view-dependent colour — float64 oracle of csrc/chs_sh.cuh.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: the reference ships no code; this is
SURVEY.md section 8(f) row f2 (the step immediately before the path in a 3DGS trainer).
colour_ch = relu(0.5 + sum_k sh[k, ch] Y_k(dir)), dir = normalize(mean - campos), campos = -R^T t.
"""
import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658, 1.445305721320277,
      -0.5900435899266435]


def sh_basis(deg: int, d: torch.Tensor) -> torch.Tensor:
    """d [...,3] unit directions -> Y [..., (deg+1)^2]."""
    x, y, z = d.unbind(-1)
    Y = [torch.full_like(x, C0)]
    if deg >= 1:
        Y += [-C1 * y, C1 * z, -C1 * x]
    if deg >= 2:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        Y += [C2[0] * xy, C2[1] * yz, C2[2] * (2 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy)]
    if deg >= 3:
        Y += [C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy), C3[3] * z * (2 * zz - 3 * xx - 3 * yy),
              C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy), C3[6] * x * (xx - 3 * yy)]
    return torch.stack(Y, dim=-1)


def sh_colors(sh: torch.Tensor, means: torch.Tensor, viewmats: torch.Tensor, deg: int) -> torch.Tensor:
    """sh [N,K,3], means [N,3], viewmats [C,4,4] -> colours [C,N,3] (float64)."""
    sh, means, viewmats = sh.double(), means.double(), viewmats.double()
    R, t = viewmats[:, :3, :3], viewmats[:, :3, 3]
    campos = -(R.transpose(-1, -2) @ t[..., None])[..., 0]
    d = means[None] - campos[:, None]
    d = d / d.norm(dim=-1, keepdim=True)
    K = (deg + 1) ** 2
    Y = sh_basis(deg, d)
    return torch.relu(0.5 + torch.einsum("cnk,nkh->cnh", Y, sh[:, :K]))
