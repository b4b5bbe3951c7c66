"""SE(3) Lie-group helpers and trajectory splines for the CPU oracle (float64, torch autograd).

TEST INFRASTRUCTURE ONLY. Nothing under ``oracle/`` is imported by the product path
(``casualhdrsplat_b200``); only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it.

PARITY UNPINNED: the reference repository ships no code (``/root/reference/Readme.md:57``
"Still working on...."), so there is no reference implementation, golden vector or test to
pin this restatement against.  What it restates is the formation model described in
``/root/reference/Readme.md:54`` and ``/root/reference/assets/pipeline.png`` (legend:
"Trajectory control knots", "Camera motion spline", "Virtual camera pose", "Exposure time
range"), completed by SURVEY.md Appendix A.2 [D3, D4].

Conventions: quaternions are ``wxyz``; spline knots are camera-to-world poses stored raw as
``[K, 7] = (tx, ty, tz, qw, qx, qy, qz)``; quaternions are normalised inside.
"""
from __future__ import annotations

import torch

# Small-angle threshold shared with the CUDA spline kernel (csrc/chs_spline.cuh: CHS_SMALL_ANGLE).
SMALL_ANGLE = 1e-4


def quat_normalize(q: torch.Tensor) -> torch.Tensor:
    return q / q.norm(dim=-1, keepdim=True)


def quat_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ],
        dim=-1,
    )


def quat_conj(q: torch.Tensor) -> torch.Tensor:
    return torch.cat([q[..., :1], -q[..., 1:]], dim=-1)


def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    """Rotation matrix of a (not necessarily unit) wxyz quaternion; normalised inside."""
    q = quat_normalize(q)
    w, x, y, z = q.unbind(-1)
    R = torch.stack(
        [
            1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y),
        ],
        dim=-1,
    )
    return R.reshape(q.shape[:-1] + (3, 3))


def hat(v: torch.Tensor) -> torch.Tensor:
    x, y, z = v.unbind(-1)
    o = torch.zeros_like(x)
    return torch.stack([o, -z, y, z, o, -x, -y, x, o], dim=-1).reshape(v.shape[:-1] + (3, 3))


def _abc(theta2: torch.Tensor):
    """A = sin t / t, B = (1 - cos t) / t^2, C = (t - sin t) / t^3 with Taylor branches below SMALL_ANGLE."""
    small = theta2 < SMALL_ANGLE * SMALL_ANGLE
    t2 = torch.where(small, torch.ones_like(theta2), theta2)
    t = torch.sqrt(t2)
    A = torch.where(small, 1 - theta2 / 6, torch.sin(t) / t)
    B = torch.where(small, 0.5 - theta2 / 24, (1 - torch.cos(t)) / t2)
    C = torch.where(small, 1.0 / 6 - theta2 / 120, (t - torch.sin(t)) / (t2 * t))
    return A, B, C


def so3_exp(phi: torch.Tensor) -> torch.Tensor:
    theta2 = (phi * phi).sum(-1)
    A, B, _ = _abc(theta2)
    K = hat(phi)
    eye = torch.eye(3, dtype=phi.dtype).expand(K.shape)
    return eye + A[..., None, None] * K + B[..., None, None] * (K @ K)


def se3_V(phi: torch.Tensor) -> torch.Tensor:
    theta2 = (phi * phi).sum(-1)
    _, B, C = _abc(theta2)
    K = hat(phi)
    eye = torch.eye(3, dtype=phi.dtype).expand(K.shape)
    return eye + B[..., None, None] * K + C[..., None, None] * (K @ K)


def se3_Vinv(phi: torch.Tensor) -> torch.Tensor:
    theta2 = (phi * phi).sum(-1)
    small = theta2 < SMALL_ANGLE * SMALL_ANGLE
    t2 = torch.where(small, torch.ones_like(theta2), theta2)
    t = torch.sqrt(t2)
    D = torch.where(
        small,
        1.0 / 12 + theta2 / 720,
        1.0 / t2 - (1 + torch.cos(t)) / (2 * t * torch.sin(t)),
    )
    K = hat(phi)
    eye = torch.eye(3, dtype=phi.dtype).expand(K.shape)
    return eye - 0.5 * K + D[..., None, None] * (K @ K)


def so3_log_quat(q: torch.Tensor) -> torch.Tensor:
    """Rotation vector of a unit quaternion; the w >= 0 representative is used (shortest path)."""
    q = torch.where(q[..., :1] < 0, -q, q)
    w = q[..., 0]
    v = q[..., 1:]
    s2 = (v * v).sum(-1)
    small = s2 < SMALL_ANGLE * SMALL_ANGLE
    s = torch.sqrt(torch.where(small, torch.ones_like(s2), s2))
    k_big = 2 * torch.atan2(s, w) / s
    k_small = (2 / w) * (1 - s2 / (3 * w * w))
    k = torch.where(small, k_small, k_big)
    return v * k[..., None]


def se3_exp(rho: torch.Tensor, phi: torch.Tensor):
    """(R, t) = Exp(rho, phi): R = so3_exp(phi), t = V(phi) rho."""
    return so3_exp(phi), (se3_V(phi) @ rho[..., None])[..., 0]


def se3_rel_log(ta, qa, tb, qb):
    """(rho, phi) = Log(T_a^-1 T_b) for poses given as translation + (unnormalised) quaternion."""
    qa = quat_normalize(qa)
    qb = quat_normalize(qb)
    Ra = quat_to_rotmat(qa)
    q_rel = quat_mul(quat_conj(qa), qb)
    t_rel = (Ra.transpose(-1, -2) @ (tb - ta)[..., None])[..., 0]
    phi = so3_log_quat(q_rel)
    rho = (se3_Vinv(phi) @ t_rel[..., None])[..., 0]
    return rho, phi


def compose(Ra, ta, Rb, tb):
    return Ra @ Rb, (Ra @ tb[..., None])[..., 0] + ta


SPLINE_LINEAR = 0
SPLINE_CUBIC = 1


def spline_segment(times: torch.Tensor, knot_t0: float, knot_dt: float, n_knots: int, kind: int):
    """Segment index s (long, no gradient) and local parameter u (carries d/dt) for each sample time."""
    x = (times - knot_t0) / knot_dt
    if kind == SPLINE_LINEAR:
        lo, hi = 0, n_knots - 2
    else:
        lo, hi = 1, n_knots - 3
    if hi < lo:
        raise ValueError("not enough knots for this spline kind")
    s = torch.floor(x.detach()).long().clamp(lo, hi)
    u = x - s.to(x.dtype)
    return s, u


def spline_c2w(knots: torch.Tensor, knot_t0: float, knot_dt: float, times: torch.Tensor, kind: int):
    """Camera-to-world (R [T,3,3], t [T,3]) of the spline at each sample time."""
    K = knots.shape[0]
    s, u = spline_segment(times, knot_t0, knot_dt, K, kind)
    tr = knots[:, :3]
    qu = knots[:, 3:]
    if kind == SPLINE_LINEAR:
        ta, qa, tb, qb = tr[s], qu[s], tr[s + 1], qu[s + 1]
        rho, phi = se3_rel_log(ta, qa, tb, qb)
        Rd, td = se3_exp(u[:, None] * rho, u[:, None] * phi)
        return compose(quat_to_rotmat(qa), ta, Rd, td)
    if kind != SPLINE_CUBIC:
        raise ValueError(f"unknown spline kind {kind}")
    u2 = u * u
    u3 = u2 * u
    basis = [
        (5 + 3 * u - 3 * u2 + u3) / 6,
        (1 + 3 * u + 3 * u2 - 2 * u3) / 6,
        u3 / 6,
    ]
    R, t = quat_to_rotmat(qu[s - 1]), tr[s - 1]
    for j in range(3):
        ia, ib = s - 1 + j, s + j
        rho, phi = se3_rel_log(tr[ia], qu[ia], tr[ib], qu[ib])
        Rd, td = se3_exp(basis[j][:, None] * rho, basis[j][:, None] * phi)
        R, t = compose(R, t, Rd, td)
    return R, t


def sample_times(frame_times: torch.Tensor, exposure: torch.Tensor, n_virtual: int) -> torch.Tensor:
    """t_{i,k} = t_i + (k/(n-1) - 1/2) dt_i  (n > 1);  t_i for n == 1.  Flattened [B*n], camera c = i*n + k."""
    if n_virtual == 1:
        return frame_times.clone()
    k = torch.arange(n_virtual, dtype=frame_times.dtype)
    w = k / (n_virtual - 1) - 0.5
    return (frame_times[:, None] + w[None, :] * exposure[:, None]).reshape(-1)


def c2w_to_viewmat(R: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    Rv = R.transpose(-1, -2)
    tv = -(Rv @ t[..., None])[..., 0]
    top = torch.cat([Rv, tv[..., None]], dim=-1)
    bottom = torch.tensor([0, 0, 0, 1], dtype=R.dtype).expand(R.shape[:-2] + (1, 4))
    return torch.cat([top, bottom], dim=-2)


def spline_viewmats(knots, knot_t0, knot_dt, frame_times, exposure, n_virtual, kind):
    """viewmats [B*n, 4, 4] (world-to-camera) of the virtual cameras inside each exposure window."""
    times = sample_times(frame_times, exposure, n_virtual)
    R, t = spline_c2w(knots, knot_t0, knot_dt, times, kind)
    return c2w_to_viewmat(R, t)
