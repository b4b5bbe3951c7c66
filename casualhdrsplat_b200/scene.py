"""Deterministic synthetic scenes for tests and benchmarks (decision D2).

The reference ships no data loaders and no datasets (``/root/reference/Readme.md:57``), so every
run uses these seeded scenes.  All draws come from CPU ``torch.Generator``s in float64 and are cast
to fp32 once, so the CPU oracle and the CUDA path see identical bits.

Seeds: scene 0, trajectory 1, upstream gradient 2, CRF 3.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Optional

import torch

SPLINE_LINEAR = 0
SPLINE_CUBIC = 1
CRF_IDENTITY = 0
CRF_MLP = 1
CRF_LUT = 2


@dataclasses.dataclass
class Scene:
    means: torch.Tensor          # [N,3] fp32 world
    quats: torch.Tensor          # [N,4] fp32 wxyz, unit
    scales: torch.Tensor         # [N,3] fp32 >0
    opacities: torch.Tensor      # [N]   fp32 in (0,1)
    colors: torch.Tensor         # [N,3] fp32 linear HDR radiance
    knots: torch.Tensor          # [K,7] fp32 camera-to-world (t, q wxyz)
    knot_t0: float
    knot_dt: float
    frame_times: torch.Tensor    # [B] fp32
    exposure_times: torch.Tensor  # [B] fp32
    spline_kind: int
    Ks: torch.Tensor             # [B,3,3] fp32
    width: int
    height: int
    n_virtual: int
    crf_kind: int
    crf_params: Optional[torch.Tensor]  # [3, 3*Hd+1] fp32 or None
    v_ldr: torch.Tensor          # [B,H,W,3] fp32 upstream gradient (loss = sum(v_ldr * ldr))
    name: str = ""

    @property
    def n_frames(self) -> int:
        return int(self.frame_times.shape[0])

    def spline(self) -> dict:
        return {"knots": self.knots, "knot_t0": self.knot_t0, "knot_dt": self.knot_dt,
                "frame_times": self.frame_times, "kind": self.spline_kind}

    def to(self, device) -> "Scene":
        kw = {}
        for f in dataclasses.fields(self):
            v = getattr(self, f.name)
            kw[f.name] = v.to(device) if isinstance(v, torch.Tensor) else v
        return Scene(**kw)


def _quat_mul(a, b):
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], -1)


def _quat_rot(q, v):
    w, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                     2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).reshape(3, 3)
    return R @ v


def _se3_exp_qt(xi):
    """xi = (rho, phi) float64 -> (q wxyz, t). Plain closed form; only used to lay out knots."""
    rho, phi = xi[:3], xi[3:]
    th = float(phi.norm())
    if th < 1e-12:
        return torch.tensor([1.0, 0, 0, 0], dtype=torch.float64), rho.clone()
    ax = phi / th
    q = torch.cat([torch.tensor([math.cos(th / 2)], dtype=torch.float64), ax * math.sin(th / 2)])
    K = torch.tensor([[0, -phi[2], phi[1]], [phi[2], 0, -phi[0]], [-phi[1], phi[0], 0]], dtype=torch.float64)
    V = torch.eye(3, dtype=torch.float64) + (1 - math.cos(th)) / th**2 * K + (th - math.sin(th)) / th**3 * (K @ K)
    return q, V @ rho


def gamma_lut_params(knots: int = 256, seed: int = 3, gamma: float = 2.2, z_min: float = -10.0, z_max: float = 1.5) -> torch.Tensor:
    """LUT-CRF parameters [3, L+2] = [z_min | z_max | v_0..v_{L-1}] sampling X**(1/gamma) (clamped to [0, 1]) at L log-exposure
    knots, with a small seeded per-channel gain so the channels differ as a calibrated table's do."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    zk = torch.linspace(z_min, z_max, knots, dtype=torch.float64)
    out = []
    for ch in range(3):
        gain = 1.0 + 0.1 * (torch.rand((), generator=g, dtype=torch.float64) - 0.5)
        v = (gain * torch.exp(zk / gamma)).clamp(0.0, 1.0)
        out.append(torch.cat([torch.tensor([z_min, z_max], dtype=torch.float64), v]))
    return torch.stack(out).to(torch.float32)


def gamma_crf_params(hidden: int = 64, seed: int = 3, gamma: float = 2.2) -> torch.Tensor:
    """MLP-CRF parameters [3, 3*Hd+1] whose response approximates X**(1/gamma) (closed form, no training).

    A ReLU layer represents a piecewise-linear function of z = ln(X + 1e-5); the hinge positions are
    spread over z in [-10, 1.5], the output weights are the slope changes of logit(exp(z/gamma))
    (capped), and a small seeded perturbation makes the three channels distinct, as a learned CRF is.
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    zk = torch.linspace(-10.0, 1.5, hidden, dtype=torch.float64)

    def target(z):
        y = torch.exp(z / gamma).clamp(1e-4, 0.998)
        return torch.log(y / (1 - y))

    tk = target(zk)
    slopes = (tk[1:] - tk[:-1]) / (zk[1:] - zk[:-1])
    slopes = torch.cat([slopes, slopes[-1:]])
    dsl = torch.cat([slopes[:1], slopes[1:] - slopes[:-1]])
    out = []
    for ch in range(3):
        gain = 1.0 + 0.25 * (torch.rand(hidden, generator=g, dtype=torch.float64) - 0.5)
        w1 = gain
        b1 = -zk * gain
        w2 = dsl / gain * (1.0 + 0.02 * (ch - 1))
        b2 = tk[:1] + 0.05 * (ch - 1)
        out.append(torch.cat([w1, b1, w2, b2]))
    return torch.stack(out).to(torch.float32)


def make_scene(n_gauss: int, width: int, height: int, n_frames: int = 1, n_virtual: int = 1,
               spline_kind: int = SPLINE_CUBIC, crf_kind: int = CRF_MLP, crf_hidden: int = 64,
               unit_exposure: bool = False, static_camera: bool = False, scale_mult: float = 1.0, name: str = "",
               scene_seed: int = 0, traj_seed: int = 1, grad_seed: int = 2, crf_seed: int = 3) -> Scene:
    """Build the D2 scene: Gaussians placed through the mid pose's frustum, a moving spline camera,
    log-uniform exposures, a gamma-like MLP CRF and an N(0,1) upstream gradient."""
    f64 = torch.float64
    g0 = torch.Generator(device="cpu").manual_seed(scene_seed)
    g1 = torch.Generator(device="cpu").manual_seed(traj_seed)
    g2 = torch.Generator(device="cpu").manual_seed(grad_seed)
    B, n = n_frames, n_virtual
    fx = fy = 0.8 * width
    cx, cy = width / 2.0, height / 2.0
    K = torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=f64)
    Ks = K[None].repeat(B, 1, 1)

    # ---- exposure ----
    if unit_exposure:
        exposure = torch.ones(B, dtype=f64)
    else:
        lo, hi = math.log(1 / 500), math.log(1 / 30)
        exposure = torch.exp(lo + (hi - lo) * torch.rand(B, generator=g1, dtype=f64))

    # ---- trajectory knots (camera-to-world) ----
    base_q = torch.tensor([0.9990482, 0.0, 0.0436194, 0.0], dtype=f64)  # 5 deg about y
    base_t = torch.tensor([0.3, -0.2, 0.5], dtype=f64)
    if spline_kind == SPLINE_LINEAR:
        n_knots = B + 1
        knot_dt = float(exposure[0]) if B == 1 else 1.0 / 30
        xi = torch.tensor([0.05, 0.0, 0.0, 0.0, math.radians(0.5), 0.0], dtype=f64)
        frame_times = (torch.arange(B, dtype=f64) + 0.5) * knot_dt
        mid_index = None
    else:
        n_knots = B + 4
        knot_dt = 1.0 / 30
        xi = torch.tensor([0.02, 0.005, 0.01, 0.002, 0.004, 0.001], dtype=f64)
        frame_times = (torch.arange(B, dtype=f64) + 1.5) * knot_dt
    if static_camera:
        xi = torch.zeros(6, dtype=f64)
    qs, ts = [base_q], [base_t]
    for _ in range(n_knots - 1):
        jitter = 1.0 + 0.1 * torch.randn(6, generator=g1, dtype=f64)
        dq, dtv = _se3_exp_qt(xi * jitter)
        ts.append(ts[-1] + _quat_rot(qs[-1], dtv))
        qs.append(_quat_mul(qs[-1], dq))
    knots = torch.cat([torch.stack(ts), torch.stack(qs)], dim=1)

    # ---- mid pose of the trajectory: Gaussians are laid out in its frame ----
    mid = (n_knots - 1) / 2.0
    i0 = int(math.floor(mid))
    i1 = min(i0 + 1, n_knots - 1)
    w = mid - i0
    q_mid = qs[i0] * (1 - w) + qs[i1] * w
    q_mid = q_mid / q_mid.norm()
    t_mid = ts[i0] * (1 - w) + ts[i1] * w

    # ---- Gaussians ----
    z = 2.0 + 10.0 * torch.rand(n_gauss, generator=g0, dtype=f64)
    u = (-0.05 + 1.10 * torch.rand(n_gauss, generator=g0, dtype=f64)) * width
    v = (-0.05 + 1.10 * torch.rand(n_gauss, generator=g0, dtype=f64)) * height
    p_cam = torch.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], dim=1)
    wq, xq, yq, zq = q_mid.unbind(-1)
    R_mid = torch.stack([1 - 2 * (yq * yq + zq * zq), 2 * (xq * yq - wq * zq), 2 * (xq * zq + wq * yq),
                         2 * (xq * yq + wq * zq), 1 - 2 * (xq * xq + zq * zq), 2 * (yq * zq - wq * xq),
                         2 * (xq * zq - wq * yq), 2 * (yq * zq + wq * xq), 1 - 2 * (xq * xq + yq * yq)]).reshape(3, 3)
    means = p_cam @ R_mid.T + t_mid
    ls_lo, ls_hi = math.log(0.004), math.log(0.04)
    scales = scale_mult * torch.exp(ls_lo + (ls_hi - ls_lo) * torch.rand(n_gauss, 3, generator=g0, dtype=f64))
    quats = torch.randn(n_gauss, 4, generator=g0, dtype=f64)
    quats = quats / quats.norm(dim=1, keepdim=True)
    opacities = 0.05 + 0.65 * torch.rand(n_gauss, generator=g0, dtype=f64)
    colors = torch.exp(1.5 * torch.randn(n_gauss, 3, generator=g0, dtype=f64))

    crf_params = gamma_crf_params(crf_hidden, crf_seed) if crf_kind == CRF_MLP else gamma_lut_params(crf_hidden, crf_seed) if crf_kind == CRF_LUT else None
    v_ldr = torch.randn(B, height, width, 3, generator=g2, dtype=f64)
    f32 = torch.float32
    return Scene(means=means.to(f32), quats=quats.to(f32), scales=scales.to(f32), opacities=opacities.to(f32),
                 colors=colors.to(f32), knots=knots.to(f32), knot_t0=0.0, knot_dt=float(knot_dt),
                 frame_times=frame_times.to(f32), exposure_times=exposure.to(f32), spline_kind=spline_kind,
                 Ks=Ks.to(f32), width=width, height=height, n_virtual=n, crf_kind=crf_kind, crf_params=crf_params,
                 v_ldr=v_ldr.to(f32), name=name)


# BASELINE.json configs (index = position in BASELINE.json "configs")
CONFIGS = {
    "c1": dict(n_gauss=10_000, width=256, height=256, n_frames=1, n_virtual=1, spline_kind=SPLINE_LINEAR,
               crf_kind=CRF_IDENTITY, unit_exposure=True, static_camera=True),
    "c2": dict(n_gauss=100_000, width=800, height=800, n_frames=1, n_virtual=4, spline_kind=SPLINE_LINEAR,
               crf_kind=CRF_MLP),
    "c3": dict(n_gauss=1_000_000, width=1920, height=1080, n_frames=1, n_virtual=8, spline_kind=SPLINE_CUBIC,
               crf_kind=CRF_MLP),
    "c4": dict(n_gauss=1_000_000, width=1920, height=1080, n_frames=8, n_virtual=8, spline_kind=SPLINE_CUBIC,
               crf_kind=CRF_MLP),
    "c5": dict(n_gauss=3_000_000, width=3840, height=2160, n_frames=8, n_virtual=16, spline_kind=SPLINE_CUBIC,
               crf_kind=CRF_MLP),
    # small cases for tests
    "tiny": dict(n_gauss=400, width=64, height=48, n_frames=2, n_virtual=3, spline_kind=SPLINE_CUBIC,
                 crf_kind=CRF_MLP, crf_hidden=16, scale_mult=12.0),
    "small": dict(n_gauss=5_000, width=200, height=136, n_frames=2, n_virtual=3, spline_kind=SPLINE_CUBIC,
                  crf_kind=CRF_MLP, crf_hidden=64, scale_mult=5.0),
}


def make_config(name: str, **overrides) -> Scene:
    kw = dict(CONFIGS[name])
    kw.update(overrides)
    return make_scene(name=name, **kw)
