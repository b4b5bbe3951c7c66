"""The two streaming steps either side of the formation path in a trainer, through the
C ABI: the photometric loss with dL/dB emitted in the same pass, and Adam applied per section of the flat gradient buffer."""
from __future__ import annotations

import torch

from . import _lib
from .api import _stream

LOSS_L2, LOSS_L1 = 0, 1


def photometric_loss(ldr: torch.Tensor, target: torch.Tensor, kind: int = LOSS_L2, scale: float = 1.0, loss_acc: torch.Tensor = None):
    """Returns (v_ldr, loss_acc): v_ldr = dL/d ldr (same shape), loss_acc a device fp64 scalar the loss was ADDED to."""
    if not (ldr.is_cuda and target.is_cuda and ldr.dtype == torch.float32 and target.dtype == torch.float32):
        raise RuntimeError("photometric_loss: CUDA float32 tensors required (no CPU path)")
    ldr, target = ldr.contiguous(), target.contiguous()
    v = torch.empty_like(ldr)
    if loss_acc is None:
        loss_acc = torch.zeros((), dtype=torch.float64, device=ldr.device)
    _lib.check(_lib.lib().chs_loss(kind, _lib.ptr(ldr), _lib.ptr(target), ldr.numel(), float(scale), _lib.ptr(v), _lib.ptr(loss_acc),
                                   _stream()), "chs_loss")
    return v, loss_acc


def ssim_loss(ldr: torch.Tensor, target: torch.Tensor, l1_weight: float = 0.8, ssim_weight: float = 0.2, loss_acc: torch.Tensor = None):
    """The 3DGS photometric loss l1_weight * mean|d| + ssim_weight * (1 - mean SSIM) on frames [n, H, W, 3] (11x11 Gaussian
    window, zero padding).  Returns (v_ldr, loss_acc) like ``photometric_loss``."""
    if not (ldr.is_cuda and target.is_cuda and ldr.dtype == torch.float32 and target.dtype == torch.float32):
        raise RuntimeError("ssim_loss: CUDA float32 tensors required (no CPU path)")
    if ldr.dim() != 4 or ldr.shape[-1] != 3 or ldr.shape != target.shape:
        raise RuntimeError("ssim_loss: ldr and target must both be [n, H, W, 3]")
    ldr, target = ldr.contiguous(), target.contiguous()
    v = torch.empty_like(ldr)
    if loss_acc is None:
        loss_acc = torch.zeros((), dtype=torch.float64, device=ldr.device)
    work = torch.empty(3 * ldr.numel(), dtype=torch.float32, device=ldr.device)
    n, h, w, _ = ldr.shape
    _lib.check(_lib.lib().chs_ssim_loss(_lib.ptr(ldr), _lib.ptr(target), n, h, w, float(l1_weight), float(ssim_weight), _lib.ptr(v),
                                        _lib.ptr(loss_acc), _lib.ptr(work), work.numel() * 4, _stream()), "chs_ssim_loss")
    return v, loss_acc


class FlatAdam:
    """Adam over named parameter tensors whose gradients are views of the flat gradient buffer (GradLayout.views)."""

    def __init__(self, params: dict, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.params = {k: v for k, v in params.items()}
        self.lr = lr if isinstance(lr, dict) else {k: lr for k in params}
        self.betas, self.eps, self.step_count = betas, eps, 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def step(self, grads: dict, grad_scale: float = 1.0) -> None:
        self.step_count += 1
        L = _lib.lib()
        for k, p in self.params.items():
            g = grads[k]
            if not (p.is_contiguous() and g.is_contiguous()):
                raise RuntimeError(f"FlatAdam: parameter / gradient `{k}` must be contiguous")
            _lib.check(L.chs_adam_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(self.m[k]), _lib.ptr(self.v[k]), p.numel(), float(self.lr[k]),
                                       float(self.betas[0]), float(self.betas[1]), float(self.eps), self.step_count, float(grad_scale),
                                       _stream()), "chs_adam_step")
