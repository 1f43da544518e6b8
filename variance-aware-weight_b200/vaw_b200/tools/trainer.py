"""The pieces of tools/trainer.py that sit directly before / after the hot path.

    sample_from_latent(latent, latent_scale)   tools/trainer.py:21-25  -> K1's input x_start, one kernel, bit-exact
    ema / gradient clipping / optimizer step   tools/trainer.py:12-18,60-62,124-133 -> vaw_b200.optim.FusedAdamW

The Trainer class itself (logging, gradient accumulation loop, checkpoints) is control plane and stays the reference's.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib as L

L.register("vaw_sample_from_latent", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_float, C.c_void_p])


def sample_from_latent(latent, latent_scale=1.0, defer=False):
    """latent [N, 2C, H, W] = (mean | std): returns (mean + std * randn_like(mean)) * latent_scale.  The noise is drawn
    with torch.randn_like on the device, as the reference does, so a seeded run sees the same Philox stream.

    defer=True returns a `DeferredLatent` instead: `GaussianDiffusion.training_losses` accepts it as `x_start` and
    draws the same eps1 INSIDE its q_sample kernel (same generator stream, same values), so neither eps1 nor, for the
    eps objective, x_start itself is written to HBM."""
    L.require_cuda(latent)
    if defer:
        from .gaussian_diffusion import DeferredLatent
        return DeferredLatent(latent, latent_scale)
    if latent.dim() < 2 or latent.shape[1] % 2:
        raise ValueError("sample_from_latent expects [N, 2C, ...] moments")
    lat = latent.contiguous().float()
    n, c2 = lat.shape[0], lat.shape[1]
    shape = (n, c2 // 2) + tuple(lat.shape[2:])
    eps = torch.randn(shape, device=lat.device, dtype=lat.dtype)
    out = torch.empty(shape, device=lat.device, dtype=torch.float32)
    if out.numel():
        L.call("vaw_sample_from_latent", lat.data_ptr(), eps.data_ptr(), out.data_ptr(), n, out.numel() // max(n, 1),
               float(latent_scale), L.stream_ptr())
    return out
