"""vaw_b200 — B200-native (sm_100a) diffusion training step behind the Variance-Aware-Weight Python API.

Layout mirrors the reference's module paths for the hot path only:
    vaw_b200.tools.gaussian_diffusion   GaussianDiffusion / FlowMatching / compute_mse_loss_weight / compute_align_loss
    vaw_b200.tools.resample             ScheduleSampler family (device importance sampling)
    vaw_b200.models.{dit,uvit,unet}     denoisers with the reference constructors and parameter names
Everything numeric runs in libvaw_b200.so (see include/vaw_b200.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
