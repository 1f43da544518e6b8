"""Host mirror of the reference's timestep respacing (tools/respace.py:9-128) for the reverse path.

`SpacedDiffusion` keeps a subset of the base process's timesteps, rebuilds betas so that the cumulative products at
the kept steps are unchanged, and hands the model the ORIGINAL timestep of every kept step.  All of it is host-side
integer / float64 work done once; the per-step arithmetic stays in vaw_reverse_step.
"""
from __future__ import annotations

import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """Reference :9-62.  "ddimN" -> the fixed integer stride that yields exactly N steps; otherwise a list (or
    comma-separated string) of per-section counts, each section strided evenly (fractional stride, accumulated and
    rounded half-to-even like the reference)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            wanted = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                kept = range(0, num_timesteps, stride)
                if len(kept) == wanted:
                    return set(kept)
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(v) for v in section_counts.split(",")]
    base, extra = divmod(num_timesteps, len(section_counts))
    kept, start = [], 0
    for k, count in enumerate(section_counts):
        size = base + (1 if k < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            kept.append(start + round(pos))
            pos += stride
        start += size
    return set(kept)


class SpacedDiffusion(GaussianDiffusion):
    """Reference :65-115: a diffusion process that skips steps of a base process."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(kwargs["betas"])
        base = GaussianDiffusion(**kwargs)
        self.timestep_map = [i for i in range(self.original_num_steps) if i in self.use_timesteps]
        kept = base.alphas_cumprod[self.timestep_map]
        prev = np.concatenate([[1.0], kept[:-1]])
        kwargs["betas"] = 1 - kept / prev          # float64, same expression per kept step as the reference's loop
        super().__init__(**kwargs)

    def p_mean_variance(self, model, *args, **kwargs):
        return super().p_mean_variance(self._wrap_model(model), *args, **kwargs)

    def training_losses(self, model, *args, **kwargs):
        return super().training_losses(self._wrap_model(model), *args, **kwargs)

    def _reverse(self, mode, model, *args, **kwargs):
        return super()._reverse(mode, self._wrap_model(model), *args, **kwargs)

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps)

    def _scale_timesteps(self, t):
        return t   # scaling is done by the wrapped model (:113-115)


class _WrappedModel:
    """Reference :118-128: maps the spaced step index to the base process's timestep before calling the model."""

    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps):
        self.model = model
        self.timestep_map = timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps
        self._maps = {}

    def parameters(self):
        return self.model.parameters()

    def __call__(self, x, ts, **kwargs):
        key = (ts.device, ts.dtype)
        m = self._maps.get(key)
        if m is None:
            m = self._maps[key] = th.tensor(self.timestep_map, device=ts.device, dtype=ts.dtype)
        new_ts = m[ts]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        hint = getattr(ts, "_vaw_host_value", None)   # the sampling loops build ts from a host integer: keep it known
        if hint is not None:                          # on the host so that IntervalCFG need not synchronise
            v = float(self.timestep_map[int(hint)])
            new_ts._vaw_host_value = v * (1000.0 / self.original_num_steps) if self.rescale_timesteps else v
        return self.model(x, new_ts, **kwargs)
