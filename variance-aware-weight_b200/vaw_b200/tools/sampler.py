"""Host mirror of the reference's guidance wrapper (tools/sampler.py:10-48) over the B200 library.

Only IntervalCFG lives on the reverse path this package accelerates; the VAE decode, the classifier
guidance and the FID plumbing of the reference's sampler module stay with the reference.
"""
from __future__ import annotations

import torch

from .. import _lib as L


class IntervalCFG(torch.nn.Module):
    """Classifier-free guidance applied inside a time interval (reference tools/sampler.py:10-48).

    Same constructor, same `forward(sample, time, **model_kwargs)`; the doubled-batch forward is the wrapped model's
    and the combine `uncond + scale * (cond - uncond)` is one vaw_cfg_combine launch instead of three elementwise ones.
    """

    def __init__(self, model, num_classes, guidance_scale=1.0, interval=(-1.0, -1.0), class_cond=True):
        super().__init__()
        self.model = model
        self.null_label = int(num_classes)
        self.guidance_scale = float(guidance_scale)
        self.interval = interval
        self.class_cond = class_cond

    def _use_cfg(self, time_value):
        if abs(self.guidance_scale - 1.0) < 1e-8:
            return False
        lo, hi = self.interval
        if lo >= 0 and hi > lo:
            return lo <= time_value < hi
        return True

    @staticmethod
    def _format_time(time_tensor, batch_size):
        hint = getattr(time_tensor, "_vaw_host_value", None)
        if time_tensor.dim() == 0:
            out = time_tensor.expand(batch_size)
        elif time_tensor.numel() == 1:
            out = time_tensor.reshape(1).expand(batch_size)
        else:
            out = time_tensor.reshape(batch_size)
        if hint is not None:
            out._vaw_host_value = hint
        return out

    def forward(self, sample_tensor, time_tensor, **model_kwargs):
        n = sample_tensor.shape[0]
        time_tensor = self._format_time(time_tensor, n)
        labels = model_kwargs.get("y", None)
        use = False
        if self.class_cond and labels is not None and abs(self.guidance_scale - 1.0) >= 1e-8:
            lo, hi = self.interval
            if not (lo >= 0 and hi > lo):
                use = True                   # no interval: the decision does not depend on the time at all
            else:
                # the samplers of this package build the time tensor from a host scalar and leave it attached
                # (`_vaw_host_value`): the interval test then costs no device synchronisation; any other caller gets the
                # reference's `time_tensor.float().mean().item()`
                tv = getattr(time_tensor, "_vaw_host_value", None)
                use = self._use_cfg(float(tv) if tv is not None else float(time_tensor.float().mean().item()))
        if not use:
            return self.model(sample_tensor, time_tensor, **model_kwargs)
        assert labels.shape[0] == n, f"CFG expects label batch size {n}, but got {labels.shape[0]}."
        kw = dict(model_kwargs)
        kw["y"] = torch.cat([labels, torch.full_like(labels, self.null_label)], dim=0)
        out = self.model(torch.cat([sample_tensor, sample_tensor], dim=0), torch.cat([time_tensor, time_tensor], dim=0),
                         **kw)
        out = out[0] if isinstance(out, tuple) else out
        L.require_cuda(out)
        if out.dtype not in (torch.float32, torch.bfloat16):
            out = out.float()
        out = out.contiguous()
        guided = torch.empty_like(out[:n])
        L.call("vaw_cfg_combine", out.data_ptr(), guided.data_ptr(), L.BF16 if out.dtype == torch.bfloat16 else L.F32,
               self.guidance_scale, guided.numel(), L.stream_ptr())
        return guided
