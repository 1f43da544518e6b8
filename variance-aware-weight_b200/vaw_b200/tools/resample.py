"""Timestep importance sampling on the device (K3), behind the reference's ScheduleSampler API.

Mirror of /root/reference/tools/resample.py:
    create_named_schedule_sampler (:9-21), ScheduleSampler.sample (:43-59), UniformSampler (:62-68),
    LossAwareSampler.update_with_local_losses (:72-112), LossSecondMomentResampler (:132-162).

What changed underneath:
  * weights -> p -> CDF -> inverse sample runs in one kernel (vaw_sampler_sample) that reproduces numpy's fp64
    arithmetic operation by operation, so timestep indices are bit-identical to `np.random.choice(T, B, p=p)`;
    the uniforms are still drawn from numpy's global MT19937 on the host (`np.random.random_sample(B)` consumes
    exactly the stream `choice` would), so seeds stay interchangeable with the reference.
  * the loss history lives on the device (fp64 [T, H] + int32 counts); update_with_local_losses does ONE
    all_gather of packed (int32 t, fp32 loss) pairs and a device kernel instead of three collectives and
    2*W*B `.item()` host syncs.  `_loss_history` / `_loss_counts` are exposed as numpy views for parity checks.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np
import torch as th
import torch.distributed as dist

from .. import _lib as L
from ..parallel import gather_tloss


def create_named_schedule_sampler(name: str, diffusion):
    if name == "uniform":
        return UniformSampler(diffusion)
    if name == "loss-second-moment":
        return LossSecondMomentResampler(diffusion)
    raise NotImplementedError(f"unknown schedule sampler: {name}")


class ScheduleSampler(ABC):
    @abstractmethod
    def weights(self):
        """numpy float64 [T] weights (not necessarily normalised)."""

    def _device_sample(self, batch_size, device, mode, w_dev, history, counts, H, uniform_prob, T):
        device = th.device(device)
        if device.type != "cuda":
            raise L.VawError("ScheduleSampler.sample needs a CUDA device (no CPU fallback)")
        u = np.random.random_sample(batch_size)  # host MT19937, identical stream consumption to np.random.choice
        u_dev = th.from_numpy(u).to(device, non_blocking=True)
        idx = th.empty(batch_size, dtype=th.int64, device=device)
        w = th.empty(batch_size, dtype=th.float32, device=device)
        L.call("vaw_sampler_sample", mode, L.ptr(w_dev), L.ptr(history), L.ptr(counts), T, H, float(uniform_prob),
               u_dev.data_ptr(), batch_size, idx.data_ptr(), w.data_ptr(), None, None, None, L.stream_ptr())
        return idx, w

    def sample(self, batch_size, device):
        """-> (timesteps int64 [B], importance weights float32 [B]) on `device` (reference :43-59)."""
        w = np.ascontiguousarray(self.weights(), dtype=np.float64)
        w_dev = th.from_numpy(w).to(device)
        return self._device_sample(batch_size, device, 0, w_dev, None, None, 0, 0.0, len(w))


class UniformSampler(ScheduleSampler):
    def __init__(self, diffusion):
        self.diffusion = diffusion
        self._weights = np.ones([diffusion.num_timesteps])
        self._w_dev = {}

    def weights(self):
        return self._weights

    def sample(self, batch_size, device):
        key = str(device)
        if key not in self._w_dev:
            self._w_dev[key] = th.from_numpy(self._weights).to(device)
        return self._device_sample(batch_size, device, 0, self._w_dev[key], None, None, 0, 0.0, len(self._weights))


class LossAwareSampler(ScheduleSampler):
    def update_with_local_losses(self, local_ts, local_losses):
        """Gather (t, loss) from every rank in rank order and apply the identical update everywhere
        (reference :72-112)."""
        L.require_cuda(local_ts, local_losses)
        ts, losses = gather_tloss(local_ts, local_losses, ragged=self.ragged_batches)
        self._device_update(ts, losses)

    # True (default): ranks may pass different batch sizes - the sizes are exchanged and padded to the largest, like the
    # reference (:85-100).  A training loop whose loader drops the last partial batch (equal sizes on every rank) can set
    # this to False and save the size exchange with its host synchronisation.
    ragged_batches = True

    @abstractmethod
    def update_with_all_losses(self, ts, losses):
        """Apply a list of (t, loss) pairs in order (reference :114-129)."""


class LossSecondMomentResampler(LossAwareSampler):
    def __init__(self, diffusion, history_per_term=10, uniform_prob=0.001):
        self.diffusion = diffusion
        self.history_per_term = history_per_term
        self.uniform_prob = uniform_prob
        self._T = diffusion.num_timesteps
        self._hist_dev = None    # fp64 [T, H]
        self._count_dev = None   # int32 [T]
        self._host_hist = np.zeros([self._T, history_per_term], dtype=np.float64)
        self._host_counts = np.zeros([self._T], dtype=int)

    # -- device state ---------------------------------------------------------------------------------
    def _ensure_device(self, device):
        device = th.device(device)
        if device.type != "cuda":
            raise L.VawError("LossSecondMomentResampler keeps its history on a CUDA device (no CPU fallback)")
        if device.index is None:   # "cuda" and "cuda:<current>" are the same place; compare like with like
            device = th.device("cuda", th.cuda.current_device())
        if self._hist_dev is None or self._hist_dev.device != device:
            self._sync_host()      # moving devices carries the history along
            self._hist_dev = th.from_numpy(self._host_hist).to(device).contiguous()
            self._count_dev = th.from_numpy(self._host_counts.astype(np.int32)).to(device).contiguous()

    def _sync_host(self):
        if self._hist_dev is not None:
            self._host_hist = self._hist_dev.cpu().numpy()
            self._host_counts = self._count_dev.cpu().numpy().astype(int)

    @property
    def _loss_history(self):
        self._sync_host()
        return self._host_hist

    @property
    def _loss_counts(self):
        self._sync_host()
        return self._host_counts

    def load_history(self, history, counts, device):
        """Install a pre-filled history (used to benchmark / test the warmed-up branch)."""
        self._host_hist = np.ascontiguousarray(history, dtype=np.float64)
        self._host_counts = np.asarray(counts).astype(int)
        self._hist_dev = None
        self._ensure_device(device)

    # -- reference API --------------------------------------------------------------------------------
    def _warmed_up(self):
        return bool((self._loss_counts == self.history_per_term).all())

    def weights(self):
        """numpy float64 [T]; computed by the kernel so it is the exact vector the sampler uses."""
        if self._hist_dev is None:
            if not (self._host_counts == self.history_per_term).all():
                return np.ones([self._T], dtype=np.float64)
            raise L.VawError("weights() of a warmed-up sampler needs its device state (call sample() first)")
        dev = self._hist_dev.device
        w = th.empty(self._T, dtype=th.float64, device=dev)
        L.call("vaw_sampler_sample", 1, None, self._hist_dev.data_ptr(), self._count_dev.data_ptr(), self._T,
               self.history_per_term, float(self.uniform_prob), None, 0, None, None, w.data_ptr(), None, None,
               L.stream_ptr())
        return w.cpu().numpy()

    def sample(self, batch_size, device):
        if type(self).weights is not LossSecondMomentResampler.weights:
            return ScheduleSampler.sample(self, batch_size, device)   # a subclass redefined weights(): honour it (:52)
        self._ensure_device(device)
        return self._device_sample(batch_size, device, 1, None, self._hist_dev, self._count_dev,
                                   self.history_per_term, self.uniform_prob, self._T)

    def _device_update(self, ts32, losses):
        self._ensure_device(ts32.device)
        L.call("vaw_sampler_update", self._hist_dev.data_ptr(), self._count_dev.data_ptr(), ts32.data_ptr(),
               losses.data_ptr(), ts32.numel(), self._T, self.history_per_term, L.stream_ptr())

    def update_with_all_losses(self, ts, losses):
        if self._hist_dev is None:
            raise L.VawError("update_with_all_losses needs the device state (call sample() or load_history() first)")
        dev = self._hist_dev.device
        t32 = th.tensor(list(ts), dtype=th.int32, device=dev)
        # python floats -> fp32 is exact for values that came from an fp32 loss tensor, as in the reference's usage
        l32 = th.tensor(list(losses), dtype=th.float64, device=dev).float()
        self._device_update(t32, l32)
