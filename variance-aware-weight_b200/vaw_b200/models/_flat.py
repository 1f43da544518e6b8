"""Shared plumbing of the engine-backed models: all parameters live in ONE flat fp32 buffer (plus a bf16 shadow for
the tensor cores and a flat fp32 gradient buffer); the nn.Parameters the reference's trainer / optimizer / checkpoint
code sees are views of it."""
from __future__ import annotations

import copy

import torch
import torch.nn as nn

from .. import _lib as L


class ParamHolder(nn.Module):
    """Carries `weight` / `bias` so that state_dict keys match the reference; computes nothing itself."""

    def __init__(self, weight_shape, bias_shape=None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(weight_shape))
        if bias_shape is not None:
            self.bias = nn.Parameter(torch.empty(bias_shape))


class Slot(nn.Module):
    """Placeholder for the parameter-free entries of the reference's nn.Sequential (SiLU) so indices line up."""


class Named(nn.Module):
    pass


class FlatEngineModule(nn.Module):
    """Subclasses provide `_slots() -> ([(parameter, flat offset)], total elements)` and
    `_workspace_bytes(batch) -> int`."""

    def __init__(self):
        super().__init__()
        self._flat = None      # fp32 [n] leaf tensor holding every parameter
        self._shadow = None    # bf16 [n]
        self._gflat = None     # fp32 [n]
        self._shadow_version = -1
        self._ws = None
        self._ws_batch = -1
        self._fwd_serial = 0
        self._events = None
        self._post_backward = None
        self._before_cast = None   # DataParallel(shard_optimizer=True): completes the fp32 master before it is re-cast
        self._slot_cache = None

    def __deepcopy__(self, memo):
        """`copy.deepcopy(model)` is how the reference makes its EMA model (main.py:344).  The copy gets its own
        parameters (nn.Parameter.__deepcopy__ clones them) and re-packs them into a fresh flat buffer on its first
        forward; the activation workspace (tens of GB), CUDA events and data-parallel hooks are not carried over."""
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        drop = {"_flat": None, "_shadow": None, "_gflat": None, "_ws": None, "_events": None, "_post_backward": None,
                "_slot_cache": None, "_ws_batch": -1, "_shadow_version": -1, "_ws_inf": None, "_ws_inf_batch": -1,
                "_before_cast": None, "_fwd_wait": None}
        for k, v in self.__dict__.items():
            new.__dict__[k] = drop[k] if k in drop else copy.deepcopy(v, memo)
        return new

    def _ensure_flat(self, device):
        """(Re)pack the parameters into the flat buffer if they are not already views of it (after .to(), deepcopy,
        load_state_dict(assign=True) ...)."""
        slots, total = self._slots()
        ok = (self._flat is not None and self._flat.device == device and all(
            p.data_ptr() == self._flat.data_ptr() + 4 * o and p.device == device for p, o in slots))
        if ok:
            return
        flat = torch.zeros(total, dtype=torch.float32, device=device)
        gflat = torch.zeros(total, dtype=torch.float32, device=device)
        with torch.no_grad():
            for p, o in slots:
                n = p.numel()
                flat[o:o + n].copy_(p.detach().reshape(-1).to(device=device, dtype=torch.float32))
                old_grad = p.grad
                p.data = flat[o:o + n].view(p.shape)
                if old_grad is not None:
                    gflat[o:o + n].copy_(old_grad.reshape(-1).to(device=device, dtype=torch.float32))
                    p.grad = gflat[o:o + n].view(p.shape)
        flat.requires_grad_(True)
        self._flat, self._gflat = flat, gflat
        self._shadow = torch.empty(total, dtype=torch.bfloat16, device=device)
        self._shadow_version = -1
        self._slot_cache = slots

    def _refresh_shadow(self):
        # every in-place update of a parameter (optimizer step, load_state_dict, init) bumps its version counter ...
        v = sum(p._version for p, _ in self._slot_cache)
        # ... except writes through `.data`, which is how the reference's ema() updates the EMA model
        # (tools/trainer.py:12-18 `target_dict[key].data.copy_`).  That model only ever runs in eval mode, so an eval
        # forward always re-casts (one 6 B/param pass, ~2 % of a DiT-XL sampling forward); training keeps the check.
        if v != self._shadow_version or not self.training:
            if self._before_cast is not None:
                self._before_cast()
            L.call("vaw_cast_f32_bf16", self._flat.data_ptr(), self._shadow.data_ptr(), self._flat.numel(), L.stream_ptr())
            self._shadow_version = v

    def _ensure_workspace(self, batch, device):
        if self._ws is None or self._ws_batch != batch or self._ws.device != device:
            nbytes = self._workspace_bytes(batch)
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._ws_batch = batch

    def flat_parameters(self):
        """(flat fp32 params, flat fp32 grads, bf16 shadow) — used by the fused optimizer and the DP all-reduce."""
        return self._flat, self._gflat, self._shadow

    def _bind_grads(self):
        """Point every trainable parameter's .grad at its slice of the flat gradient buffer.  Returns True when the
        gradients were unset (zero_grad(set_to_none=True)): the next backward then OVERWRITES the buffer."""
        fresh = any(p.grad is None for p, _ in self._slot_cache if p.requires_grad)
        if fresh:
            for p, o in self._slot_cache:
                if p.requires_grad:
                    p.grad = self._gflat[o:o + p.numel()].view(p.shape)
        return fresh
