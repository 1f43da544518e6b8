"""Fused optimizer step over the engine's flat parameter buffer, and the data-parallel model wrapper.

FusedAdamW replaces, for models that expose `flat_parameters()` (vaw_b200.models.DiT), the reference's
`optim.AdamW(model.parameters(), lr, betas, weight_decay, eps)` (main.py:354) + GradScaler unscale (trainer.py:124-129)
+ rank-0 EMA (trainer.py:12-18) with ONE HBM-bound pass (vaw_adamw_step): read p, g, m, v; write p, m, v, the bf16
shadow used by the tensor cores, and optionally the EMA copy.  torch.optim.AdamW over `model.parameters()` keeps
working (the per-tensor Parameters are views of the flat buffer); this class is the fast path.
"""
from __future__ import annotations

import ctypes as C
from contextlib import nullcontext

import torch

from . import _lib as L
from .parallel import FlatGradSync, dist_ready

L.register("vaw_adamw_step", [C.c_void_p] * 6 + [C.c_longlong] + [C.c_double] * 5 + [C.c_longlong, C.c_double, C.c_double,
                                                                                      C.c_void_p, C.c_void_p])
L.register("vaw_grad_clip_coef", [C.c_void_p, C.c_longlong, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p])


class FusedAdamW:
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, ema_decay=None):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.ema_decay = ema_decay
        self.step_count = 0
        self.m = self.v = self.ema = None
        self.grad_norm = self._norm_ws = None
        self._ranges = []
        self.param_groups = [dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)]  # LambdaLR-style access

    def _ensure_state(self):
        flat, gflat, shadow = self.model.flat_parameters()
        if flat is None:
            raise L.VawError("FusedAdamW: run a forward pass (or model._ensure_flat(device)) before the first step")
        if self.m is None or self.m.device != flat.device or self.m.numel() != flat.numel():
            self.m = torch.zeros_like(flat, requires_grad=False)
            self.v = torch.zeros_like(flat, requires_grad=False)
            if self.ema_decay is not None:
                self.ema = flat.detach().clone()
            # contiguous element ranges of the TRAINABLE tensors (frozen ones - DiT's pos_embed - get no update and no
            # weight decay, like torch.optim.AdamW skipping parameters without a gradient); padding rides along
            slots = sorted(((o, p.numel(), p.requires_grad) for p, o in self.model._slot_cache), key=lambda x: x[0])
            ranges = []
            for i, (o, n, train) in enumerate(slots):
                end = slots[i + 1][0] if i + 1 < len(slots) else flat.numel()
                if not train:
                    continue
                if ranges and ranges[-1][1] == o:
                    ranges[-1][1] = end
                else:
                    ranges.append([o, end])
            self._ranges = [(a, b - a) for a, b in ranges]
        return flat, gflat, shadow

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0, max_grad_norm: float | None = None):
        """One AdamW step.  `grad_scale` multiplies the gradients (GradScaler unscale / accumulation average);
        `max_grad_norm` applies torch.nn.utils.clip_grad_norm_ semantics (trainer.py:60-62) without a host sync: the
        norm and the clip coefficient stay on the device (`self.grad_norm` holds [norm, coef] after the call)."""
        flat, gflat, shadow = self._ensure_state()
        self.step_count += 1
        g = self.param_groups[0]
        clip = None
        if max_grad_norm:
            if self.grad_norm is None or self.grad_norm.device != flat.device:
                self.grad_norm = torch.zeros(2, device=flat.device)
                self._norm_ws = torch.empty(1024, device=flat.device)
            L.call("vaw_grad_clip_coef", gflat.data_ptr(), gflat.numel(), float(grad_scale), float(max_grad_norm),
                   self._norm_ws.data_ptr(), self.grad_norm.data_ptr(), L.stream_ptr())
            clip = self.grad_norm.data_ptr() + 4
        for off, n in self._ranges:
            L.call("vaw_adamw_step", flat.data_ptr() + 4 * off, gflat.data_ptr() + 4 * off, self.m.data_ptr() + 4 * off,
                   self.v.data_ptr() + 4 * off, shadow.data_ptr() + 2 * off,
                   self.ema.data_ptr() + 4 * off if self.ema is not None else None, n, float(g["lr"]),
                   float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]),
                   self.step_count, float(grad_scale), float(self.ema_decay if self.ema_decay is not None else 0.0),
                   clip, L.stream_ptr())
        # the kernel refreshed the bf16 shadow itself: mark it current so the next forward skips the cast pass
        self.model._shadow_version = sum(p._version for p, _ in self.model._slot_cache)

    def ema_state_dict(self):
        """The EMA weights (trainer.py:12-18 keeps them in a second model) under the model's parameter names, as views
        of the flat EMA buffer; load them into a model copy with load_state_dict(..., strict=False)."""
        if self.ema is None:
            raise L.VawError("FusedAdamW was built without ema_decay")
        by_id = {id(p): o for p, o in self.model._slot_cache}
        return {k: self.ema[by_id[id(p)]:by_id[id(p)] + p.numel()].view(p.shape)
                for k, p in self.model.named_parameters() if id(p) in by_id}

    def zero_grad(self, set_to_none: bool = True):
        for p in self.model.parameters():
            p.grad = None  # the next backward overwrites the flat gradient buffer (no memset pass needed)


class DataParallel(torch.nn.Module):
    """Drop-in for the reference's DDP wrapper (main.py:347): `.module`, `.no_sync()`, same forward.  Gradients are
    averaged over ranks by FlatGradSync on the flat gradient buffer, bucketed per transformer block and overlapped
    with backward through the engine's per-block events."""

    def __init__(self, module, process_group=None):
        super().__init__()
        self.module = module
        self._sync = None
        self._group = process_group
        self._events = None
        first = next(module.parameters())
        if dist_ready() and first.is_cuda:
            self._setup()

    def _setup(self):
        m = self.module
        m._ensure_flat(next(m.parameters()).device)
        # every rank starts from rank 0's weights, like DDP's constructor broadcast (main.py:347)
        import torch.distributed as dist
        with torch.no_grad():
            dist.broadcast(m._flat.data, 0, group=self._group)
        m._shadow_version = -1
        per_block, tail, total = m.block_grad_ranges()
        # gradients become final block L-1 first ... block 0, then the embedders / final layer / projectors
        buckets = list(reversed(per_block)) + [tail]
        self._sync = FlatGradSync(m._gflat, buckets, self._group)
        ev = [torch.cuda.Event() for _ in range(m.depth + 1)]
        for e in ev:
            e.record()  # materialise the CUDA events so their handles can be passed through the C ABI
        m._events = ev
        self._events = list(reversed(ev[: m.depth])) + [ev[m.depth]]
        m._post_backward = self._after_backward

    def _after_backward(self):
        # backward is fully enqueued at this point: the bucket all-reduces (gated by the per-block events) overlap
        # with it on the side stream, and whatever the caller enqueues next (optimizer) is ordered after them.
        self._sync.launch(self._events)
        self._sync.wait()

    def no_sync(self):
        if self._sync is None:
            return nullcontext()
        return self._sync.no_sync()

    def wait_grads(self):
        if self._sync is not None:
            self._sync.wait()

    def forward(self, *a, **k):
        out = self.module(*a, **k)
        if self._sync is None and dist_ready():
            self._setup()
        return out

    def flat_parameters(self):
        return self.module.flat_parameters()

    @property
    def _slot_cache(self):
        return self.module._slot_cache

    @property
    def _shadow_version(self):
        return self.module._shadow_version

    @_shadow_version.setter
    def _shadow_version(self, v):
        self.module._shadow_version = v
