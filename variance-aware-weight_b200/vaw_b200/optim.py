"""Fused optimizer step over the engine's flat parameter buffer, and the data-parallel model wrapper.

FusedAdamW replaces, for models that expose `flat_parameters()` (vaw_b200.models.DiT / UViT), the reference's
`optim.AdamW(model.parameters(), lr, betas, weight_decay, eps)` (main.py:354) + GradScaler unscale / inf check
(trainer.py:124-129) + rank-0 EMA (trainer.py:12-18) with ONE HBM-bound pass (vaw_adamw_step_amp): read p, g, m, v;
write p, m, v, the bf16 shadow used by the tensor cores, and optionally the EMA copy.

It IS a `torch.optim.Optimizer`: `param_groups[i]["params"]` hold the model's real Parameters, `state[p]` carries
`step / exp_avg / exp_avg_sq` (views of the flat moment buffers, torch.optim.AdamW's key names, so `state_dict()`
round-trips with the reference's checkpoints, tools/utils.py:93-120), which is what the reference's
`LambdaLR(optimizer, ...)` (main.py:355), `scaler.unscale_(optimizer)` / `scaler.step(optimizer)` (trainer.py:124-129)
and `optimizer.zero_grad()` (:133) need.  torch.optim.AdamW over `model.parameters()` keeps working too (the
per-tensor Parameters are views of the flat buffer); this class is the fast path.
"""
from __future__ import annotations

import ctypes as C
from contextlib import nullcontext

import torch

from . import _lib as L
from .parallel import FlatGradSync, ShardedGradSync, dist_ready

_PTR, _LL, _D = C.c_void_p, C.c_longlong, C.c_double
L.register("vaw_adamw_step", [_PTR] * 6 + [_LL] + [_D] * 5 + [_LL, _D, _D, _PTR, _PTR])
L.register("vaw_adamw_step_amp", [_PTR] * 6 + [_LL] + [_D] * 5 + [_LL, _D, _D, _PTR, _PTR, _PTR, _PTR])
L.register("vaw_adamw_step_ranges", [_PTR] * 7 + [C.c_int, _LL] + [_D] * 5 + [_LL, _D, _D, _PTR, _PTR, _PTR, _PTR])
L.register("vaw_grad_clip_coef", [_PTR, _LL, _D, _D, _PTR, _PTR, _PTR])


def _intersect(ranges_a, ranges_b):
    """Intersection of two lists of half-open element ranges."""
    out = []
    for a0, a1 in ranges_a:
        for b0, b1 in ranges_b:
            lo, hi = max(a0, b0), min(a1, b1)
            if hi > lo:
                out.append((lo, hi))
    return sorted(out)


def _unwrap(model):
    return model.module if isinstance(model, DataParallel) else model


class FusedAdamW(torch.optim.Optimizer):
    """FusedAdamW(model, lr=..., betas=..., eps=..., weight_decay=..., ema_decay=None)

    `model` is an engine-backed module (or its DataParallel wrapper).  `params` may restrict / group the parameters the
    way torch optimizers allow (an iterable of Parameters or of param-group dicts); every one of them must belong to
    `model`.  Frozen parameters (DiT's pos_embed) are never updated."""

    # torch.cuda.amp.GradScaler.step() hands such optimizers `self.grad_scale` / `self.found_inf` (device tensors) and
    # calls step() unconditionally: the kernel unscales and skips on the device, no `.item()` sync (trainer.py:128)
    _step_supports_amp_scaling = True

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, ema_decay=None, params=None):
        if not isinstance(model, torch.nn.Module):
            raise TypeError("FusedAdamW(model, ...): pass the engine-backed module (its parameters are found through it)")
        self.model = model
        self.ema_decay = ema_decay
        self.step_count = 0
        self.m = self.v = self.ema = None
        self.grad_norm = self._norm_ws = None
        self._group_ranges = None
        self._flat_ptr = None
        self._loaded = None
        self._shard_ranges = None
        if params is None:
            params = [p for p in model.parameters() if p.requires_grad]
        # the param-group keys of torch.optim.AdamW, so that a state_dict saved here loads into torch's AdamW with the
        # same meaning (without `decoupled_weight_decay` torch's Adam.__setstate__ falls back to L2 decay) and back
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False,
                                      maximize=False, foreach=None, capturable=False, differentiable=False,
                                      fused=None, decoupled_weight_decay=True))

    # the single-group shorthands the round-1 API exposed
    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    # -------------------------------------------------------------------------------------------------
    def _ensure_state(self):
        flat, gflat, shadow = self.model.flat_parameters()
        if flat is None:
            raise L.VawError("FusedAdamW: run a forward pass (or model._ensure_flat(device)) before the first step")
        stale = (self.m is None or self.m.device != flat.device or self.m.numel() != flat.numel()
                 or self._flat_ptr != flat.data_ptr())
        if stale:
            old = (self.m, self.v) if self.m is not None and self.m.numel() == flat.numel() else None
            self.m = torch.zeros_like(flat, requires_grad=False)
            self.v = torch.zeros_like(flat, requires_grad=False)
            if old is not None:        # the model repacked its buffer (.to(), load_state_dict(assign=True)): keep the moments
                self.m.copy_(old[0])
                self.v.copy_(old[1])
            if self.ema_decay is not None and (self.ema is None or self.ema.numel() != flat.numel()
                                               or self.ema.device != flat.device):
                self.ema = flat.detach().clone()
            self._flat_ptr = flat.data_ptr()
            self._build_ranges(flat)
            if self._loaded is not None:   # moments restored by load_state_dict before the buffers existed
                for p, o in _unwrap(self.model)._slot_cache:
                    st = self._loaded.get(id(p))
                    if st is not None and "exp_avg" in st:
                        self.m[o:o + p.numel()].copy_(st["exp_avg"].reshape(-1))
                        self.v[o:o + p.numel()].copy_(st["exp_avg_sq"].reshape(-1))
                self._loaded = None
        return flat, gflat, shadow

    def _build_ranges(self, flat):
        """Per param group: contiguous element ranges of its TRAINABLE tensors in the flat buffer (the alignment
        padding between two neighbours of the same group rides along), and the per-parameter state views."""
        slots = sorted(((o, p) for p, o in _unwrap(self.model)._slot_cache), key=lambda x: x[0])
        owner = {}
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                owner[id(p)] = gi
        known = {id(p) for _, p in slots}
        missing = [1 for g in self.param_groups for p in g["params"] if id(p) not in known]
        if missing:
            raise L.VawError(f"FusedAdamW: {len(missing)} parameter(s) in param_groups do not belong to the model")
        self._group_ranges = [[] for _ in self.param_groups]
        self._step_t = step_t = torch.tensor(float(self.step_count))   # one tensor shared by every state entry
        for i, (o, p) in enumerate(slots):
            gi = owner.get(id(p))
            if gi is None or not p.requires_grad:
                continue
            # up to the next tensor's (64-element aligned) offset: the zero padding rides along and stays zero
            end = slots[i + 1][0] if i + 1 < len(slots) else flat.numel()
            r = self._group_ranges[gi]
            if r and r[-1][1] == o:
                r[-1][1] = end
            else:
                r.append([o, end])
            self.state[p] = {"step": step_t, "exp_avg": self.m[o:o + p.numel()].view(p.shape),
                             "exp_avg_sq": self.v[o:o + p.numel()].view(p.shape)}

    # -------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, *, grad_scale: float = 1.0, max_grad_norm: float | None = None):
        """One AdamW step.  `grad_scale` multiplies the gradients (an accumulation average); `max_grad_norm` applies
        torch.nn.utils.clip_grad_norm_ semantics (trainer.py:60-62) without a host sync: the norm and the clip
        coefficient stay on the device (`self.grad_norm` holds [norm, coef] after the call).  Under
        `GradScaler.step(optimizer)` the scaler's `grad_scale` / `found_inf` device tensors are honoured in-kernel."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        flat, gflat, shadow = self._ensure_state()
        amp_scale = getattr(self, "grad_scale", None)      # set by GradScaler.step; None once unscale_() already ran
        found_inf = getattr(self, "found_inf", None)
        inv_scale = None
        if isinstance(amp_scale, torch.Tensor):
            inv_scale = amp_scale.to(device=flat.device, dtype=torch.float32).reciprocal().reshape(1)
        if isinstance(found_inf, torch.Tensor):
            found_inf = found_inf.to(device=flat.device, dtype=torch.float32).reshape(1)
        else:
            found_inf = None
        shard = getattr(self.model, "_shard_sync", None)
        if (found_inf is not None or inv_scale is not None) and shard is not None:
            # sharded mode: a rank sees the averaged gradient only on the slices it owns (elsewhere its own local values),
            # so GradScaler's per-rank overflow check - and with it the loss scale - could differ between ranks
            raise L.VawError("FusedAdamW: an enabled GradScaler is not supported with DataParallel(shard_optimizer=True) "
                             "(bf16 autocast needs no loss scaling: GradScaler(enabled=False))")
        self.step_count += 1   # host-side bias-correction counter; a skipped (inf) step is undone below
        if found_inf is not None:
            # torch's own fused AdamW does `step -= found_inf` on the device; the counter here lives on the host because
            # the bias corrections are kernel ARGUMENTS - one 4-byte read, only on the GradScaler path (which the
            # reference syncs on anyway, trainer.py:128 -> _maybe_opt_step's .item())
            if float(found_inf.item()) != 0.0:
                self.step_count -= 1
                return loss
        clip = None
        if max_grad_norm and shard is not None:
            raise L.VawError("FusedAdamW: gradient clipping needs the full gradient on every rank; it is not "
                             "available with DataParallel(shard_optimizer=True)")
        if max_grad_norm:
            if self.grad_norm is None or self.grad_norm.device != flat.device:
                self.grad_norm = torch.zeros(2, device=flat.device)
                self._norm_ws = torch.empty(1024, device=flat.device)
            L.call("vaw_grad_clip_coef", gflat.data_ptr(), gflat.numel(), float(grad_scale), float(max_grad_norm),
                   self._norm_ws.data_ptr(), self.grad_norm.data_ptr(), L.stream_ptr())
            clip = self.grad_norm.data_ptr() + 4
        ema_decay = float(self.ema_decay if self.ema_decay is not None else 0.0)
        if shard is not None:
            self._step_sharded(shard, flat, gflat, shadow, grad_scale, ema_decay, inv_scale, found_inf)
            return loss
        for g, ranges in zip(self.param_groups, self._group_ranges):
            if g.get("amsgrad") or g.get("maximize") or not g.get("decoupled_weight_decay", True):
                raise L.VawError("FusedAdamW implements torch.optim.AdamW's default update only "
                                 "(amsgrad=False, maximize=False, decoupled weight decay)")
            b1, b2 = g["betas"]
            for off, end in ranges:
                n = end - off
                L.call("vaw_adamw_step_amp", flat.data_ptr() + 4 * off, gflat.data_ptr() + 4 * off,
                       self.m.data_ptr() + 4 * off, self.v.data_ptr() + 4 * off, shadow.data_ptr() + 2 * off,
                       self.ema.data_ptr() + 4 * off if self.ema is not None else None, n, float(g["lr"]),
                       float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]), self.step_count,
                       float(grad_scale), ema_decay, clip, L.ptr(inv_scale), L.ptr(found_inf), L.stream_ptr())
        self._step_t.fill_(float(self.step_count))
        # the kernel refreshed the bf16 shadow itself: mark it current so the next forward skips the cast pass
        m = _unwrap(self.model)
        m._shadow_version = sum(p._version for p, _ in m._slot_cache)
        return loss

    def _step_sharded(self, shard, flat, gflat, shadow, grad_scale, ema_decay, inv_scale, found_inf):
        """DataParallel(shard_optimizer=True): update this rank's slice of every reduce-scattered range and the
        replicated ranges in ONE launch per param group, then complete the bf16 shadows from their owners."""
        m = _unwrap(self.model)
        lo, hi = m.stacked_range() if hasattr(m, "stacked_range") else (0, 0)
        if self._shard_ranges is None or self._shard_ranges[0] is not shard:
            owned, repl = shard.owned_ranges()
            mine = sorted(owned + repl)
            per_group = []
            for ranges in self._group_ranges:
                sel = _intersect([(a, b) for a, b in ranges], mine)
                # the stacked adaLN weights are what the next forward reads first: their slices are updated by a launch
                # of their own so that their all-gather runs under the update of everything else
                parts = []
                for part in ([r for r in sel if lo <= r[0] and r[1] <= hi], [r for r in sel if not (lo <= r[0] and r[1] <= hi)]):
                    t = torch.tensor([[a, b - a] for a, b in part], dtype=torch.int64, device=flat.device).reshape(-1, 2)
                    parts.append((t, len(part), max([b - a for a, b in part], default=0)))
                per_group.append(parts)
            self._shard_ranges = (shard, per_group)
        first_done = None
        for phase in (0, 1):
            for g, parts in zip(self.param_groups, self._shard_ranges[1]):
                if g.get("amsgrad") or g.get("maximize") or not g.get("decoupled_weight_decay", True):
                    raise L.VawError("FusedAdamW implements torch.optim.AdamW's default update only")
                t, n, mx = parts[phase]
                if n == 0:
                    continue
                b1, b2 = g["betas"]
                L.call("vaw_adamw_step_ranges", flat.data_ptr(), gflat.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                       shadow.data_ptr(), self.ema.data_ptr() if self.ema is not None else None, t.data_ptr(), n, mx,
                       float(g["lr"]), float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]), self.step_count,
                       float(grad_scale), ema_decay, None, L.ptr(inv_scale), L.ptr(found_inf), L.stream_ptr())
            if phase == 0 and hi > lo:
                first_done = shard.gather_first(shadow, lambda b, e: lo <= b and e <= hi)
        self._step_t.fill_(float(self.step_count))
        self.model._after_sharded_step(shadow, first_done)
        m._shadow_version = sum(p._version for p, _ in m._slot_cache)

    # -------------------------------------------------------------------------------------------------
    def state_dict(self):
        if self.m is None and self.model.flat_parameters()[0] is not None:
            self._ensure_state()
        sd = super().state_dict()
        # inside this class every entry shares ONE step tensor; a consumer such as torch.optim.AdamW increments the step
        # of every parameter separately (`_foreach_add_` over the list), so each entry leaves with its own copy
        sd["state"] = {k: {kk: (vv.clone() if kk == "step" else vv) for kk, vv in st.items()}
                       for k, st in sd["state"].items()}
        return sd

    def load_state_dict(self, state_dict):
        """Accepts what torch.optim.AdamW.state_dict() / this class's state_dict() produce (tools/utils.py:109-120)."""
        super().load_state_dict(state_dict)
        loaded = {id(p): dict(st) for p, st in self.state.items()}
        self.state.clear()
        steps = [int(float(st["step"])) for st in loaded.values() if "step" in st]
        self.step_count = max(steps) if steps else 0
        self.m = None
        self._loaded = loaded
        if self.model.flat_parameters()[0] is not None:
            self._ensure_state()

    def ema_state_dict(self):
        """The EMA weights (trainer.py:12-18 keeps them in a second model) under the model's parameter names, as views
        of the flat EMA buffer; load them into a model copy with load_state_dict(..., strict=False)."""
        if self.ema is None:
            raise L.VawError("FusedAdamW was built without ema_decay")
        shard = getattr(self.model, "_shard_sync", None)
        if shard is not None:
            # sharded mode keeps the EMA of a reduce-scattered range on its owner only: complete it (a collective)
            shard.all_gather(self.ema, wait=True)
        m = _unwrap(self.model)
        by_id = {id(p): o for p, o in m._slot_cache}
        return {k: self.ema[by_id[id(p)]:by_id[id(p)] + p.numel()].view(p.shape)
                for k, p in m.named_parameters() if id(p) in by_id}

    def zero_grad(self, set_to_none: bool = True):
        """`set_to_none=True` (torch's default, what trainer.py:133 gets) makes the next backward OVERWRITE the flat
        gradient buffer - no memset pass; False zeroes the buffer in place."""
        if set_to_none:
            for p in _unwrap(self.model).parameters():
                p.grad = None
        else:
            super().zero_grad(set_to_none=False)


class DataParallel(torch.nn.Module):
    """Drop-in for the reference's DDP wrapper (main.py:347): `.module`, `.no_sync()`, same forward.  Gradients are
    averaged over ranks by FlatGradSync on the flat gradient buffer, bucketed per transformer block and overlapped
    with backward through the engine's per-block events (DiT and U-ViT); a module without per-block ranges gets one
    whole-buffer bucket after backward."""

    def __init__(self, module, process_group=None, device_ids=None, output_device=None, shard_optimizer=False,
                 **_ddp_kwargs):
        """shard_optimizer=True (NCCL, models with `block_shard_ranges`, FusedAdamW): reduce-scatter the large tensors'
        gradients, update 1/W of them per rank, all-gather the bf16 shadows (parallel.ShardedGradSync).  The fp32
        `.data` of a large parameter is then complete on its owning ranks only until `gather_master()` - which
        `state_dict()` and every eval forward call for you; gradients of the large tensors are per-rank slices, so code
        that reads `.grad` itself (clip_grad_norm_, another optimizer) needs the default mode."""
        super().__init__()
        self.module = module
        self._sync = None
        self._shard = bool(shard_optimizer)
        self._shard_sync = None
        self._master_stale = False
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
        if self._shard and hasattr(m, "block_shard_ranges"):
            big, small, tail, total = m.block_shard_ranges()
            buckets = [[("rs", b, e) for b, e in big[i]] + [("ar", b, e) for b, e in small[i]]
                       for i in reversed(range(len(big)))] + [[("ar", b, e) for b, e in tail]]
            n_blocks = len(big)
            ev = [torch.cuda.Event() for _ in range(n_blocks + 1)]
            for e in ev:
                e.record()
            m._events = ev
            self._events = list(reversed(ev[:n_blocks])) + [ev[n_blocks]]
            self._sync = self._shard_sync = ShardedGradSync(m._gflat, buckets, self._group)
            m._post_backward = self._after_backward
            m._before_cast = self.gather_master
            return
        if hasattr(m, "block_grad_ranges"):
            per_block, tail, total = m.block_grad_ranges()
            # gradients become final block L-1 first ... block 0, then the embedders / final layer / projectors
            buckets = list(reversed(per_block)) + [tail]
            n_blocks = len(per_block)
            ev = [torch.cuda.Event() for _ in range(n_blocks + 1)]
            for e in ev:
                e.record()  # materialise the CUDA events so their handles can be passed through the C ABI
            m._events = ev
            self._events = list(reversed(ev[:n_blocks])) + [ev[n_blocks]]
        else:
            buckets = [(0, m._gflat.numel())]
            self._events = None
        self._sync = FlatGradSync(m._gflat, buckets, self._group)
        m._post_backward = self._after_backward

    def _after_backward(self):
        # backward is fully enqueued at this point: the bucket all-reduces (gated by the per-block events) overlap
        # with it on the side stream, and whatever the caller enqueues next (optimizer) is ordered after them.
        if self._sync.gflat.data_ptr() != self.module._gflat.data_ptr():
            # the module repacked its flat buffers (.to(), load_state_dict(assign=True)) after the wrapper was built
            self._sync.gflat = self.module._gflat
        self._sync.launch(self._events)
        self._sync.wait()

    def no_sync(self):
        if self._sync is None:
            return nullcontext()
        return self._sync.no_sync()

    # ---- sharded-optimizer mode ----------------------------------------------------------------------------
    def _after_sharded_step(self, shadow, first_done=None):
        """Called by FusedAdamW after it updated this rank's slices: complete the bf16 shadows (what the next forward's
        GEMMs read) from their owners; the fp32 master copies are completed lazily."""
        m = self.module
        lo, hi = m.stacked_range() if hasattr(m, "stacked_range") else (0, 0)
        # the stacked adaLN weights are read by ONE GEMM at the top of the forward: gather them first; then block 0, 1, ...
        # each with its own event, which the engine's forward waits on right before the block (vaw_dit_forward_ev)
        events = self._shard_sync.all_gather(shadow, wait=False, first=lambda b, e: lo <= b and e <= hi,
                                             first_done=first_done)
        m._fwd_wait = events if len(events) == m.depth + 1 else None
        if m._fwd_wait is None:
            torch.cuda.current_stream().wait_stream(self._shard_sync.gather_stream)
        self._master_stale = True

    def gather_master(self):
        """Complete the fp32 master parameters of the sharded tensors on every rank (collective: call on all ranks)."""
        if self._shard_sync is not None and self._master_stale:
            self._shard_sync.all_gather(self.module._flat.data, wait=True)
            self._master_stale = False

    def state_dict(self, *a, **k):
        self.gather_master()
        return super().state_dict(*a, **k)

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
