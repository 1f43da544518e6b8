"""CUDA-graph capture of the launch-heavy part of a training step (K1 -> denoiser forward -> K2 -> backward).

A DiT-S step is ~390 library launches of 10-20 us; replaying them as one graph removes the per-launch host cost
(B = 64: 7.15 -> 6.31 ms per step on a B200, scripts/dev_graph_dits.py).  The sampler draw (host RNG + a small H2D copy)
and the optimizer (host-side step count; two launches) stay outside the graph.  Single-process only: the data-parallel
gradient all-reduce is issued from Python per bucket and is not captured.

Two things a replay cannot see and that are therefore handled around it: (1) the bf16 weight shadow - the version check
that triggers the fp32 -> bf16 cast runs in Python, so `__call__` refreshes the shadow before every replay (a no-op after
FusedAdamW, which writes the shadow itself; one cast pass after torch.optim.AdamW / load_state_dict / an EMA swap);
(2) overwrite-vs-accumulate is a launch argument of the backward and the graph is captured in overwrite mode, so
`accumulate=True` (gradient accumulation over micro-batches) saves the flat gradient buffer before the replay and adds
it back afterwards (vaw_add_f32: one extra pass over the gradients, small next to the step for the models that need a
graph at all).
"""
from __future__ import annotations

import torch

from . import _lib as L


class GraphedTrainingLosses:
    """`terms = graphed(x0, t, w, y=labels[, features][, noise])` runs

        terms = diffusion.training_losses(model, x0, features, t=t, model_kwargs={"y": y}, noise=noise)
        (terms["loss"] * w).mean().backward()

    as one graph replay.  The returned tensors are static buffers (overwritten by the next call); the gradients land in
    the model's flat gradient buffer exactly as after an eager backward (`.grad` views are re-bound).
    """

    def __init__(self, diffusion, model, x0_shape, *, class_cond=True, feature_shape=None, warmup=3, device=None):
        dev = torch.device(device) if device is not None else next(model.parameters()).device
        if dev.type != "cuda":
            raise L.VawError("CUDA graphs need a CUDA model (no CPU fallback)")
        if getattr(model, "_post_backward", None) is not None:
            raise L.VawError("GraphedTrainingLosses: the data-parallel gradient hook cannot be captured; "
                             "wrap the bare module (single process) instead")
        self.diffusion, self.model = diffusion, model
        B = x0_shape[0]
        self.x0 = torch.zeros(x0_shape, device=dev)
        self.noise = torch.zeros(x0_shape, device=dev)
        self.t = torch.zeros(B, dtype=torch.int64, device=dev)
        self.w = torch.ones(B, device=dev)
        self.y = torch.zeros(B, dtype=torch.int64, device=dev) if class_cond else None
        self.features = torch.zeros(feature_shape, device=dev) if feature_shape is not None else None
        self._terms = None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self.noise.normal_()
            for _ in range(max(1, warmup)):     # allocator warm-up, lazy module loads, weight shadows
                self._run()
                self._clear_grads()
        torch.cuda.current_stream(dev).wait_stream(side)
        self._clear_grads()                 # captured in overwrite mode: a replay REPLACES the gradients
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.launch_count()
        with torch.cuda.graph(self.graph):
            self._terms = self._run()
        self.launches_per_replay = L.launch_count() - n0   # library kernels inside one replay (accounting, bench.py)
        self._terms = {k: v.detach() for k, v in self._terms.items()}
        self._saved_grads = None

    def _clear_grads(self):
        for p in self.model.parameters():
            p.grad = None

    def _run(self):
        kw = {"y": self.y} if self.y is not None else {}
        terms = self.diffusion.training_losses(self.model, self.x0, self.features, t=self.t, model_kwargs=kw,
                                               noise=self.noise)
        (terms["loss"] * self.w).mean().backward()
        return terms

    @torch.no_grad()
    def __call__(self, x0, t, w=None, y=None, features=None, noise=None, accumulate=False):
        self.x0.copy_(x0)
        self.t.copy_(t)
        if w is None:
            self.w.fill_(1.0)
        else:
            self.w.copy_(w)
        if self.y is not None:
            if y is None:
                raise ValueError("class-conditional capture needs labels y")
            self.y.copy_(y)
        if self.features is not None:
            if features is None:
                raise ValueError("captured with an alignment loss: features are required")
            self.features.copy_(features)
        if noise is None:
            self.noise.normal_()
        else:
            self.noise.copy_(noise)
        refresh = getattr(self.model, "_refresh_shadow", None)
        if refresh is not None:
            refresh()                  # weights changed outside FusedAdamW since the last replay -> re-cast the shadow
        gflat = None
        if accumulate:
            flat_parameters = getattr(self.model, "flat_parameters", None)
            if flat_parameters is None:
                raise L.VawError("gradient accumulation across replays needs an engine-backed model")
            gflat = flat_parameters()[1]
            if self._saved_grads is None or self._saved_grads.shape != gflat.shape:
                self._saved_grads = torch.empty_like(gflat)
            self._saved_grads.copy_(gflat)
        self.graph.replay()
        terms = self._terms
        if gflat is not None:
            L.call("vaw_add_f32", self._saved_grads.data_ptr(), gflat.data_ptr(), gflat.numel(), L.stream_ptr())
        bind = getattr(self.model, "_bind_grads", None)
        if bind is not None:
            bind()                     # `.grad` views of the flat gradient buffer, as after an eager backward
        return terms
