"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch) for the two natural
collectives of the training step:

  * gradient all-reduce (average) — the reference gets it from DDP's bucketed hooks (main.py:347, trainer.py:94-108).
    Here the engine writes gradients straight into one flat buffer, so the all-reduce runs on contiguous per-block
    slices of that buffer, launched on a side stream as soon as the engine's backward has recorded the block's event
    (blocks finish in reverse order), overlapping with the rest of backward.
  * all_gather of the per-sample (t, loss) pairs for LossAwareSampler.update_with_local_losses (resample.py:85-106):
    one collective of packed int32 pairs instead of three collectives + 2*W*B host syncs.

Everything in this file is backend-agnostic tensor plumbing (works with gloo on CPU tensors, which is how the
world_size=2 CPU tests exercise it); the kernels it feeds are CUDA-only.
"""
from __future__ import annotations

import os
from contextlib import contextmanager

import torch
import torch.distributed as dist


def dist_ready() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def _pack(local_ts, l32, B, Bpad):
    """[2, Bpad] int32: row 0 = timesteps (-1 in the padding), row 1 = the fp32 losses' bit patterns."""
    packed = torch.empty(2, Bpad, dtype=torch.int32, device=l32.device)
    if l32.is_cuda:
        from . import _lib as L   # one launch instead of ~6 indexing ops (include/vaw_b200.h: vaw_pack_tloss)
        t64 = local_ts.to(torch.int64).reshape(-1).contiguous()
        L.call("vaw_pack_tloss", t64.data_ptr(), l32.data_ptr(), packed[0].data_ptr(), packed[1].data_ptr(), B, Bpad,
               L.stream_ptr())
        return packed
    packed[0, :B] = local_ts.to(torch.int32).reshape(-1)      # gloo / CPU tensors (the world_size-2 CPU tests)
    packed[1, :B] = l32.view(torch.int32)
    if Bpad > B:
        packed[0, B:] = -1
        packed[1, B:] = 0
    return packed


def gather_tloss(local_ts: torch.Tensor, local_losses: torch.Tensor, ragged: bool = True):
    """All-gather (timestep, loss) pairs in rank order.  Returns (ts int32 [sum B], losses fp32 [sum B]); padding
    entries (ragged batches only) carry t = -1 and are skipped by the update kernel.
    Bit-exactness: the fp32 loss travels as its int32 bit pattern, so no value is rounded on the way.
    ragged=True exchanges the batch sizes first and pads to the largest, like the reference (resample.py:85-100);
    ragged=False skips that exchange (and its host sync) when every rank is known to pass the same batch size."""
    l32 = local_losses.detach().to(torch.float32).reshape(-1).contiguous()
    if not dist_ready():
        return local_ts.to(torch.int32).reshape(-1).contiguous(), l32
    W = dist.get_world_size()
    B = l32.numel()
    Bpad = B
    if ragged:
        bmax = torch.tensor([B], dtype=torch.int32, device=l32.device)
        dist.all_reduce(bmax, op=dist.ReduceOp.MAX)
        Bpad = int(bmax.item())
    packed = _pack(local_ts, l32, B, Bpad)
    t32 = packed[0]
    flat = torch.empty(W * 2 * Bpad, dtype=torch.int32, device=t32.device)
    dist.all_gather_into_tensor(flat, packed.view(-1))
    gathered = flat.view(W, 2, Bpad)
    ts = gathered[:, 0, :].reshape(-1).contiguous()
    losses = gathered[:, 1, :].reshape(-1).contiguous().view(torch.float32)
    return ts, losses


class FlatGradSync:
    """Bucketed average all-reduce over a flat gradient buffer.

    buckets: list of buckets in the order their gradients become final during backward; a bucket is one (begin, end)
    element range or a list of ranges (a transformer block's tensors plus its slice of the stacked adaLN weights).
    With CUDA tensors the collectives run on a dedicated stream gated by per-bucket events recorded by the engine;
    call `launch(events)` right after backward was enqueued and `wait()` before the optimizer reads the gradients.
    """

    def __init__(self, gflat: torch.Tensor, buckets, process_group=None):
        self.gflat = gflat
        norm = []
        for bk in buckets:
            ranges = [bk] if isinstance(bk[0], (int, float)) else list(bk)
            norm.append([(int(b), int(e)) for b, e in ranges if e > b])
        self.buckets = norm
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist_ready() else 1
        self.stream = torch.cuda.Stream() if gflat.is_cuda else None
        self.enabled = True
        self._native_avg = dist_ready() and dist.get_backend(process_group) == "nccl"

    @contextmanager
    def no_sync(self):
        """Gradient accumulation window (the reference uses DDP.no_sync, trainer.py:94-101)."""
        old, self.enabled = self.enabled, False
        try:
            yield
        finally:
            self.enabled = old

    @contextmanager
    def _one_launch(self):
        """Collectives of ONE kind issued inside become a single NCCL group (one kernel launch): c10d's coalescing
        manager; plain back-to-back calls if this torch does not have it."""
        try:
            from torch.distributed.distributed_c10d import _coalescing_manager
        except ImportError:  # pragma: no cover
            yield
            return
        with _coalescing_manager(group=self.group):
            yield

    def _reduce(self, view):
        if self._native_avg:
            dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
            view.div_(self.world)

    def launch(self, events=None):
        if not self.enabled or self.world == 1:
            return
        if os.environ.get("VAW_DP_NOSYNC") == "1":   # measurement knob: data-parallel step without the gradient all-reduce
            return
        if self.stream is None:
            for ranges in self.buckets:
                for b, e in ranges:
                    self._reduce(self.gflat[b:e])
            return
        with torch.cuda.stream(self.stream):
            for i, ranges in enumerate(self.buckets):
                if events is not None and events[i] is not None:
                    self.stream.wait_event(events[i])
                else:
                    self.stream.wait_stream(torch.cuda.default_stream())
                if self._native_avg and len(ranges) > 1:
                    with self._one_launch():     # a bucket's ranges (weights, biases, adaLN slice) as ONE NCCL launch
                        for b, e in ranges:
                            dist.all_reduce(self.gflat[b:e], op=dist.ReduceOp.AVG, group=self.group)
                else:
                    for b, e in ranges:
                        self._reduce(self.gflat[b:e])

    def wait(self):
        if self.stream is not None and self.enabled and self.world > 1:
            torch.cuda.current_stream().wait_stream(self.stream)


class ShardedGradSync(FlatGradSync):
    """Data parallelism with the optimizer state sharded over ranks (ZeRO-1 style) on the flat buffers.

    The LARGE tensors of a bucket (a block's weight matrices) are reduce-scattered: rank r receives the average of its
    1/W slice only, updates that slice (fp32 master, moments, bf16 shadow) and the bf16 shadows are all-gathered before
    the next forward.  The SMALL tensors (biases - the forward reads them from the fp32 master buffer) and the tail
    (embedders, final layer, projectors) stay replicated: all-reduced and updated on every rank.  Against all-reduce +
    replicated AdamW this moves 0.75x the bytes over NVLink (2.7 GB reduce-scatter + 1.35 GB bf16 all-gather instead of a
    2 x 2.7 GB all-reduce), takes the second half of it out of the backward pass, and divides the AdamW pass by W.
    The fp32 master copy of a large tensor is complete only on its owners until `gather_master()`.

    buckets: in completion order, each a list of (kind, begin, end) with kind 'rs' or 'ar'."""

    def __init__(self, gflat, buckets, process_group=None):
        if not (dist_ready() and dist.get_backend(process_group) == "nccl" and gflat.is_cuda):
            raise RuntimeError("the sharded optimizer mode needs CUDA tensors and the NCCL backend")
        self.kinds = [[(k, int(b), int(e)) for k, b, e in bk if e > b] for bk in buckets]
        super().__init__(gflat, [[(b, e) for _, b, e in bk] for bk in self.kinds], process_group)
        self.rank = dist.get_rank(process_group)
        W = self.world
        for bk in self.kinds:
            for k, b, e in bk:
                if k == "rs" and ((e - b) % (4 * W) or b % 4):
                    raise RuntimeError(f"reduce-scatter range [{b}, {e}) does not split into {W} float4-aligned slices")
        self.gather_stream = torch.cuda.Stream()

    def _slice(self, b, e):
        n = (e - b) // self.world
        return b + self.rank * n, b + (self.rank + 1) * n

    def owned_ranges(self):
        """Element ranges this rank updates: its slice of every reduce-scattered range, and every all-reduced range."""
        owned, replicated = [], []
        for bk in self.kinds:
            for k, b, e in bk:
                (owned if k == "rs" else replicated).append(self._slice(b, e) if k == "rs" else (b, e))
        return owned, replicated

    def launch(self, events=None):
        """Per bucket (gated by its event): the reduce-scatters of its large tensors as one launch.  The small
        replicated tensors of ALL buckets (biases; a few MB in total, each all-reduce latency-bound) are averaged by one
        grouped all-reduce after the last bucket instead of two ~100 us collectives per block on the critical tail."""
        if not self.enabled or self.world == 1:
            return
        with torch.cuda.stream(self.stream):
            deferred = []
            for i, bk in enumerate(self.kinds):
                if events is not None and events[i] is not None:
                    self.stream.wait_event(events[i])
                else:
                    self.stream.wait_stream(torch.cuda.default_stream())
                rs = [(b, e) for k, b, e in bk if k == "rs"]
                deferred += [(b, e) for k, b, e in bk if k == "ar"]
                if rs:
                    with self._one_launch():
                        for b, e in rs:
                            lo, hi = self._slice(b, e)
                            dist.reduce_scatter_tensor(self.gflat[lo:hi], self.gflat[b:e], op=dist.ReduceOp.AVG,
                                                       group=self.group)
            if deferred:
                with self._one_launch():
                    for b, e in deferred:
                        dist.all_reduce(self.gflat[b:e], op=dist.ReduceOp.AVG, group=self.group)

    def _gather(self, buf, ranges):
        with self._one_launch():
            for b, e in ranges:
                lo, hi = self._slice(b, e)
                dist.all_gather_into_tensor(buf[b:e], buf[lo:hi], group=self.group)

    def gather_first(self, buf, first):
        """Start completing `buf` with the reduce-scattered ranges for which first(b, e) holds (one launch), after
        everything enqueued so far on the current stream; returns the event that marks their completion.  Lets the
        optimizer update and publish the tensors the forward reads first while it is still updating the rest."""
        self.gather_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.gather_stream):
            self._gather(buf, [(b, e) for bk in reversed(self.kinds) for k, b, e in bk if k == "rs" and first(b, e)])
            ev = torch.cuda.Event()
            ev.record(self.gather_stream)
        return ev

    def all_gather(self, buf, wait=True, first=None, first_done=None):
        """Complete `buf` (the bf16 shadow, or the fp32 master) from the owners' slices, in place, on a side stream that
        starts after everything enqueued so far (the optimizer step).  Buckets are gathered in FORWARD order (the reverse
        of their completion order in backward); ranges for which `first(b, e)` is true go before everything else (or
        were already started by gather_first(), whose event is passed as `first_done`).
        wait=False returns one CUDA event per step of that schedule instead of blocking the current stream:
        [after the `first` ranges, after bucket L-1 (= block 0), after bucket L-2, ...]."""
        cur = torch.cuda.current_stream()
        self.gather_stream.wait_stream(cur)
        events = []
        order = list(reversed(self.kinds))
        with torch.cuda.stream(self.gather_stream):
            if wait:      # nobody consumes it piecewise: one launch for everything
                self._gather(buf, [(b, e) for bk in order for k, b, e in bk if k == "rs"])
            else:
                if first_done is not None:
                    events.append(first_done)
                elif first is not None:
                    self._gather(buf, [(b, e) for bk in order for k, b, e in bk if k == "rs" and first(b, e)])
                    ev = torch.cuda.Event()
                    ev.record(self.gather_stream)
                    events.append(ev)
                for bk in order:
                    rest = [(b, e) for k, b, e in bk if k == "rs" and not (first is not None and first(b, e))]
                    if rest:
                        self._gather(buf, rest)
                        ev = torch.cuda.Event()
                        ev.record(self.gather_stream)
                        events.append(ev)
        if wait:
            cur.wait_stream(self.gather_stream)
            return None
        return events


def shard_seed(base_seed: int, rank: int) -> int:
    """Per-rank RNG seed rule of the reference (tools/utils.py:62-69): seed + rank."""
    return int(base_seed) + int(rank)
