#!/usr/bin/env python
"""bench.py — training imgs/s of the diffusion training step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--config 2|3|4|5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workloads (BASELINE.json `configs`, 0-based index = --config; all synthetic data, random-init weights of the reference's
init scheme, eps-prediction, cosine schedule, weight_type=lambda, bf16 tensor-core math with fp32 master weights, fused
AdamW lr=1e-4 betas=(0.9, 0.95)):
  2  DiT-S/2 on 32x32x4 latents, class-cond 1000, UniformSampler, batch 256 per GPU (single GPU: K1 -> forward -> K2 ->
     backward replayed as one CUDA graph)
  3  DiT-XL/2 on 32x32x4 latents, class-cond 1000, LossSecondMomentResampler with a warmed-up history, batch 64 per GPU
     — the DEFAULT: the configuration BASELINE.json's metric is quoted on (it fits one GPU)
  4  U-ViT-M/4 on 64x64x3 pixels, class-cond 1000, UniformSampler, batch 64 per GPU
  5  DiT-XL/2 + REPA projector (encoder_depth 8, 2048-2048-768) with the alignment loss (gamma 0.5, 'mse') against the
     features of a frozen random-init MoCo-v3 ViT-B/16 computed INSIDE the step from synthetic 256-px pixels, batch 64

One step = sampler.sample -> [teacher features] -> training_losses (K1, denoiser forward, K2) -> backward -> [grad
all-reduce] -> AdamW -> sampler.update_with_local_losses.  `value` times K steps with inputs resident in HBM; `e2e` times
the same steps through the public API with the batch coming from pinned host memory and the loss read back every step.

--impl reference: the CPU arm.  The reference itself (a git-ignored copy under baseline/_ref, made by
__graft_entry__.build() where /root/reference exists; it travels to the GPU box with the snapshot) is run through its own
classes — models.dit / models.uvit, tools.gaussian_diffusion.GaussianDiffusion.training_losses, tools.resample, torch's
AdamW — on all host threads on a bounded batch of the same workload (`cpu_baseline.kind` = "reference").  Its three
missing third-party imports come from oracle/ref_stubs.  Without baseline/_ref the oracle port (oracle/train_step.py,
kind "port") is timed instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200")]

# train GFLOP per image (BASELINE.md §3 / SURVEY §8d: 2 FLOP per MAC, training = 3x forward, attention 4 T^2 D per block)
CONFIGS = {
    2: dict(model="DiT-S", family="dit", batch=256, cpu_batch=16, sampler="uniform", img=32, chans=4, patch=2,
            gflop=36.32, D=384, tokens=256, graph=True,
            desc="DiT-S/2 diffusion training step on 32x32x4 latents (256px), class-cond 1000"),
    3: dict(model="DiT-XL", family="dit", batch=64, cpu_batch=4, sampler="loss-second-moment", img=32, chans=4, patch=2,
            gflop=711.7, D=1152, tokens=256,
            desc="DiT-XL/2 diffusion training step on 32x32x4 latents (256px), class-cond 1000"),
    4: dict(model="UViT-M", family="uvit", batch=64, cpu_batch=4, sampler="uniform", img=64, chans=3, patch=4,
            gflop=211.4, D=768, tokens=258,
            desc="U-ViT-M/4 diffusion training step on 64x64x3 pixels, class-cond 1000"),
    5: dict(model="DiT-XL", family="dit", batch=64, cpu_batch=4, sampler="uniform", img=32, chans=4, patch=2,
            gflop=724.2 + 46.4, D=1152, tokens=256, repa=True,
            desc="DiT-XL/2 + REPA projector (encoder_depth 8) training step on 32x32x4 latents with the alignment loss "
                 "(gamma 0.5, mse) against frozen random-init MoCo-v3 ViT-B/16 features of synthetic 256px pixels, "
                 "teacher forward inside the step"),
}
DIT_GFLOP = {"DiT-S": 36.32, "DiT-B": 138.0, "DiT-L": 484.2, "DiT-XL": 711.7}
DIT_DIM = {"DiT-S": 384, "DiT-B": 768, "DiT-L": 1024, "DiT-XL": 1152}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--model", default=None, choices=list(DIT_GFLOP), help="override the DiT size of config 2/3/5")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (weak scaling); default per config")
    ap.add_argument("--sampler", default=None, choices=["uniform", "loss-second-moment"])
    ap.add_argument("--cpu-batch", type=int, default=None, help="batch of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-leg", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="config 2: eager launches instead of the CUDA graph")
    ap.add_argument("--no-shard", action="store_true", help="N > 1: all-reduce + replicated AdamW instead of the sharded optimizer")
    a = ap.parse_args()
    c = dict(CONFIGS[a.config])
    if a.model and c["family"] == "dit":
        base = DIT_GFLOP[c["model"]]
        c.update(model=a.model, gflop=c["gflop"] - base + DIT_GFLOP[a.model], D=DIT_DIM[a.model],
                 desc=c["desc"].replace(c["model"], a.model))
    c["batch"] = a.batch or c["batch"]
    c["cpu_batch"] = a.cpu_batch or c["cpu_batch"]
    c["sampler"] = a.sampler or c["sampler"]
    a.cfg = c
    return a


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sustained=d.get("bf16_tflops_sustained", 1400.0), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def synthetic_history(seed=0, T=1000, H=10):
    """Deterministic warmed-up loss history (every timestep has H entries) so that the benchmark exercises the
    non-uniform branch of LossSecondMomentResampler.weights (resample.py:145-149).  Same generator as the test fixture."""
    import numpy as np
    rng = np.random.RandomState(1234 + seed)
    base = 0.02 + 0.5 * np.exp(-np.arange(T) / 300.0)
    hist = np.abs(base[:, None] * (1.0 + 0.1 * rng.randn(T, H))).astype(np.float32).astype(np.float64)
    return hist, np.full(T, H, dtype=int)


# ------------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the reference itself (baseline/_ref) through its own classes, else the oracle port
# ------------------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def _reference_available():
    return os.path.exists(os.path.join(REF_DIR, "tools", "gaussian_diffusion.py"))


class ReferenceCpuStep:
    """One training step of the UNMODIFIED reference on the host: models.dit.DiT_* / models.uvit.UViT_M,
    GaussianDiffusion.training_losses (tools/gaussian_diffusion.py:834), tools.resample samplers composed the upstream
    way, torch.optim.AdamW (main.py:354); config 5 adds encoders.mocov3_vit.vit_base + align_utils.get_feature."""

    def __init__(self, cfg, batch):
        import warnings
        from types import SimpleNamespace
        import numpy as np
        import torch
        warnings.filterwarnings("ignore")
        # the reference imports timm / diffusers / torchdiffeq at module import; none is installed: import stubs
        sys.path[:0] = [os.path.join(ROOT, "oracle", "ref_stubs"), REF_DIR]
        import models.dit as rdit
        import models.uvit as ruvit
        import tools.gaussian_diffusion as rgd
        import tools.resample as rrs
        self.torch, self.B, self.cfg = torch, batch, cfg
        repa = bool(cfg.get("repa"))
        args = SimpleNamespace(weight_type="lambda", gamma=0.5, learn_sigma=False, p2_gamma=1.0, p2_k=1.0,
                               time_dist=["uniform"], learn_align=repa, align_type="mse", amp=False,
                               enc_type="mocov3-vit-b")
        self.args = args
        self.diffusion = rgd.GaussianDiffusion(
            args=args, betas=rgd.get_named_beta_schedule("cosine", 1000), model_mean_type=rgd.ModelMeanType.EPSILON,
            model_var_type=rgd.ModelVarType.FIXED_LARGE, loss_type=rgd.LossType.MSE, rescale_timesteps=True, device="cpu")
        torch.manual_seed(42)
        np.random.seed(42)
        if cfg["family"] == "dit":
            kw = dict(learn_align=True, encoder_depth=8, z_dims=768, projector_dim=2048) if repa else {}
            self.model = rdit.DiT_models[cfg["model"]](image_size=32, patch_size=2, in_channels=4, class_dropout_prob=0.0,
                                                       num_classes=1000, learn_sigma=False, **kw)
        else:
            self.model = ruvit.UViT_M(image_size=64, patch_size=4, in_channels=3, num_classes=1000, class_dropout_prob=0.0)
        self.model.train()
        self.sampler = rrs.create_named_schedule_sampler(cfg["sampler"], self.diffusion)
        if cfg["sampler"] == "loss-second-moment":
            hist, counts = synthetic_history(0)
            self.sampler._loss_history[:] = hist
            self.sampler._loss_counts[:] = counts
        self.opt = torch.optim.AdamW(self.model.parameters(), lr=1e-4, betas=(0.9, 0.95), weight_decay=0.0, eps=1e-8)
        self.x = torch.randn(batch, cfg["chans"], cfg["img"], cfg["img"])
        if cfg["family"] == "uvit":
            self.x.clamp_(-1, 1)
        self.y = torch.randint(0, 1000, (batch,))
        self.teacher = None
        if repa:
            import encoders.mocov3_vit as rvit
            import tools.align_utils as rau
            self.teacher = rvit.vit_base(num_classes=0).eval()
            self.get_feature = rau.get_feature
            self.pixels = torch.randint(0, 256, (batch, 3, 256, 256)).float()

    def step(self):
        feats = self.get_feature(self.args, self.pixels, self.teacher) if self.teacher is not None else None
        t, w = self.sampler.sample(self.B, "cpu")
        terms = self.diffusion.training_losses(self.model, self.x, feats, t=t, model_kwargs={"y": self.y})
        if self.cfg["sampler"] == "loss-second-moment":   # single process: update_with_local_losses == all_losses
            self.sampler.update_with_all_losses(t.tolist(), terms["loss"].detach().tolist())
        loss = (terms["loss"] * w).mean()
        loss.backward()
        self.opt.step()
        self.opt.zero_grad()
        return float(loss)


class PortCpuStep:
    """Fallback when baseline/_ref is absent: the oracle's restatement of the step (DiT configs only)."""

    def __init__(self, cfg, batch):
        import numpy as np
        import torch
        from oracle.train_step import OracleTrainer
        if cfg["family"] != "dit" or cfg.get("repa"):
            raise RuntimeError("the oracle port covers the plain DiT step only; baseline/_ref is needed for this config")
        np.random.seed(42)
        torch.manual_seed(42)
        self.tr = OracleTrainer(cfg["model"], seed=0, sampler=cfg["sampler"])
        self.x = torch.randn(batch, 4, 32, 32)
        self.y = torch.randint(0, 1000, (batch,))

    def step(self):
        return float(self.tr.step(self.x, self.y)[0])


def cpu_step_rate(cfg, batch, steps, warmup):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    import contextlib
    kind = "reference" if _reference_available() else "port"
    # the reference prints at import time ("attention mode is ..."): stdout carries the JSON line and nothing else
    with contextlib.redirect_stdout(sys.stderr):
        runner = (ReferenceCpuStep if kind == "reference" else PortCpuStep)(cfg, batch)
        for _ in range(warmup):
            runner.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            runner.step()
        dt = (time.perf_counter() - t0) / max(1, steps)
    return batch / dt, dt, torch.get_num_threads(), kind


def run_reference(args, rank):
    if rank != 0:
        return
    cfg = args.cfg
    ips, dt, threads, kind = cpu_step_rate(cfg, cfg["cpu_batch"], args.steps, args.warmup)
    what = ("the reference's own classes (baseline/_ref: models, GaussianDiffusion.training_losses, resample, AdamW)"
            if kind == "reference" else "oracle port of the reference step")
    sample = (f"{cfg['desc']}, batch {cfg['cpu_batch']}, fp32, {what}: forward+backward+AdamW+sampler, "
              f"{args.warmup} warm-up + {args.steps} timed steps")
    line = {
        "impl": "reference", "metric": "training imgs/sec", "value": ips, "unit": "imgs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cpu=True),
        "cpu_baseline": {"value": ips, "unit": "imgs/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": ips, "unit": "imgs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, cpu=False):
    cfg = args.cfg
    b = cfg["cpu_batch"] if cpu else cfg["batch"]
    return {"workload": f"{cfg['desc']}, eps-pred, cosine schedule, weight_type=lambda, {cfg['sampler']} sampler",
            "baseline_config_index": args.config, "per_gpu_batch": b,
            "global_batch": b if cpu else b * args.gpus,
            "parallelism": f"dp{args.gpus}", "optimizer": "AdamW lr=1e-4 betas=(0.9,0.95)",
            "l2": "per-step working set (GBs of activations + 3x the parameters) exceeds the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------------------
def build_native(args, dev):
    """(model, diffusion, teacher or None) of the chosen config, from the package's public constructors."""
    from vaw_b200.models import dit as vdit
    from vaw_b200.models import uvit as vuvit
    from vaw_b200.tools import gaussian_diffusion as gd
    cfg = args.cfg
    repa = bool(cfg.get("repa"))
    if cfg["family"] == "dit":
        kw = dict(learn_align=True, encoder_depth=8, z_dims=768, projector_dim=2048) if repa else {}
        model = vdit.DiT_models[cfg["model"]](image_size=32, patch_size=2, in_channels=4, class_dropout_prob=0.0,
                                              num_classes=1000, learn_sigma=False, **kw)
    else:
        model = vuvit.UViT_M(image_size=64, patch_size=4, in_channels=3, num_classes=1000, class_dropout_prob=0.0)
    model = model.to(dev).train()
    extra = dict(learn_align=True, gamma=0.5, align_type="mse") if repa else {}
    diffusion = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda", **extra)
    teacher = None
    if repa:
        from vaw_b200.encoders.mocov3_vit import vit_base
        teacher = vit_base().to(dev).eval()
    return model, diffusion, teacher


def assert_ranks_agree(model, sampler, dist, dev, net_or_model=None):
    """After an optimizer step every rank must hold bit-identical parameters (identical all-reduced gradients ->
    identical AdamW) and an identical sampler history (every rank applies the same gathered update, resample.py:76-79).
    Exact integer checksums over the raw bit patterns, compared across ranks; outside the timed region."""
    import torch
    if hasattr(net_or_model, "gather_master"):
        net_or_model.gather_master()          # sharded optimizer: complete the fp32 master before comparing it
    sums = [model._flat.detach().view(torch.int32).to(torch.int64).sum()]
    if getattr(sampler, "_hist_dev", None) is not None:
        sums.append(sampler._hist_dev.view(torch.int64).sum())
        sums.append(sampler._count_dev.to(torch.int64).sum())
    mine = torch.stack(sums)
    everyone = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(everyone, mine)
    for r, other in enumerate(everyone):
        if not torch.equal(other, everyone[0]):
            raise RuntimeError(f"rank {r} diverged from rank 0 after the update: {other.tolist()} vs {everyone[0].tolist()}")


def run_native(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from types import SimpleNamespace
    from vaw_b200 import _lib as L
    from vaw_b200.optim import DataParallel, FusedAdamW
    from vaw_b200.parallel import shard_seed
    from vaw_b200.tools import resample as rs

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L.call("vaw_device_check")
    seed = shard_seed(42, rank)
    torch.manual_seed(seed)
    np.random.seed(seed)

    cfg = args.cfg
    B = cfg["batch"]
    model, diffusion, teacher = build_native(args, dev)
    # N > 1: gradients reduce-scattered per block, AdamW on 1/N of the large tensors per rank, bf16 shadows all-gathered
    # (vaw_b200.parallel.ShardedGradSync); --no-shard keeps all-reduce + replicated AdamW
    net = DataParallel(model, shard_optimizer=not args.no_shard) if world > 1 else model
    sampler = rs.create_named_schedule_sampler(cfg["sampler"], diffusion)
    loss_aware = cfg["sampler"] == "loss-second-moment"
    if loss_aware:
        hist, counts = synthetic_history(0)
        sampler.load_history(hist, counts, dev)
        sampler.ragged_batches = False     # every rank passes B samples: no size exchange, no host sync
    opt = FusedAdamW(net, lr=1e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0)

    # synthetic data: a small pool of batches, resident in HBM (value) and in pinned host memory (e2e)
    pool = 4
    gen = torch.Generator().manual_seed(seed)
    shape = (B, cfg["chans"], cfg["img"], cfg["img"])
    host_x = [torch.randn(shape, generator=gen) for _ in range(pool)]
    if cfg["family"] == "uvit":
        host_x = [x.clamp_(-1, 1) for x in host_x]
    host_x = [x.pin_memory() for x in host_x]
    host_y = [torch.randint(0, 1000, (B,), generator=gen).pin_memory() for _ in range(pool)]
    host_px = None
    if teacher is not None:   # raw 0..255 pixels as fp32, what the reference's Latent_Pixel loader hands the trainer
        host_px = [torch.randint(0, 256, (B, 3, 256, 256), generator=gen).float().pin_memory() for _ in range(2)]
    dev_x = [x.to(dev) for x in host_x]
    dev_y = [y.to(dev) for y in host_y]
    dev_px = [p.to(dev) for p in host_px] if host_px else None
    enc_args = SimpleNamespace(enc_type="mocov3-vit-b")
    if teacher is not None:
        from vaw_b200.encoders.mocov3_vit import get_feature

    graphed = None
    if cfg.get("graph") and world == 1 and not args.no_graph:
        from vaw_b200.graph import GraphedTrainingLosses
        graphed = GraphedTrainingLosses(diffusion, model, shape)

    def step(x, y, px=None):
        feats = get_feature(enc_args, px, teacher) if teacher is not None else None
        t, w = sampler.sample(B, dev)
        if graphed is not None:
            terms = graphed(x, t, w, y=y)
            loss = None
        else:
            terms = diffusion.training_losses(net, x, feats, t=t, model_kwargs={"y": y})
            if loss_aware:
                sampler.update_with_local_losses(t, terms["loss"].detach())
            loss = (terms["loss"] * w).mean()
            loss.backward()
        opt.step()
        opt.zero_grad()
        return loss if loss is not None else terms["loss"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        barrier()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    px_of = (lambda i: dev_px[i % 2]) if dev_px else (lambda i: None)
    for i in range(max(3, args.warmup)):
        step(dev_x[i % pool], dev_y[i % pool], px_of(i))
        if i == 0 and world > 1:
            torch.cuda.synchronize()
            assert_ranks_agree(model, sampler, dist, dev, net)
    torch.cuda.synchronize()

    with ClockSampler(local_rank) as clk:
        n0 = L.launch_count()
        ms_dev = timed(lambda i: step(dev_x[i % pool], dev_y[i % pool], px_of(i)), args.steps)
        launches = int(L.launch_count() - n0)
        if graphed is not None:   # kernels replayed from the graph do not pass through the library's launch counter
            launches += graphed.launches_per_replay * args.steps
    clocks = clk.summary()
    if world > 1:
        assert_ranks_agree(model, sampler, dist, dev, net)

    # end-to-end: batch from pinned host memory every step, loss read back every step
    def e2e_step(i):
        x = host_x[i % pool].to(dev, non_blocking=True)
        y = host_y[i % pool].to(dev, non_blocking=True)
        px = host_px[i % 2].to(dev, non_blocking=True) if host_px else None
        return float(step(x, y, px).float().mean().item())
    for i in range(2):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps)

    imgs = B * world * args.steps
    value = imgs / (ms_dev / 1e3)
    e2e_value = imgs / (ms_e2e / 1e3)
    peaks = measured_peaks()
    gflop = cfg["gflop"]
    step_tflops_per_gpu = value / world * gflop / 1e3
    h2d = int(host_x[0].numel() * 4 + host_y[0].numel() * 8 + (host_px[0].numel() * 4 if host_px else 0))

    line = {
        "metric": "training imgs/sec", "value": value, "unit": "imgs/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": "imgs/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "clocks": clocks,
        "step_tensor_util": {"achieved_tflops_per_gpu": step_tflops_per_gpu, "train_gflop_per_img": gflop,
                             "frac_of_measured_sustained": step_tflops_per_gpu / peaks["tf_sustained"],
                             "frac_of_nominal_2250": step_tflops_per_gpu / 2250.0, "peaks": peaks["src"]},
    }
    if graphed is not None:
        line["config"]["cuda_graph"] = "K1 -> forward -> K2 -> backward replayed as one CUDA graph per step"
    if world > 1:
        line["rank_consistency"] = "flat parameters and sampler history bit-identical on all ranks after step 1 and after the timed region"
        line["config"]["data_parallel"] = ("all-reduce + replicated AdamW" if (args.no_shard or getattr(net, "_shard_sync", None) is None)
                                           else "reduce-scatter + 1/N AdamW + bf16 all-gather (sharded optimizer)")

    if rank == 0 and not args.no_kernel_leg:
        line["roofline"] = kernel_leg(args, dev, peaks)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        del net, model, opt, graphed
        torch.cuda.empty_cache()
        try:
            ips, dt, threads, kind = cpu_step_rate(cfg, cfg["cpu_batch"], 4, 1)   # bounded: ~10-30 s of host work
            line["cpu_baseline"] = {"value": ips, "unit": "imgs/s", "cores": threads, "kind": kind,
                                    "sample": f"{'the reference itself (baseline/_ref)' if kind == 'reference' else 'oracle port of the reference step'}, "
                                              f"{cfg['model']}, batch {cfg['cpu_batch']}, fp32, 1 warm-up + 4 timed steps "
                                              f"({dt:.1f} s/step)"}
        except Exception as e:   # the GPU line must not be lost to a host-side failure of the reported baseline
            line["cpu_baseline"] = {"value": None, "unit": "imgs/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {type(e).__name__}: {e}"}
    if rank == 0:
        print(json.dumps(line), flush=True)


def block_gemm_calls(batch, D, dev, tokens=256, family="dit"):
    """The twelve tcgen05 GEMM launches of one transformer block (forward + backward) exactly as dit_engine.cu /
    uvit_engine.cu issue them (shapes, operand majorness, fused epilogues, tail split-K for the weight gradients).
    -> (closures, flops)."""
    import ctypes as C
    import torch
    from vaw_b200 import _lib as L
    M, Hd, T = batch * tokens, 4 * D, tokens
    dit = family == "dit"
    EPI_RESID = L.EPI_RES                                     # U-ViT: plain residual in the proj / fc2 epilogues
    EPI_ACT = L.EPI_GELU_TANH if dit else L.EPI_GELU_ERF
    EPI_DACT = L.EPI_DGELU_TANH if dit else L.EPI_DGELU_ERF
    bf = lambda *s: torch.randn(*s, device=dev).bfloat16()
    f32 = lambda *s: torch.randn(*s, device=dev)
    # The LayerNorm outputs carry a [1, 0 x 31] block behind every row; the fc1 (and, in DiT, where qkv has a bias, the
    # qkv) weight-gradient GEMM reads them as [M, D + 32] and returns the bias gradient as an extra output column
    # (VAW_EPI_F32 row-sum form)
    ldx = D + 32 if dit else D
    ldf = D + 32
    xn, xn2, attn_o, h_act, h_pre = bf(M, ldx), bf(M, ldf), bf(M, D), bf(M, Hd), bf(M, Hd)
    gb_fc1, gb_qkv = f32(Hd), f32(3 * D)
    qkv, dqkv, dh, dy = bf(M, 3 * D), bf(M, 3 * D), bf(M, Hd), bf(M, D)
    Wqkv, Wproj, Wfc1, Wfc2 = bf(3 * D, D), bf(D, D), bf(Hd, D), bf(D, Hd)
    x_res, gate = f32(M, D), f32(batch, D)
    y_bf, x_out, o_bf = bf(M, D), f32(M, D), bf(M, D)
    gWqkv, gWproj, gWfc1, gWfc2 = f32(3 * D, D), f32(D, D), f32(Hd, D), f32(D, Hd)
    bias = {n: torch.zeros(n, device=dev) for n in (D, 3 * D, Hd)}
    ws = torch.empty(148 * 128 * 256 * 2, device=dev)

    def mk(A, lda, a_mn, Bm, ldb, b_mn, m, n, k, epi, out, out2=None, bias_=None, resid=None, gate_=None, aux=None,
           split=False):
        g = L.GemmArgs()
        g.A, g.B, g.lda, g.ldb, g.a_mn, g.b_mn = A.data_ptr(), Bm.data_ptr(), lda, ldb, a_mn, b_mn
        g.M, g.N, g.K, g.epilogue = m, n, k, epi
        n_alg = n - 32 if (epi == L.EPI_F32 and out2 is not None) else n     # the ones block is not algorithmic work
        g.out, g.out2, g.bias, g.resid = L.ptr(out), L.ptr(out2), L.ptr(bias_), L.ptr(resid)
        g.gate, g.aux, g.rows_per_sample, g.ldg = L.ptr(gate_), L.ptr(aux), T, D
        if split:
            g.k_splits, g.split_ws, g.split_ws_elems = -1, ws.data_ptr(), ws.numel()
        return (lambda: L.call("vaw_gemm_bf16", C.byref(g), L.stream_ptr())), 2.0 * m * n_alg * k

    calls = [
        mk(xn, ldx, 0, Wqkv, D, 0, M, 3 * D, D, L.EPI_BF16, qkv, bias_=bias[3 * D]),                     # qkv
        # DiT: proj / fc2 are plain bf16-output GEMMs, the gated residual update runs in the next LayerNorm pass
        (mk(attn_o, D, 0, Wproj, D, 0, M, D, D, L.EPI_BF16, y_bf, bias_=bias[D]) if dit else
         mk(attn_o, D, 0, Wproj, D, 0, M, D, D, EPI_RESID, None, x_out, bias[D], x_res, None)),          # proj
        mk(xn2, ldf, 0, Wfc1, D, 0, M, Hd, D, EPI_ACT, h_pre, h_act, bias[Hd]),                           # fc1
        (mk(h_act, Hd, 0, Wfc2, Hd, 0, M, D, Hd, L.EPI_BF16, y_bf, bias_=bias[D]) if dit else
         mk(h_act, Hd, 0, Wfc2, Hd, 0, M, D, Hd, EPI_RESID, None, x_out, bias[D], x_res, None)),         # fc2
        mk(dy, D, 1, h_act, Hd, 1, D, Hd, M, L.EPI_F32, gWfc2, split=True),                               # wgrad fc2
        mk(dy, D, 0, Wfc2, Hd, 1, M, Hd, D, EPI_DACT, dh, aux=h_pre),                                     # dgrad fc2
        mk(dh, Hd, 1, xn2, ldf, 1, Hd, ldf, M, L.EPI_F32, gWfc1, gb_fc1, split=True),                     # wgrad fc1 + bias
        mk(dh, Hd, 0, Wfc1, D, 1, M, D, Hd, L.EPI_BF16, o_bf),                                            # dgrad fc1
        mk(dy, D, 1, attn_o, D, 1, D, D, M, L.EPI_F32, gWproj, split=True),                               # wgrad proj
        mk(dy, D, 0, Wproj, D, 1, M, D, D, L.EPI_BF16, o_bf),                                             # dgrad proj
        (mk(dqkv, 3 * D, 1, xn, ldx, 1, 3 * D, ldx, M, L.EPI_F32, gWqkv, gb_qkv, split=True) if dit else
         mk(dqkv, 3 * D, 1, xn, D, 1, 3 * D, D, M, L.EPI_F32, gWqkv, split=True)),                        # wgrad qkv (+ bias)
        mk(dqkv, 3 * D, 0, Wqkv, D, 1, M, D, 3 * D, L.EPI_BF16, o_bf),                                    # dgrad qkv
    ]
    return [c for c, _ in calls], [f for _, f in calls]


def kernel_leg(args, dev, peaks):
    """Dominant kernel = gemm_bf16_tcgen05_kernel (62 % of the step, profiles/).  Its launches are timed through the
    C ABI as the twelve GEMMs of one block (forward + backward shapes and epilogues of dit_engine.cu), back to back with
    CUDA events on the launch stream; `achieved` = flops per launch / average launch duration over these launches.  The
    sequence streams ~3 GB of distinct operands per pass, far more than the 126 MB L2, so operands are cold."""
    import torch
    cfg = args.cfg
    D = cfg["D"]
    calls, flops = block_gemm_calls(cfg["batch"], D, dev, cfg["tokens"], cfg["family"])
    for _ in range(3):
        for c in calls:
            c()
    torch.cuda.synchronize()
    iters = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        for c in calls:
            c()
    e1.record()
    torch.cuda.synchronize()
    n = iters * len(calls)
    us = e0.elapsed_time(e1) / n * 1e3
    fl = sum(flops) / len(flops)
    achieved = fl / (us * 1e-6) / 1e12
    traffic = None
    # ncu `--set full` DRAM bytes of these twelve launches; only valid for the shape and library build it was taken on
    tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("model", "DiT-XL") == cfg["model"] and tj.get("batch", 64) == cfg["batch"]:
            traffic = tj.get("dram_bytes_per_launch")
    return {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel (average over the 12 GEMM launches of a DiT block, "
                                         "fwd + bwd; split-K fix-up launches included in the time)",
            "shape": "M=%d tokens, D=%d: qkv/proj/fc1/fc2 + their dgrad/wgrad" % (cfg["batch"] * cfg["tokens"], D),
            "achieved": achieved, "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_burst"],
            "us_per_launch": us, "flops_per_launch": fl, "launches_timed": n, "traffic": traffic,
            "peak_source": peaks["src"] + " burst (kernel timed alone)"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_native(args, rank, world, local_rank)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
