#!/usr/bin/env python
"""bench.py — training imgs/s of the diffusion training step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the configuration the metric is quoted on; it fits one GPU):
  DiT-XL/2 on synthetic 32x32x4 latents, class-conditional (1000 classes), eps-prediction, cosine schedule,
  weight_type=lambda, LossSecondMomentResampler with a warmed-up history, batch 64 per GPU, bf16 tensor-core math
  with fp32 master weights, fused AdamW.  Random-init weights (reference init), synthetic data.

One step = sampler.sample -> training_losses (K1, DiT forward, K2) -> backward -> [grad all-reduce] -> AdamW ->
sampler.update_with_local_losses.  `value` times K steps with inputs resident in HBM; `e2e` times the same steps
through the public API with the batch coming from pinned host memory and the loss read back every step.

--impl reference: the CPU arm — the oracle port of the reference's training step (oracle/train_step.py; the
reference itself cannot travel to the GPU box) on all host threads, on a bounded batch of the same workload.
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

TRAIN_GFLOP_PER_IMG = {"DiT-S": 36.32, "DiT-B": 138.0, "DiT-L": 484.2, "DiT-XL": 711.7}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--model", default="DiT-XL", choices=list(TRAIN_GFLOP_PER_IMG))
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch (weak scaling)")
    ap.add_argument("--sampler", default="loss-second-moment", choices=["uniform", "loss-second-moment"])
    ap.add_argument("--cpu-batch", type=int, default=4, help="batch of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-leg", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sustained=d.get("bf16_tflops_sustained", 1400.0), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


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
# CPU arm: oracle port of the reference training step
# ------------------------------------------------------------------------------------------------------------
def cpu_step_rate(model, batch, steps, warmup):
    import numpy as np
    import torch
    from oracle.train_step import OracleTrainer
    torch.set_num_threads(os.cpu_count() or 1)
    np.random.seed(42)
    torch.manual_seed(42)
    tr = OracleTrainer(model, seed=0)
    x = torch.randn(batch, 4, 32, 32)
    y = torch.randint(0, 1000, (batch,))
    for _ in range(warmup):
        tr.step(x, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(x, y)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return batch / dt, dt, torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    ips, dt, threads = cpu_step_rate(args.model, args.cpu_batch, args.steps, args.warmup)
    sample = (f"{args.model}/2 32x32x4 latents, batch {args.cpu_batch}, fp32, oracle port of the reference step "
              f"(forward+backward+AdamW+sampler), {args.warmup} warm-up + {args.steps} timed steps")
    line = {
        "impl": "reference", "metric": "training imgs/sec", "value": ips, "unit": "imgs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cpu=True),
        "cpu_baseline": {"value": ips, "unit": "imgs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "imgs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, cpu=False):
    return {"workload": f"{args.model}/2 diffusion training step on 32x32x4 latents (256px), class-cond 1000, "
                        f"eps-pred, cosine schedule, weight_type=lambda, {args.sampler} sampler",
            "per_gpu_batch": args.cpu_batch if cpu else args.batch,
            "global_batch": (args.cpu_batch if cpu else args.batch * args.gpus),
            "parallelism": f"dp{args.gpus}", "optimizer": "AdamW lr=1e-4 betas=(0.9,0.95)",
            "l2": "per-step working set (>20 GB of activations) exceeds the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------------------
def run_native(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from vaw_b200 import _lib as L
    from vaw_b200.models import dit as vdit
    from vaw_b200.optim import DataParallel, FusedAdamW
    from vaw_b200.parallel import shard_seed
    from vaw_b200.tools import gaussian_diffusion as gd
    from vaw_b200.tools import resample as rs
    from oracle.train_step import synthetic_history  # fixture only: the warmed-up sampler history

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L.call("vaw_device_check")
    seed = shard_seed(42, rank)
    torch.manual_seed(seed)
    np.random.seed(seed)

    B = args.batch
    model = vdit.DiT_models[args.model](image_size=32, patch_size=2, in_channels=4, class_dropout_prob=0.0,
                                        num_classes=1000, learn_sigma=False).to(dev)
    model.train()
    net = DataParallel(model) if world > 1 else model
    diffusion = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda")
    sampler = rs.create_named_schedule_sampler(args.sampler, diffusion)
    if args.sampler == "loss-second-moment":
        hist, counts = synthetic_history(0)
        sampler.load_history(hist, counts, dev)
    opt = FusedAdamW(model, lr=1e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0)

    # synthetic data: a small pool of batches, resident in HBM (value) and in pinned host memory (e2e)
    pool = 4
    gen = torch.Generator().manual_seed(seed)
    host_x = [torch.randn(B, 4, 32, 32, generator=gen).pin_memory() for _ in range(pool)]
    host_y = [torch.randint(0, 1000, (B,), generator=gen).pin_memory() for _ in range(pool)]
    dev_x = [x.to(dev) for x in host_x]
    dev_y = [y.to(dev) for y in host_y]

    def step(x, y):
        t, w = sampler.sample(B, dev)
        terms = diffusion.training_losses(net, x, None, t=t, model_kwargs={"y": y})
        if args.sampler == "loss-second-moment":
            sampler.update_with_local_losses(t, terms["loss"].detach())
        loss = (terms["loss"] * w).mean()
        loss.backward()
        opt.step()
        opt.zero_grad()
        return loss

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

    for i in range(max(3, args.warmup)):
        step(dev_x[i % pool], dev_y[i % pool])
    torch.cuda.synchronize()

    lc = L.lib().vaw_launch_count
    lc.restype = __import__("ctypes").c_ulonglong
    with ClockSampler(local_rank) as clk:
        n0 = lc()
        ms_dev = timed(lambda i: step(dev_x[i % pool], dev_y[i % pool]), args.steps)
        launches = int(lc() - n0)
    clocks = clk.summary()

    # end-to-end: batch from pinned host memory every step, loss read back every step
    def e2e_step(i):
        x = host_x[i % pool].to(dev, non_blocking=True)
        y = host_y[i % pool].to(dev, non_blocking=True)
        return float(step(x, y).item())
    for i in range(2):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps)

    imgs = B * world * args.steps
    value = imgs / (ms_dev / 1e3)
    e2e_value = imgs / (ms_e2e / 1e3)
    peaks = measured_peaks()
    gflop = TRAIN_GFLOP_PER_IMG[args.model]
    step_tflops_per_gpu = value / world * gflop / 1e3

    line = {
        "metric": "training imgs/sec", "value": value, "unit": "imgs/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": "imgs/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(host_x[0].numel() * 4 + host_y[0].numel() * 8), "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "clocks": clocks,
        "step_tensor_util": {"achieved_tflops_per_gpu": step_tflops_per_gpu, "train_gflop_per_img": gflop,
                             "frac_of_measured_sustained": step_tflops_per_gpu / peaks["tf_sustained"],
                             "frac_of_nominal_2250": step_tflops_per_gpu / 2250.0, "peaks": peaks["src"]},
    }

    if rank == 0 and not args.no_kernel_leg:
        line["roofline"] = kernel_leg(args, dev, peaks)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        del net, model, opt
        torch.cuda.empty_cache()
        ips, dt, threads = cpu_step_rate(args.model, args.cpu_batch, 6, 1)   # ~12 s of host work
        line["cpu_baseline"] = {"value": ips, "unit": "imgs/s", "cores": threads, "kind": "port",
                                "sample": f"oracle port of the reference step, {args.model}/2, batch {args.cpu_batch}, "
                                          f"fp32, 1 warm-up + 6 timed steps ({dt:.1f} s/step)"}
    if rank == 0:
        print(json.dumps(line), flush=True)


def block_gemm_calls(batch, D, dev):
    """The twelve tcgen05 GEMM launches of one DiT block (forward + backward) exactly as dit_engine.cu issues them
    (shapes, operand majorness, fused epilogues, tail split-K for the weight gradients).  -> (closures, flops)."""
    import ctypes as C
    import torch
    from vaw_b200 import _lib as L
    M, Hd, T = batch * 256, 4 * D, 256
    bf = lambda *s: torch.randn(*s, device=dev).bfloat16()
    f32 = lambda *s: torch.randn(*s, device=dev)
    xn, attn_o, h_act, h_pre = bf(M, D), bf(M, D), bf(M, Hd), bf(M, Hd)
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
        g.out, g.out2, g.bias, g.resid = L.ptr(out), L.ptr(out2), L.ptr(bias_), L.ptr(resid)
        g.gate, g.aux, g.rows_per_sample, g.ldg = L.ptr(gate_), L.ptr(aux), T, D
        if split:
            g.k_splits, g.split_ws, g.split_ws_elems = -1, ws.data_ptr(), ws.numel()
        return (lambda: L.call("vaw_gemm_bf16", C.byref(g), L.stream_ptr())), 2.0 * m * n * k

    calls = [
        mk(xn, D, 0, Wqkv, D, 0, M, 3 * D, D, L.EPI_BF16, qkv, bias_=bias[3 * D]),                       # qkv
        mk(attn_o, D, 0, Wproj, D, 0, M, D, D, L.EPI_GATE_RES, y_bf, x_out, bias[D], x_res, gate),        # proj
        mk(xn, D, 0, Wfc1, D, 0, M, Hd, D, L.EPI_GELU_TANH, h_pre, h_act, bias[Hd]),                      # fc1
        mk(h_act, Hd, 0, Wfc2, Hd, 0, M, D, Hd, L.EPI_GATE_RES, y_bf, x_out, bias[D], x_res, gate),       # fc2
        mk(dy, D, 1, h_act, Hd, 1, D, Hd, M, L.EPI_F32, gWfc2, split=True),                               # wgrad fc2
        mk(dy, D, 0, Wfc2, Hd, 1, M, Hd, D, L.EPI_DGELU_TANH, dh, aux=h_pre),                             # dgrad fc2
        mk(dh, Hd, 1, xn, D, 1, Hd, D, M, L.EPI_F32, gWfc1, split=True),                                  # wgrad fc1
        mk(dh, Hd, 0, Wfc1, D, 1, M, D, Hd, L.EPI_BF16, o_bf),                                            # dgrad fc1
        mk(dy, D, 1, attn_o, D, 1, D, D, M, L.EPI_F32, gWproj, split=True),                               # wgrad proj
        mk(dy, D, 0, Wproj, D, 1, M, D, D, L.EPI_BF16, o_bf),                                             # dgrad proj
        mk(dqkv, 3 * D, 1, xn, D, 1, 3 * D, D, M, L.EPI_F32, gWqkv, split=True),                          # wgrad qkv
        mk(dqkv, 3 * D, 0, Wqkv, D, 1, M, D, 3 * D, L.EPI_BF16, o_bf),                                    # dgrad qkv
    ]
    return [c for c, _ in calls], [f for _, f in calls]


def kernel_leg(args, dev, peaks):
    """Dominant kernel = gemm_bf16_tcgen05_kernel (62 % of the step, profiles/).  Its launches are timed through the
    C ABI as the twelve GEMMs of one block (forward + backward shapes and epilogues of dit_engine.cu), back to back with
    CUDA events on the launch stream; `achieved` = flops per launch / average launch duration over these launches.  The
    sequence streams ~3 GB of distinct operands per pass, far more than the 126 MB L2, so operands are cold."""
    import torch
    D = {"DiT-S": 384, "DiT-B": 768, "DiT-L": 1024, "DiT-XL": 1152}[args.model]
    calls, flops = block_gemm_calls(args.batch, D, dev)
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
    tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    return {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel (average over the 12 GEMM launches of a DiT block, "
                                         "fwd + bwd; split-K fix-up launches included in the time)",
            "shape": "M=%d tokens, D=%d: qkv/proj/fc1/fc2 + their dgrad/wgrad" % (args.batch * 256, D),
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
