"""Timeline of ONE data-parallel training step on rank 0 (CUPTI through torch.profiler; there is no nsys in the image):
when the NCCL kernels run, how long they take and how much of that lies under the compute stream's kernels.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      scripts/dev_dp_timeline.py            # SHARD=0 for the all-reduce path, CONFIG=3|4

Prints a text summary (commit it under profiles/); the step measured under the profiler is slower than a bench step,
so the numbers that matter are the shares and the overlap, not the absolute step time."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from vaw_b200.optim import DataParallel, FusedAdamW  # noqa: E402
from vaw_b200.parallel import shard_seed  # noqa: E402
from vaw_b200.tools import resample as rs  # noqa: E402


def union_len(iv):
    iv = sorted(iv)
    tot, cur_b, cur_e = 0.0, None, None
    for b, e in iv:
        if cur_e is None or b > cur_e:
            if cur_e is not None:
                tot += cur_e - cur_b
            cur_b, cur_e = b, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        tot += cur_e - cur_b
    return tot


def overlap_len(a, b):
    """length of (union of a) ∩ (union of b)"""
    return union_len(a) + union_len(b) - union_len(a + b)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    shard = os.environ.get("SHARD", "1") == "1"
    sys.argv = ["bench.py", "--config", os.environ.get("CONFIG", "3")]
    args = bench.parse()
    cfg = args.cfg
    seed = shard_seed(42, rank)
    torch.manual_seed(seed)
    np.random.seed(seed)
    model, diffusion, _ = bench.build_native(args, dev)
    net = DataParallel(model, shard_optimizer=shard)
    sampler = rs.create_named_schedule_sampler(cfg["sampler"], diffusion)
    loss_aware = cfg["sampler"] == "loss-second-moment"
    if loss_aware:
        sampler.load_history(*bench.synthetic_history(0), dev)
        sampler.ragged_batches = False
    opt = FusedAdamW(net, lr=1e-4, betas=(0.9, 0.95))
    B = cfg["batch"]
    x = torch.randn(B, cfg["chans"], cfg["img"], cfg["img"], device=dev)
    y = torch.randint(0, 1000, (B,), device=dev)

    def step():
        t, w = sampler.sample(B, dev)
        terms = diffusion.training_losses(net, x, None, t=t, model_kwargs={"y": y})
        if loss_aware:
            sampler.update_with_local_losses(t, terms["loss"].detach())
        (terms["loss"] * w).mean().backward()
        opt.step()
        opt.zero_grad()

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    if rank != 0:        # the other ranks run the same three steps (the collectives need them), unprofiled
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    else:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
        path = os.path.join(tempfile.gettempdir(), "vaw_dp_trace.json")
        prof.export_chrome_trace(path)
        ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
        ev.sort(key=lambda e: e["ts"])
        # the middle step: from the second sampler_sample kernel to the third
        marks = [e["ts"] for e in ev if "sampler_sample" in e["name"]]
        lo, hi = (marks[1], marks[2]) if len(marks) >= 3 else (ev[0]["ts"], ev[-1]["ts"] + ev[-1]["dur"])
        ks = [e for e in ev if lo <= e["ts"] < hi]
        nccl = [e for e in ks if "nccl" in e["name"].lower()]
        comp = [e for e in ks if "nccl" not in e["name"].lower()]
        iv = lambda es: [(e["ts"], e["ts"] + e["dur"]) for e in es]  # noqa: E731
        span = hi - lo
        print(f"data-parallel step timeline, rank 0 of {world}, config {os.environ.get('CONFIG', '3')}, "
              f"{'sharded optimizer (reduce-scatter / all-gather)' if shard else 'all-reduce + replicated AdamW'}")
        print(f"step span under the profiler: {span / 1e3:.2f} ms; {len(comp)} compute kernels busy "
              f"{union_len(iv(comp)) / 1e3:.2f} ms; {len(nccl)} NCCL kernels busy {union_len(iv(nccl)) / 1e3:.2f} ms, "
              f"of which {overlap_len(iv(nccl), iv(comp)) / 1e3:.2f} ms under compute kernels")
        idle = span - union_len(iv(ks))
        print(f"no kernel at all running: {idle / 1e3:.2f} ms")
        by = {}
        for e in nccl:
            n = e["name"].split("(")[0][:60]
            c = by.setdefault(n, [0, 0.0])
            c[0] += 1
            c[1] += e["dur"]
        for n, (c, d) in sorted(by.items(), key=lambda kv: -kv[1][1]):
            print(f"  {d / 1e3:8.2f} ms  n={c:4d}  avg {d / c:8.1f} us  {n}")
        # phases: where the NCCL kernels sit relative to forward / backward / optimizer
        first = lambda pat: next((e["ts"] for e in ks if pat in e["name"]), None)  # noqa: E731
        t_bwd = first("attn_bwd") or first("ln_bwd")
        t_opt = first("adamw")
        print(f"phase starts (ms after the step's first kernel): backward's first attention/LN kernel "
              f"{(t_bwd - lo) / 1e3 if t_bwd else -1:.2f}, optimizer {(t_opt - lo) / 1e3 if t_opt else -1:.2f}")
        print("NCCL kernels in time order (start ms, duration us, stream):")
        for e in nccl:
            print(f"  {(e['ts'] - lo) / 1e3:8.2f}  {e['dur']:9.1f}  s{e.get('args', {}).get('stream', '?')}  "
                  f"{e['name'].split('(')[0][:48]}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
