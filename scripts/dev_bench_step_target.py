"""ncu target: the bench.py step of a config (sampler -> [teacher] -> training_losses -> backward -> fused AdamW ->
sampler update), WARM warm-up steps + 1 measured step, for per-kernel launch lists without paying for bench.py's timed /
e2e / roofline legs under the profiler.  CONFIG=3|4|5 (default 3), WARM (default 1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200")]
import numpy as np
import torch
import bench
from types import SimpleNamespace
from vaw_b200.optim import FusedAdamW
from vaw_b200.tools import resample as rs
sys.argv = ["bench.py", "--config", os.environ.get("CONFIG", "3")]
args = bench.parse()
cfg = args.cfg
dev = torch.device("cuda", 0)
torch.manual_seed(42); np.random.seed(42)
model, diffusion, teacher = bench.build_native(args, dev)
sampler = rs.create_named_schedule_sampler(cfg["sampler"], diffusion)
if cfg["sampler"] == "loss-second-moment":
    sampler.load_history(*bench.synthetic_history(0), dev)
    sampler.ragged_batches = False
opt = FusedAdamW(model, lr=1e-4, betas=(0.9, 0.95))
B = cfg["batch"]
x = torch.randn(B, cfg["chans"], cfg["img"], cfg["img"], device=dev)
y = torch.randint(0, 1000, (B,), device=dev)
px = torch.randint(0, 256, (B, 3, 256, 256), device=dev).float() if teacher is not None else None
if teacher is not None:
    from vaw_b200.encoders.mocov3_vit import get_feature
for i in range(int(os.environ.get("WARM", 1)) + 1):
    feats = get_feature(SimpleNamespace(enc_type="mocov3-vit-b"), px, teacher) if teacher is not None else None
    t, w = sampler.sample(B, dev)
    terms = diffusion.training_losses(model, x, feats, t=t, model_kwargs={"y": y})
    if cfg["sampler"] == "loss-second-moment":
        sampler.update_with_local_losses(t, terms["loss"].detach())
    (terms["loss"] * w).mean().backward()
    opt.step(); opt.zero_grad()
    torch.cuda.synchronize()
    print("step", i, "done", flush=True)
