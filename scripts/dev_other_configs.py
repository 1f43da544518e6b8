"""Throughput of the other BASELINE configs on one B200 (forward + backward + fused AdamW through the public API):
config 2 = DiT-S/2 (B = 64 and 256, UniformSampler), config 4 = U-ViT-M/4 on 3x64x64 (B = 64)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from vaw_b200.models.dit import DiT_S
from vaw_b200.models.uvit import UViT_M
from vaw_b200.optim import FusedAdamW
from vaw_b200.tools import gaussian_diffusion as gd, resample as rs
from gpu_util import dezero
dev = torch.device("cuda", 0)
d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda")
def run(name, net, shape, B, gflop):
    net = net.to(dev); dezero(net)
    s = rs.UniformSampler(d); opt = FusedAdamW(net, lr=1e-4, betas=(0.9, 0.95))
    x = torch.randn(B, *shape, device=dev).clamp(-1, 1); y = torch.randint(0, 1000, (B,), device=dev)
    def step():
        t, w = s.sample(B, dev)
        terms = d.training_losses(net, x, None, t=t, model_kwargs={"y": y})
        (terms["loss"] * w).mean().backward(); opt.step(); opt.zero_grad()
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10; e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:28s} B={B:4d}: {ms:7.2f} ms/step  {B / ms * 1e3:8.0f} img/s  {B / ms * gflop:6.0f} TFLOP/s", flush=True)
    del net, opt; torch.cuda.empty_cache()
mk = lambda: DiT_S(image_size=32, patch_size=2, in_channels=4, class_dropout_prob=0.0, num_classes=1000, learn_sigma=False)
run("DiT-S/2 (config 2)", mk(), (4, 32, 32), 64, 36.32)
run("DiT-S/2 (config 2)", mk(), (4, 32, 32), 256, 36.32)
run("U-ViT-M/4 @64 (config 4)", UViT_M(image_size=64, patch_size=4, in_channels=3, num_classes=1000, class_dropout_prob=0.0), (3, 64, 64), 64, 211.4)
