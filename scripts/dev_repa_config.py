"""Config 5 (DiT-XL/2 + REPA projector, align loss 'mse'): throughput of the training step with synthetic teacher
features, the frozen MoCo-v3 ViT-B/16 teacher forward alone, and the full step with the teacher inside."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
import torch
from vaw_b200.models.dit import DiT_XL
from vaw_b200.optim import FusedAdamW
from vaw_b200.tools import gaussian_diffusion as gd, resample as rs
from gpu_util import dezero
dev = torch.device("cuda", 0); B = 64
net = DiT_XL(image_size=32, patch_size=2, in_channels=4, class_dropout_prob=0.0, num_classes=1000, learn_sigma=False,
             learn_align=True, encoder_depth=8, z_dims=768, projector_dim=2048).to(dev)
dezero(net)
d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda", learn_align=True,
                                 gamma=0.5, align_type="mse")
s = rs.UniformSampler(d); opt = FusedAdamW(net, lr=1e-4, betas=(0.9, 0.95))
x = torch.randn(B, 4, 32, 32, device=dev); y = torch.randint(0, 1000, (B,), device=dev)
feats = torch.randn(B, 256, 768, device=dev)
def step():
    t, w = s.sample(B, dev)
    terms = d.training_losses(net, x, feats, t=t, model_kwargs={"y": y})
    loss = (terms["loss"] * w).mean(); loss.backward(); opt.step(); opt.zero_grad()
    return terms
for _ in range(3): terms = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 8; e0.record()
for _ in range(n): terms = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"DiT-XL/2 + REPA (config 5, synthetic teacher features) B={B}: {ms:.2f} ms/step  {B / ms * 1e3:.0f} img/s  "
      f"{B / ms * 724.2:.0f} TFLOP/s   mse {terms['mse'].mean().item():.4f} align {float(terms['align']):.4f}")

# the frozen teacher: raw 256-px pixels -> [B, 256, 768] features, every step (align_utils.py:43-50)
from types import SimpleNamespace
from vaw_b200.encoders.mocov3_vit import vit_base, get_feature
teacher = vit_base().to(dev).eval()
pixels = torch.randint(0, 256, (B, 3, 256, 256), device=dev).float()
ns = SimpleNamespace(enc_type="mocov3-vit-b")
for _ in range(3): f = get_feature(ns, pixels, teacher)
torch.cuda.synchronize(); e0.record()
for _ in range(n): f = get_feature(ns, pixels, teacher)
e1.record(); torch.cuda.synchronize()
tms = e0.elapsed_time(e1) / n
print(f"MoCo-v3 ViT-B/16 teacher forward B={B}: {tms:.2f} ms  {B / tms * 46.4:.0f} TFLOP/s (46.4 GFLOP/img)")
def full_step():
    global feats
    feats = get_feature(ns, pixels, teacher)
    return step()
for _ in range(3): terms = full_step()
torch.cuda.synchronize(); e0.record()
for _ in range(n): terms = full_step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"DiT-XL/2 + REPA with the teacher in the step B={B}: {ms:.2f} ms/step  {B / ms * 1e3:.0f} img/s  "
      f"{B / ms * (724.2 + 46.4):.0f} TFLOP/s   mse {terms['mse'].mean().item():.4f} align {float(terms['align']):.4f}")
