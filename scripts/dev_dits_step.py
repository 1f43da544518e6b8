"""Two training steps of DiT-S/2 (config 2) at B = 256 for an ncu launch list (where does a small-D step spend time)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
import torch
from vaw_b200.models.dit import DiT_S
from vaw_b200.optim import FusedAdamW
from vaw_b200.tools import gaussian_diffusion as gd, resample as rs
from gpu_util import dezero
dev = torch.device("cuda", 0); B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda")
net = DiT_S(image_size=32, patch_size=2, in_channels=4, class_dropout_prob=0.0, num_classes=1000, learn_sigma=False).to(dev)
dezero(net)
s = rs.UniformSampler(d); opt = FusedAdamW(net, lr=1e-4, betas=(0.9, 0.95))
x = torch.randn(B, 4, 32, 32, device=dev); y = torch.randint(0, 1000, (B,), device=dev)
for _ in range(2):
    t, w = s.sample(B, dev)
    terms = d.training_losses(net, x, None, t=t, model_kwargs={"y": y})
    (terms["loss"] * w).mean().backward(); opt.step(); opt.zero_grad()
torch.cuda.synchronize()
print("done")
