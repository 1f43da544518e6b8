"""One U-ViT-M/4 (config 4) training step at B = 64: ncu target for the attention kernels of the 258-token sequence."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200")]
import torch
from vaw_b200.models.uvit import UViT_M
from vaw_b200.tools import gaussian_diffusion as gd
dev = torch.device("cuda", 0); B = int(os.environ.get("B", 64))
net = UViT_M(image_size=64, patch_size=4, in_channels=3, num_classes=1000, class_dropout_prob=0.0).to(dev).train()
d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda")
x = torch.randn(B, 3, 64, 64, device=dev).clamp(-1, 1); y = torch.randint(0, 1000, (B,), device=dev)
for _ in range(int(os.environ.get("STEPS", 2))):
    terms = d.training_losses(net, x, None, model_kwargs={"y": y})
    terms["loss"].mean().backward()
    for p in net.parameters():
        p.grad = None
torch.cuda.synchronize()
print("ok", float(terms["loss"].mean()))
