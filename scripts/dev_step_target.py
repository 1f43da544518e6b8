"""ncu target: 1 warm-up + 1 measured DiT-XL/2 B=64 forward+backward (no optimizer), for per-kernel launch lists."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
import torch
from vaw_b200.models.dit import DiT_XL
from gpu_util import dezero
dev = "cuda"
B = 64
m = DiT_XL(image_size=32, patch_size=2, in_channels=4, class_dropout_prob=0.0, num_classes=1000, learn_sigma=False).to(dev)
dezero(m)
xx = torch.randn(B, 4, 32, 32, device=dev); tt = torch.rand(B, device=dev) * 999; yy = torch.randint(0, 1000, (B,), device=dev)
g = torch.randn(B, 4, 32, 32, device=dev).bfloat16()
for _ in range(int(os.environ.get("STEPS", 2))):
    out, _ = m(xx, tt, yy); out.backward(g)
    torch.cuda.synchronize()
print("done")
