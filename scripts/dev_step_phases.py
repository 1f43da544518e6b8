"""Where does the bench step's time go?  CUDA-event timing of the phases of the training step (steady state)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from vaw_b200.models.dit import DiT_XL
from vaw_b200.optim import FusedAdamW
from vaw_b200.tools import gaussian_diffusion as gd, resample as rs
from oracle.train_step import synthetic_history
from gpu_util import dezero
dev = torch.device("cuda", 0); B = 64
net = DiT_XL(image_size=32, patch_size=2, in_channels=4, class_dropout_prob=0.0, num_classes=1000, learn_sigma=False).to(dev)
dezero(net)
d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda")
s = rs.LossSecondMomentResampler(d); h, c = synthetic_history(0); s.load_history(h, c, dev)
opt = FusedAdamW(net, lr=1e-4, betas=(0.9, 0.95))
x = torch.randn(B, 4, 32, 32, device=dev); y = torch.randint(0, 1000, (B,), device=dev)
names = ["sample", "losses(fwd)", "sampler update", "loss glue", "backward", "optimizer"]
def step(ev):
    ev[0].record(); t, w = s.sample(B, dev)
    ev[1].record(); terms = d.training_losses(net, x, None, t=t, model_kwargs={"y": y})
    ev[2].record(); s.update_with_local_losses(t, terms["loss"].detach())
    ev[3].record(); loss = (terms["loss"] * w).mean()
    ev[4].record(); loss.backward()
    ev[5].record(); opt.step(); opt.zero_grad()
    ev[6].record()
for _ in range(3): step([torch.cuda.Event(enable_timing=True) for _ in range(7)])
torch.cuda.synchronize()
N = 8
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(7)] for _ in range(N)]
import time
t0 = time.perf_counter()
for i in range(N): step(evs[i])
t_cpu = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
acc = np.zeros(6)
for e in evs:
    for k in range(6): acc[k] += e[k].elapsed_time(e[k + 1])
gaps = sum(evs[i][6].elapsed_time(evs[i + 1][0]) for i in range(N - 1)) / (N - 1)
print("phase ms/step:", {n: round(a / N, 3) for n, a in zip(names, acc)}, " sum", round(acc.sum() / N, 3), " between steps", round(gaps, 3))
print(f"CPU enqueue time per step {t_cpu / N * 1e3:.2f} ms; wall per step {t_all / N * 1e3:.2f} ms")
