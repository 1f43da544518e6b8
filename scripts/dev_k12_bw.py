"""K1 / K2 / K3 measurements asked for by SURVEY 8(d): achieved GB/s at a saturating size (N = 16384 latents) and the
launch latency at the config sizes; K3 in microseconds."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from vaw_b200.tools import gaussian_diffusion as gd, resample as rs
dev = "cuda"
def timeit(fn, iters=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3
d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda")
dv = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="velocity", weight_type="lambda")
for N, shape in ((16384, (4, 32, 32)), (64, (4, 32, 32)), (16, (3, 32, 32)), (64, (3, 64, 64))):
    x0 = torch.randn(N, *shape, device=dev); eps = torch.randn_like(x0); t = torch.randint(0, 1000, (N,), device=dev)
    el = x0.numel()
    us = timeit(lambda: d.q_sample(x0, t, eps))
    print(f"K1 q_sample eps-pred  N={N:6d} {shape}: {us:8.1f} us  {12*el/us/1e3:7.0f} GB/s (12 B/element)")
    us = timeit(lambda: (dv.q_sample(x0, t, eps), dv.compute_target(x0, eps, t)))
    print(f"K1 q_sample+v target  N={N:6d} {shape}: {us:8.1f} us")
    for dt, bpe in ((torch.float32, 12), (torch.bfloat16, 8)):
        out = torch.randn(N, *shape, device=dev, dtype=dt).requires_grad_(True)
        def k2():
            terms = d.training_losses(lambda x, ts, **k: out, x0, None, t=t, noise=eps)
            terms["loss"].mean().backward()
            out.grad = None
        us = timeit(k2, iters=20)
        print(f"K2 training_losses(model=identity)+backward {str(dt)[6:]:8s} N={N:6d}: {us:8.1f} us (incl. K1 and torch's mean/backward glue; K2 alone moves {bpe} B/element = {bpe*el/1e6:.1f} MB)")
s = rs.LossSecondMomentResampler(d)
from oracle.train_step import synthetic_history
h, c = synthetic_history(0)
s.load_history(h, c, dev)
np.random.seed(0)
us = timeit(lambda: s.sample(64, dev), iters=50)
print(f"K3 sample(64) warmed-up history: {us:.1f} us (includes the host MT19937 draw and a 512-byte H2D copy)")
tt = torch.randint(0, 1000, (64,), device=dev); ll = torch.rand(64, device=dev)
us = timeit(lambda: s.update_with_local_losses(tt, ll), iters=50)
print(f"K3 update_with_local_losses(64): {us:.1f} us")
