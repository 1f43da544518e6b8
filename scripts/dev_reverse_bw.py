"""K8 (fused reverse step) and the guidance combine: achieved GB/s at a saturating size and launch latency at the
sampling batch sizes; then the forward-only DiT step a DDIM sampler runs (IntervalCFG doubled batch + K8)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
import torch
from vaw_b200 import _lib as L
from vaw_b200.tools import gaussian_diffusion as gd
from vaw_b200.tools.respace import SpacedDiffusion, space_timesteps
from vaw_b200.tools.sampler import IntervalCFG
dev = "cuda"
def timeit(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3
d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda")
kw = dict(clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None)
for N in (16384, 64, 8):
    x = torch.randn(N, 4, 32, 32, device=dev); z = torch.randn_like(x); t = torch.randint(0, 1000, (N,), device=dev)
    el = x.numel()
    for dt, ob in ((torch.float32, 4), (torch.bfloat16, 2)):
        mo = torch.randn(N, 4, 32, 32, device=dev, dtype=dt)
        model = lambda a, b, **k: mo
        for name, mode, eta, extra in (("ddim eta=0", L.RS_DDIM, 0.0, 0), ("ddim eta=.5", L.RS_DDIM, 0.5, 4), ("p_sample", L.RS_DDPM, 0.0, 4)):
            bpe = ob + 4 + extra + 8   # model output + x_t (+ noise) in, sample + pred_xstart out
            us = timeit(lambda: d._reverse(mode, model, x, t, eta=eta, noise=z, **kw))
            print(f"K8 {name:11s} out={str(dt)[6:]:8s} N={N:6d}: {us:8.1f} us  {bpe*el/us/1e3:7.0f} GB/s ({bpe} B/element)")
    both = torch.randn(2 * N, 4, 32, 32, device=dev, dtype=torch.bfloat16)
    cfg = IntervalCFG(lambda a, b, **k: both, 1000, 1.5)
    y = torch.zeros(N, dtype=torch.long, device=dev)
    us = timeit(lambda: cfg(x, t.float(), y=y))
    print(f"CFG combine bf16 N={N:6d}: {us:8.1f} us  {6*el/us/1e3:7.0f} GB/s (6 B/element; includes the two torch.cat of the wrapper)")
if "--model" in sys.argv:
    from vaw_b200.models import dit
    m = dit.DiT_XL(32, 2, 4, 0.1, 1000, False).to(dev).eval()
    sd = SpacedDiffusion(use_timesteps=space_timesteps(1000, "ddim50"), args=gd.default_args(),
                         betas=gd.get_named_beta_schedule("cosine", 1000), model_mean_type=gd.ModelMeanType.EPSILON,
                         model_var_type=gd.ModelVarType.FIXED_LARGE, loss_type=gd.LossType.MSE, rescale_timesteps=True)
    for B in (32, 64):
        y = torch.randint(0, 1000, (B,), device=dev)
        cfgm = IntervalCFG(m, 1000, 1.5).eval()
        x = torch.randn(B, 4, 32, 32, device=dev); t = torch.full((B,), 25, device=dev)
        with torch.no_grad():
            us_f = timeit(lambda: m(torch.cat([x, x]), torch.cat([t, t]).float(), torch.cat([y, y])), iters=10, warm=3)
            us_s = timeit(lambda: sd.ddim_sample(cfgm, x, t, model_kwargs={"y": y}), iters=10, warm=3)
        fl = 2 * B * 237.23e9
        print(f"DiT-XL/2 guided DDIM step B={B}: forward(2B) {us_f/1e3:.2f} ms ({fl/us_f/1e6:.0f} TFLOP/s), whole step {us_s/1e3:.2f} ms "
              f"-> {B/us_s*1e6:.0f} img-steps/s; ddim50 = {B/(50*us_s)*1e6:.1f} img/s")
