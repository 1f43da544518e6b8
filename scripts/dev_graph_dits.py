"""Experiment: CUDA-graph capture of K1 -> DiT forward -> K2 -> backward (the launch-heavy part of a DiT-S step),
optimizer outside the graph.  Prints eager vs replayed time per step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
import torch
from vaw_b200 import _lib as L
from vaw_b200.models.dit import DiT_S
from vaw_b200.optim import FusedAdamW
from vaw_b200.tools import gaussian_diffusion as gd, resample as rs
from gpu_util import dezero
dev = torch.device("cuda", 0)
d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda")
def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for B in (64, 256):
    net = DiT_S(image_size=32, patch_size=2, in_channels=4, class_dropout_prob=0.0, num_classes=1000, learn_sigma=False).to(dev)
    dezero(net)
    s = rs.UniformSampler(d); opt = FusedAdamW(net, lr=1e-4, betas=(0.9, 0.95))
    x = torch.randn(B, 4, 32, 32, device=dev); y = torch.randint(0, 1000, (B,), device=dev)
    st_t = torch.zeros(B, dtype=torch.int64, device=dev); st_w = torch.ones(B, device=dev)
    st_eps = torch.randn_like(x)
    def fwd_bwd():
        terms = d.training_losses(net, x, None, t=st_t, model_kwargs={"y": y}, noise=st_eps)
        (terms["loss"] * st_w).mean().backward()
        return terms
    def eager():
        t, w = s.sample(B, dev); st_t.copy_(t); st_w.copy_(w); st_eps.normal_()
        fwd_bwd(); opt.step(); opt.zero_grad()
    ms_e = timeit(eager)
    n0 = L.launch_count(); fwd_bwd(); opt.zero_grad(); nl = L.launch_count() - n0
    # capture
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): fwd_bwd(); opt.step(); opt.zero_grad()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            st_terms = fwd_bwd()
    except Exception as e:
        print("capture failed:", repr(e)[:400]); continue
    def graphed():
        t, w = s.sample(B, dev); st_t.copy_(t); st_w.copy_(w); st_eps.normal_()
        g.replay(); opt.step()
    ms_g = timeit(graphed)
    # same gradients from a replay as from an eager pass on the same static inputs?
    g.replay(); ga = net._gflat.clone()
    opt.zero_grad(); fwd_bwd(); gb = net._gflat.clone()
    err = ((ga - gb).norm() / gb.norm()).item()
    print(f"DiT-S/2 B={B}: eager {ms_e:.2f} ms ({B/ms_e*1e3:.0f} img/s), graph replay + optimizer {ms_g:.2f} ms ({B/ms_g*1e3:.0f} img/s); "
          f"{nl} library launches in fwd+bwd; replay-vs-eager grad rel diff {err:.2e}", flush=True)
    del net, opt, g; torch.cuda.empty_cache()
