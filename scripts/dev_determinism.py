"""Run-to-run determinism of forward + backward on fixed inputs: the flat gradient must be bit-identical."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
import torch
from vaw_b200.models.dit import DiT_S, DiT_XL
from vaw_b200.tools import gaussian_diffusion as gd
from gpu_util import dezero
dev = torch.device("cuda", 0)
d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda")
for name, mk, Bs in (("DiT-S", DiT_S, (64, 256, 96)), ("DiT-XL", DiT_XL, (64,))):
    for B in Bs:
        torch.manual_seed(0)
        net = mk(image_size=32, patch_size=2, in_channels=4, class_dropout_prob=0.0, num_classes=1000, learn_sigma=False).to(dev)
        dezero(net)
        x = torch.randn(B, 4, 32, 32, device=dev); y = torch.randint(0, 1000, (B,), device=dev)
        t = torch.randint(0, 1000, (B,), device=dev); eps = torch.randn_like(x)
        outs, grads = [], []
        for r in range(3):
            for p in net.parameters(): p.grad = None
            terms = d.training_losses(net, x, None, t=t, model_kwargs={"y": y}, noise=eps)
            terms["loss"].mean().backward()
            torch.cuda.synchronize()
            outs.append(terms["mse"].detach().clone()); grads.append(net._gflat.clone())
        same_out = all(torch.equal(outs[0], o) for o in outs[1:])
        same_g = all(torch.equal(grads[0], g) for g in grads[1:])
        print(f"{name} B={B}: loss identical {same_out}, grads identical {same_g}", flush=True)
        if not same_g:
            for k, p in net.named_parameters():
                if p.grad is None: continue
                off = dict((id(q), o) for q, o in net._slot_cache)[id(p)]
                a, b = grads[0][off:off + p.numel()], grads[1][off:off + p.numel()]
                if not torch.equal(a, b):
                    print(f"   {k:45s} rel diff {((a - b).norm() / (b.norm() + 1e-30)).item():.2e}")
        del net; torch.cuda.empty_cache()

# U-ViT-M/4 (config 4) and the REPA student (config 5 geometry at a small batch)
from vaw_b200.models.uvit import UViT_M
torch.manual_seed(0)
net = UViT_M(image_size=64, patch_size=4, in_channels=3, num_classes=1000, class_dropout_prob=0.0).to(dev)
B = 64
x = torch.randn(B, 3, 64, 64, device=dev).clamp(-1, 1); y = torch.randint(0, 1000, (B,), device=dev)
t = torch.randint(0, 1000, (B,), device=dev); eps = torch.randn_like(x)
res = []
for r in range(3):
    for p in net.parameters(): p.grad = None
    terms = d.training_losses(net, x, None, t=t, model_kwargs={"y": y}, noise=eps)
    terms["loss"].mean().backward(); torch.cuda.synchronize()
    res.append((terms["mse"].detach().clone(), torch.cat([p.grad.flatten() for p in net.parameters() if p.grad is not None]).clone()))
print(f"U-ViT-M B={B}: loss identical {all(torch.equal(res[0][0], a) for a, _ in res[1:])}, grads identical {all(torch.equal(res[0][1], g) for _, g in res[1:])}", flush=True)
del net; torch.cuda.empty_cache()
net = DiT_XL(image_size=32, patch_size=2, in_channels=4, class_dropout_prob=0.0, num_classes=1000, learn_sigma=False,
             learn_align=True, encoder_depth=8, z_dims=768, projector_dim=2048).to(dev)
dezero(net)
d5 = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda", learn_align=True, gamma=0.5)
B = 32
x = torch.randn(B, 4, 32, 32, device=dev); y = torch.randint(0, 1000, (B,), device=dev)
t = torch.randint(0, 1000, (B,), device=dev); eps = torch.randn_like(x); feats = torch.randn(B, 256, 768, device=dev)
res = []
for r in range(3):
    for p in net.parameters(): p.grad = None
    terms = d5.training_losses(net, x, feats, t=t, model_kwargs={"y": y}, noise=eps)
    terms["loss"].mean().backward(); torch.cuda.synchronize()
    res.append((terms["loss"].detach().clone(), net._gflat.clone()))
print(f"DiT-XL + REPA B={B}: loss identical {all(torch.equal(res[0][0], a) for a, _ in res[1:])}, grads identical {all(torch.equal(res[0][1], g) for _, g in res[1:])}", flush=True)
