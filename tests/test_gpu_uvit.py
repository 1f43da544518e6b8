"""GPU tier: the U-ViT engine vs the oracle (bf16 autocast, 2e-2 rel-L2) and vs the reference's golden fixture."""
import os

import numpy as np
import pytest
import torch

from gpu_util import relerr
from oracle.uvit import uvit_forward
from vaw_b200.models.uvit import UViT
from vaw_b200.tools import gaussian_diffusion as gd

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = os.path.join(os.path.dirname(__file__), "golden")
TOL = 2e-2


# The last row is U-ViT-M's width and sequence (BASELINE.json config 4): D = 768, 12 heads of 64, 64x64x3 pixels at patch 4
# -> T = 256 + 2 extra tokens = 258, depth 3 (one in-block, the mid block, one out-block with its skip_linear).
@pytest.mark.parametrize("dim,heads,depth,img,patch,classes,B", [(128, 2, 3, 16, 2, 10, 4), (128, 2, 5, 16, 4, -1, 3),
                                                                 (384, 6, 5, 32, 4, 100, 8), (768, 12, 3, 64, 4, 1000, 4)])
def test_forward_backward_vs_oracle(dim, heads, depth, img, patch, classes, B):
    torch.manual_seed(1)
    m = UViT(image_size=img, patch_size=patch, in_channels=3 if patch == 4 else 4, embed_dim=dim, depth=depth,
             num_heads=heads, num_classes=classes).to(DEV).train()
    C = m.in_channels
    x = torch.randn(B, C, img, img, device=DEV); t = torch.rand(B, device=DEV) * 999
    y = torch.randint(0, classes, (B,), device=DEV) if classes > 0 else None
    gout = torch.randn(B, C, img, img, device=DEV)
    out = m(x, t, y)
    assert out.shape == x.shape and out.dtype == torch.float32
    (out * gout).sum().backward()
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict(keep_vars=True).items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o_ref = uvit_forward(sd, x, t, y, patch_size=patch, num_heads=heads, depth=depth)
    (o_ref.float() * gout).sum().backward()
    sd32 = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict(keep_vars=True).items()}
    o32 = uvit_forward(sd32, x, t, y, patch_size=patch, num_heads=heads, depth=depth)
    (o32 * gout).sum().backward()
    assert relerr(out, o_ref) < TOL
    for k, p in m.named_parameters():
        # compare against the fp32 oracle with the tolerance the bf16 oracle itself needs (+ slack), and directly
        e_engine, e_oracle = relerr(p.grad, sd32[k].grad), relerr(sd[k].grad, sd32[k].grad)
        assert e_engine < max(TOL, 1.5 * e_oracle), (k, e_engine, e_oracle)


def test_reference_golden_weights():
    g = np.load(os.path.join(G, "uvit_golden.npz"))
    m = UViT(image_size=8, patch_size=2, in_channels=4, embed_dim=64, depth=3, num_heads=1, mlp_ratio=4, num_classes=10)
    m.load_state_dict({k[len("param::"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("param::")})
    m = m.to(DEV).train()
    x0, eps, t, y = (torch.from_numpy(g[k]).to(DEV) for k in ("x0", "eps", "t", "y"))
    d = gd.create_gaussian_diffusion(noise_schedule="linear", mean_type="epsilon", weight_type="lambda")
    out = m(d.q_sample(x0, t, eps), d._scale_timesteps(t), y)
    assert relerr(out, torch.from_numpy(g["fwd_out"]).to(DEV)) < TOL
    terms = d.training_losses(m, x0, None, t=t, model_kwargs={"y": y}, noise=eps)
    terms["loss"].mean().backward()
    np.testing.assert_allclose(terms["mse"].detach().cpu().numpy(), g["mse"], rtol=TOL)
    n = 0
    for k, p in m.named_parameters():
        assert relerr(p.grad, torch.from_numpy(g["grad::" + k]).to(DEV)) < 3e-2, k
        n += 1
    assert n > 40
