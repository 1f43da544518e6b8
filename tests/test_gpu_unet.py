"""GPU tier: BASELINE config 1 style step — the UNet (PyTorch module, §8 a17) driven by the B200 diffusion kernels
(K1 q_sample/target, K2 weighted MSE + gradient) in fp32 — against the fixture produced by the reference on CPU."""
import os
import sys

import numpy as np
import pytest
import torch

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, G)
from fill import fill_by_name, grad_digest  # noqa: E402

from test_unet_cpu import CASES  # noqa: E402
from vaw_b200 import _lib  # noqa: E402
from vaw_b200.models import unet as vunet  # noqa: E402
from vaw_b200.tools import gaussian_diffusion as gd  # noqa: E402
from vaw_b200.tools import resample as rs  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _fp32_math():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("tag", ["a", "b"])
def test_unet_training_losses_vs_reference_golden(tag):
    g = np.load(os.path.join(G, "unet_golden.npz"))
    cfg = CASES[tag]
    m = fill_by_name(vunet.create_unet_model(**cfg)).to(DEV).train()
    x0, eps, t, y = (torch.from_numpy(g[f"{tag}::{k}"]).to(DEV) for k in ("x0", "eps", "t", "y"))
    kw = {"y": y} if cfg["class_cond"] else {}
    d = gd.create_gaussian_diffusion(noise_schedule="linear", mean_type="epsilon", weight_type="min_snr_5")
    n0 = _lib.launch_count()
    terms = d.training_losses(m, x0, None, t=t, model_kwargs=kw, noise=eps)
    terms["loss"].mean().backward()
    assert _lib.launch_count() > n0          # the diffusion maths ran in the CUDA library, not in torch
    np.testing.assert_allclose(terms["mse"].detach().cpu().numpy(), g[f"{tag}::mse"], rtol=2e-5)
    for k, p in m.named_parameters():
        want = g[f"{tag}::grad::{k}"]
        got = grad_digest(p.grad.cpu())
        scale = max(float(np.abs(want).max()), 1e-6)
        assert np.abs(got - want).max() <= 5e-4 * scale + 1e-6, (k, np.abs(got - want).max(), scale)


def test_unet32_config1_step_with_loss_aware_sampler():
    """configs[0]: UNet-32 on 3x32x32, loss-second-moment sampler, a few optimiser steps; loss finite and decreasing
    on a fixed batch, sampler history filled through the device path."""
    torch.manual_seed(0)
    np.random.seed(0)
    m = vunet.UNet_32(num_classes=10).to(DEV).train()
    d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda")
    s = rs.LossSecondMomentResampler(d)
    opt = torch.optim.AdamW(m.parameters(), lr=2e-4)
    x0 = torch.randn(16, 3, 32, 32, device=DEV).clamp(-1, 1)
    y = torch.randint(0, 10, (16,), device=DEV)
    losses = []
    for _ in range(6):
        t, w = s.sample(16, DEV)
        terms = d.training_losses(m, x0, None, t=t, model_kwargs={"y": y})
        s.update_with_local_losses(t, terms["loss"].detach())
        loss = (terms["loss"] * w).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert all(np.isfinite(losses))
    assert int(s._loss_counts.sum()) == 6 * 16, (s._loss_counts.sum(), losses)
