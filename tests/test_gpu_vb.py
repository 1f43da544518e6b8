"""GPU tier: the variational-bound term (vaw_vb_terms, learned variance / KL objectives; reference
tools/gaussian_diffusion.py:775-808, :862-906, tools/losses.py) through training_losses, against
  * the fixture produced by executing the reference (vb_golden.npz) - forward terms and the autograd gradient w.r.t. the
    2C-channel model output,
  * the oracle (oracle/vb.py, autograd) at the latent shape with a bf16 model output.
Tolerances: the KL samples (t > 0) 1e-5-level fp32 agreement; the decoder-NLL samples (t == 0) live where fp32 tanh
saturates (cdf = 1.0 exactly, 1e-12 clamps) and the device's tanhf differs from the host's in the last bit, so single
elements can jump between log(1e-12) and log(6e-8): those samples are compared at 2e-2."""
import os

import numpy as np
import pytest
import torch

from gpu_util import relerr
from oracle import diffusion as odiff
from oracle import vb as ovb
from vaw_b200.tools import gaussian_diffusion as gd

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "vb_golden.npz"))
CASES = [("EPSILON", "LEARNED_RANGE", "MSE", "lambda"), ("EPSILON", "LEARNED", "RESCALED_MSE", "min_snr_5"),
         ("START_X", "LEARNED_RANGE", "KL", "constant"), ("EPSILON", "LEARNED_RANGE", "RESCALED_KL", "constant"),
         ("PREVIOUS_X", "LEARNED", "KL", "constant"), ("START_X", "LEARNED_RANGE", "RESCALED_MSE", "lambda")]


def _diffusion(sched, mean, var, loss, wt):
    return gd.create_gaussian_diffusion(noise_schedule=sched, mean_type=mean.lower(), var_type=var.lower(),
                                        loss_type=loss.lower(), weight_type=wt, learn_sigma=True)


@pytest.mark.parametrize("sched", ("linear", "cosine"))
@pytest.mark.parametrize("mean,var,loss,wt", CASES)
def test_vb_training_losses_vs_reference_golden(sched, mean, var, loss, wt):
    x0, eps, t = (torch.from_numpy(G[k]).to(DEV) for k in ("x0", "eps", "t"))
    mo = torch.from_numpy(G["model_out"]).to(DEV).requires_grad_(True)
    d = _diffusion(sched, mean, var, loss, wt)
    terms = d.training_losses(lambda x, ts, **k: mo, x0, None, t=t, noise=eps)
    terms["loss"].mean().backward()
    key = f"{sched}::{mean}::{var}::{loss}::{wt}"
    kl = G["t"] != 0
    for k in ("mse", "vb", "loss"):
        if f"{k}::{key}" not in G.files:
            assert k not in terms
            continue
        got, want = terms[k].detach().cpu().numpy(), G[f"{k}::{key}"]
        np.testing.assert_allclose(got[kl], want[kl], rtol=2e-5, atol=1e-6, err_msg=k)
        np.testing.assert_allclose(got[~kl], want[~kl], rtol=2e-2, err_msg=k + " (decoder NLL)")
    want = G[f"grad::{key}"]
    got = mo.grad.cpu().numpy()
    assert np.abs(got[kl] - want[kl]).max() <= 3e-5 * np.abs(want[kl]).max() + 1e-9
    assert relerr(torch.from_numpy(got[~kl]), torch.from_numpy(want[~kl])) < 5e-2


@pytest.mark.parametrize("dtype", (torch.float32, torch.bfloat16))
@pytest.mark.parametrize("mean,var,loss", [("EPSILON", "LEARNED_RANGE", "MSE"), ("EPSILON", "LEARNED", "KL"),
                                           ("START_X", "LEARNED_RANGE", "RESCALED_MSE"), ("VELOCITY", "LEARNED_RANGE", "KL")])
def test_vb_vs_oracle_at_latent_shape(dtype, mean, var, loss):
    """N = 16 latents [4, 32, 32]; a model output near the truth so that the t == 0 samples are well conditioned (a trained
    model's regime), bf16 outputs with the reference's bf16 intermediates."""
    torch.manual_seed(7)
    N, C, H = 16, 4, 32
    tb = odiff.tables(odiff.named_beta_schedule("cosine", 1000))
    x0 = torch.randn(N, C, H, H).clamp(-1, 1)
    x0[:, :, 0, :8] = -1.0
    x0[:, :, 1, :8] = 1.0
    eps = torch.randn(N, C, H, H)
    t = torch.randint(0, 1000, (N,))
    t[:4] = 0
    x_t = torch.from_numpy(odiff.q_sample(tb, x0.numpy(), t.numpy(), eps.numpy()))
    truth = {"EPSILON": eps, "START_X": x0, "VELOCITY": torch.from_numpy(odiff.target(tb, "VELOCITY", x0.numpy(), t.numpy(), eps.numpy()))}[mean]
    mo = torch.cat([truth + 0.05 * torch.randn(N, C, H, H), 0.4 * torch.randn(N, C, H, H) + (0.9 if var == "LEARNED_RANGE" else -4.0)], dim=1)
    mo = mo.to(dtype).float()          # the values the kernel sees; the oracle widens the same numbers
    d = _diffusion("cosine", mean, var, loss, "lambda" if loss != "KL" else "constant")
    out = mo.to(DEV).to(dtype).requires_grad_(True)
    terms = d.training_losses(lambda x, ts, **k: out, x0.to(DEV), None, t=t.to(DEV), noise=eps.to(DEV))
    terms["loss"].mean().backward()
    ref_in = mo.clone().requires_grad_(True)
    ref = ovb.training_losses(tb, mean, var, loss, "lambda" if loss != "KL" else "constant", ref_in, x0.numpy(), t.numpy(),
                              eps.numpy(), bf16_out=dtype == torch.bfloat16)
    ref["loss"].mean().backward()
    kl = (t != 0).numpy()
    for k in ref:
        got, want = terms[k].detach().cpu().numpy(), ref[k].detach().numpy()
        np.testing.assert_allclose(got[kl], want[kl], rtol=3e-5, atol=1e-6, err_msg=k)
        np.testing.assert_allclose(got[~kl], want[~kl], rtol=2e-3, err_msg=k + " (decoder NLL)")
    assert out.grad.dtype == dtype
    tol = 1e-2 if dtype == torch.bfloat16 else 2e-4
    assert relerr(out.grad.float().cpu()[kl], ref_in.grad[kl]) < tol
    assert relerr(out.grad.float().cpu()[~kl], ref_in.grad[~kl]) < max(tol, 2e-3)
    assert float(out.grad[:, C:].abs().sum()) > 0          # the variance channels receive a gradient


def test_vb_argument_errors_and_fixed_variance_kl():
    d = gd.create_gaussian_diffusion(noise_schedule="linear", mean_type="epsilon", var_type="fixed_small", loss_type="kl",
                                     weight_type="constant")
    tb = odiff.tables(odiff.named_beta_schedule("linear", 1000))
    torch.manual_seed(1)
    x0 = torch.randn(5, 3, 8, 8).clamp(-1, 1); eps = torch.randn(5, 3, 8, 8); t = torch.tensor([3, 100, 500, 900, 999])
    out = torch.randn(5, 3, 8, 8, device=DEV, requires_grad=True)
    terms = d.training_losses(lambda x, ts, **k: out, x0.to(DEV), None, t=t.to(DEV), noise=eps.to(DEV))
    assert set(terms) == {"loss"}
    terms["loss"].sum().backward()
    # KL between two Gaussians of equal (fixed-small) variance: 0.5 (m1 - m2)^2 / var, in bits
    x_t = odiff.q_sample(tb, x0.numpy(), t.numpy(), eps.numpy())
    pm = odiff.p_mean_variance(tb, "EPSILON", "FIXED_SMALL", out.detach().cpu().numpy(), x_t, t.numpy(), clip_denoised=False)
    c1, c2 = odiff.extract(tb["posterior_mean_coef1"], t.numpy()), odiff.extract(tb["posterior_mean_coef2"], t.numpy())
    tm = c1 * x0.numpy() + c2 * x_t
    want = (0.5 * (tm - pm["mean"]) ** 2 / pm["variance"]).reshape(5, -1).mean(1) / np.log(2.0)
    np.testing.assert_allclose(terms["loss"].detach().cpu().numpy(), want, rtol=2e-4)
    assert torch.isfinite(out.grad).all() and float(out.grad.abs().sum()) > 0
    with pytest.raises(AssertionError):   # a learned variance needs 2C output channels (:890)
        gd.create_gaussian_diffusion(var_type="learned_range", learn_sigma=True).training_losses(
            lambda x, ts, **k: out, x0.to(DEV), None, t=t.to(DEV), noise=eps.to(DEV))
