"""ORACLE (test infrastructure, not product code): the variational-bound term of the training objective.

Follows /root/reference/tools/gaussian_diffusion.py:
    q_posterior_mean_variance   :254-276
    p_mean_variance             :278-384  (learned-variance parameterisations :312-324, clip_denoised=False here)
    _vb_terms_bpd               :775-808  KL(q(x_{t-1}|x_t,x_0) || p(x_{t-1}|x_t)) in bits, decoder NLL at t == 0
    training_losses             :862-875  (LossType.KL / RESCALED_KL) and :886-906 (MSE + vb with the mean detached,
                                           RESCALED_MSE scales vb by T / 1000), :921-922 loss = mse + vb
and /root/reference/tools/losses.py:12-77 (normal_kl, approx_standard_normal_cdf, discretized_gaussian_log_likelihood).

torch float32 on the CPU, the reference's operation order (the decoder NLL at t = 0 lives where fp32 tanh saturates:
cdf values of exactly 1.0 and the 1e-12 clamps are part of the reference's result, so float64 would be a DIFFERENT
function there), with autograd supplying the gradients: the CUDA kernel's hand-derived gradient is checked against an
independent derivation.  Pinned against the executed reference by tests/golden/make_golden.py ->
tests/golden/vb_golden.npz (tests/test_oracle_vb.py).  Only tests/ may import this module.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import diffusion as odiff


DT = torch.float32


def _gather(table, t, like):
    v = torch.from_numpy(np.asarray(table, dtype=np.float64)[np.asarray(t)].astype(np.float32)).to(DT)
    return v.reshape(-1, *([1] * (like.dim() - 1)))


def normal_kl(mean1, logvar1, mean2, logvar2):
    return 0.5 * (-1.0 + logvar2 - logvar1 + torch.exp(logvar1 - logvar2) + (mean1 - mean2) ** 2 * torch.exp(-logvar2))


def _cdf(x):
    return 0.5 * (1.0 + torch.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * torch.pow(x, 3))))


def discretized_gaussian_log_likelihood(x, means, log_scales):
    centered = x - means
    inv_stdv = torch.exp(-log_scales)
    cdf_plus = _cdf(inv_stdv * (centered + 1.0 / 255.0))
    cdf_min = _cdf(inv_stdv * (centered - 1.0 / 255.0))
    log_cdf_plus = torch.log(cdf_plus.clamp(min=1e-12))
    log_one_minus_cdf_min = torch.log((1.0 - cdf_min).clamp(min=1e-12))
    delta = cdf_plus - cdf_min
    return torch.where(x < -0.999, log_cdf_plus,
                       torch.where(x > 0.999, log_one_minus_cdf_min, torch.log(delta.clamp(min=1e-12))))


def vb_terms(tb, mean_type, var_type, model_output, x_start, x_t, t, bf16_out=False):
    """_vb_terms_bpd with clip_denoised=False.  model_output: torch float32 [N, 2C, ...] (requires_grad for gradients).
    Returns the per-sample bound in bits [N]."""
    x0, xt = torch.as_tensor(x_start).to(DT), torch.as_tensor(x_t).to(DT)
    C = x0.shape[1]
    o, v = model_output[:, :C], model_output[:, C:]
    c1, c2 = _gather(tb["posterior_mean_coef1"], t, x0), _gather(tb["posterior_mean_coef2"], t, x0)
    true_mean = c1 * x0 + c2 * xt
    true_lv = _gather(tb["posterior_log_variance_clipped"], t, x0)
    if var_type == "LEARNED":
        lv = v
    elif var_type == "LEARNED_RANGE":
        frac = (v + 1) / 2
        max_log = _gather(np.log(tb["betas"]), t, x0)
        lv = frac * max_log + (1 - frac) * true_lv
        if bf16_out:
            # the reference evaluates (v + 1) / 2 and 1 - frac in the model's output dtype (bf16 under autocast) before
            # the products with the fp32 tables promote to fp32; the roundings are transparent to the gradient
            rb, f = odiff._bf16_round, np.float32
            fr = rb(rb(v.detach().float().numpy() + f(1)) / f(2))
            lv_r = (torch.from_numpy(fr).to(DT) * max_log + torch.from_numpy(rb(f(1) - fr)).to(DT) * true_lv)
            lv = lv + (lv_r - lv.detach())
    else:
        raise NotImplementedError(var_type)
    if mean_type == "PREVIOUS_X":
        mean = o
    else:
        if mean_type == "START_X":
            xs = o
        elif mean_type == "EPSILON":
            xs = _gather(tb["sqrt_recip_alphas_cumprod"], t, x0) * xt - _gather(tb["sqrt_recipm1_alphas_cumprod"], t, x0) * o
        elif mean_type == "VELOCITY":
            xs = _gather(tb["sqrt_alphas_cumprod"], t, x0) * xt - _gather(tb["sqrt_one_minus_alphas_cumprod"], t, x0) * o
        else:
            raise NotImplementedError(mean_type)
        mean = c1 * xs + c2 * xt
    dims = list(range(1, x0.dim()))
    kl = normal_kl(true_mean, true_lv, mean, lv).mean(dim=dims) / np.log(2.0)
    nll = (-discretized_gaussian_log_likelihood(x0, mean, 0.5 * lv)).mean(dim=dims) / np.log(2.0)
    return torch.where(torch.as_tensor(np.asarray(t)) == 0, nll, kl)


def training_losses(tb, mean_type, var_type, loss_type, weight_type, model_output, x_start, t, noise, T=1000,
                    bf16_out=False):
    """training_losses (:834-930) for the learned-variance configurations.  model_output: torch float32 leaf
    [N, 2C, ...].  -> dict(mse=?, vb=?, loss=) of torch tensors (graph attached)."""
    x0, eps = np.asarray(x_start, dtype=np.float32), np.asarray(noise, dtype=np.float32)
    x_t = odiff.q_sample(tb, x0, t, eps)
    C = x0.shape[1]
    if loss_type in ("KL", "RESCALED_KL"):
        out = vb_terms(tb, mean_type, var_type, model_output, x0, x_t, t, bf16_out)
        return {"loss": out * (T if loss_type == "RESCALED_KL" else 1.0)}
    frozen = torch.cat([model_output[:, :C].detach(), model_output[:, C:]], dim=1)
    vb = vb_terms(tb, mean_type, var_type, frozen, x0, x_t, t, bf16_out)
    if loss_type == "RESCALED_MSE":
        vb = vb * (T / 1000.0)
    a = odiff.extract(tb["sqrt_alphas_cumprod"], t).reshape(-1)
    s = odiff.extract(tb["sqrt_one_minus_alphas_cumprod"], t).reshape(-1)
    w = torch.from_numpy(odiff.loss_weight(mean_type, weight_type, a, s)).to(DT)
    tg = torch.from_numpy(np.ascontiguousarray(odiff.target(tb, mean_type, x0, t, eps))).to(DT)
    mse = w * ((tg - model_output[:, :C]) ** 2).mean(dim=list(range(1, tg.dim())))
    return {"mse": mse, "vb": vb, "loss": mse + vb}
