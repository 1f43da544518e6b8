"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's EDM-style sampler.

Follows /root/reference/tools/cfg_edm.py:
    Net.__init__ / alpha_bar / round_sigma   :42-47, :83-107   sigma table u (float32), nearest-entry rounding
    Net.forward                              :51-79            c_in, c_noise, c_skip / c_out per prediction type
    ablation_sampler                         :109-210          time grid (vp / ve / iddpm / edm), sigma(t) and s(t)
                                                               families, churn, Euler / Heun (alpha-generalised) steps
torch on the CPU: float64 state and time scalars, float32 preconditioning, the operations in the reference's order (the
whole computation is deterministic IEEE arithmetic plus libm pow / log / sin on scalars, so the result is bit-identical
to the executed reference for the same denoiser).  Pinned by tests/golden/make_golden.py -> tests/golden/edm_golden.npz
(tests/test_oracle_edm.py).  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np
import torch


class SigmaTable:
    """The sigma levels of the M-step DDPM schedule (Net.__init__ :42-47) and the lookups on them."""

    def __init__(self, noise_schedule="linear", M=1000, C_1=0.001, C_2=0.008, lambda_max=10.0, lambda_min=-10.0):
        self.M, self.C_2, self.schedule, self.lmax, self.lmin = M, C_2, noise_schedule, lambda_max, lambda_min
        u = torch.zeros(M + 1)
        for j in range(M, 0, -1):
            r = self.alpha_bar(j - 1) / self.alpha_bar(j)
            u[j - 1] = ((u[j] ** 2 + 1) / r.clip(min=C_1) - 1).sqrt()
        self.u = u
        self.sigma_min, self.sigma_max = float(u[M - 1]), float(u[0])

    def alpha_bar(self, j):
        j = torch.as_tensor(j)
        if self.schedule == "cosine":
            return (0.5 * np.pi * j / self.M / (self.C_2 + 1)).sin() ** 2
        if self.schedule == "linear":
            return np.cumprod(1.0 - np.linspace(0.0001, 0.02, self.M + 1, dtype=np.float64), axis=0)[self.M - j]
        if self.schedule == "linear_logsnr":
            return torch.sigmoid(self.lmax + (self.M - j) / self.M * (self.lmin - self.lmax))
        raise NotImplementedError(self.schedule)

    def round(self, sigma, return_index=False):
        sigma = torch.as_tensor(sigma)
        idx = torch.cdist(sigma.to(torch.float32).reshape(1, -1, 1), self.u.reshape(1, -1, 1)).argmin(2)
        res = idx if return_index else self.u[idx.flatten()].to(sigma.dtype)
        return res.reshape(sigma.shape)


def denoise(table, pred_type, model_fn, x, sigma, channels):
    """Net.forward (:51-79) with amp off.  x: float64/32 [N, C, H, W], sigma: 0-dim tensor.  model_fn(x_in fp32, t int32 [N])
    -> denoiser output.  Returns float32 (START_X: the model's dtype)."""
    x = x.to(torch.float32)
    sigma = sigma.to(torch.float32).reshape(-1, 1, 1, 1)
    c_noise = table.M - 1 - table.round(sigma, return_index=True).to(torch.float32)
    c_in = 1 / (sigma ** 2 + 1).sqrt()
    out = model_fn(c_in * x, c_noise.flatten().repeat(x.shape[0]).int())
    if pred_type == "EPSILON":
        return 1 * x + (-sigma) * out[:, :channels].to(torch.float32)
    if pred_type == "START_X":
        return out
    if pred_type == "VELOCITY":
        return (c_in ** 2) * x + (-sigma * c_in) * out[:, :channels].to(torch.float32)
    raise ValueError(pred_type)


def sample(table, pred_type, model_fn, latents, noises, num_steps=18, sigma_min=None, sigma_max=None, rho=7, solver="heun",
           discretization="edm", schedule="linear", scaling="none", epsilon_s=1e-3, C_1=0.001, C_2=0.008, M=1000, alpha=1,
           S_churn=0, S_min=0, S_max=float("inf"), S_noise=1):
    """ablation_sampler (:109-210); `noises` is an iterable of the randn_like draws (float64)."""
    vp_sig = lambda bd, bm: lambda t: (np.e ** (0.5 * bd * (t ** 2) + bm * t) - 1) ** 0.5
    if sigma_min is None:
        sigma_min = {"vp": vp_sig(19.9, 0.1)(epsilon_s), "ve": 0.02, "iddpm": 0.002, "edm": 0.002}[discretization]
    if sigma_max is None:
        sigma_max = {"vp": vp_sig(19.9, 0.1)(1), "ve": 100, "iddpm": 81, "edm": 80}[discretization]
    sigma_min, sigma_max = max(sigma_min, table.sigma_min), min(sigma_max, table.sigma_max)
    bd = 2 * (np.log(sigma_min ** 2 + 1) / epsilon_s - np.log(sigma_max ** 2 + 1)) / (epsilon_s - 1)
    bm = np.log(sigma_max ** 2 + 1) - 0.5 * bd
    k = torch.arange(num_steps, dtype=torch.float64)
    if discretization == "vp":
        sig_steps = vp_sig(bd, bm)(1 + k / (num_steps - 1) * (epsilon_s - 1))
    elif discretization == "ve":
        sig_steps = ((sigma_max ** 2) * ((sigma_min ** 2 / sigma_max ** 2) ** (k / (num_steps - 1)))).sqrt()
    elif discretization == "iddpm":
        u = torch.zeros(M + 1, dtype=torch.float64)
        ab = lambda j: (0.5 * np.pi * j / M / (C_2 + 1)).sin() ** 2
        for j in torch.arange(M, 0, -1):
            u[j - 1] = ((u[j] ** 2 + 1) / (ab(j - 1) / ab(j)).clip(min=C_1) - 1).sqrt()
        kept = u[torch.logical_and(u >= sigma_min, u <= sigma_max)]
        sig_steps = kept[((len(kept) - 1) / (num_steps - 1) * k).round().to(torch.int64)]
    else:
        sig_steps = (sigma_max ** (1 / rho) + k / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    if schedule == "vp":
        sig = vp_sig(bd, bm)
        dsig = lambda t: 0.5 * (bm + bd * t) * (sig(t) + 1 / sig(t))
        sig_inv = lambda s_: ((bm ** 2 + 2 * bd * (s_ ** 2 + 1).log()).sqrt() - bm) / bd
    elif schedule == "ve":
        sig, dsig, sig_inv = (lambda t: t.sqrt()), (lambda t: 0.5 / t.sqrt()), (lambda s_: s_ ** 2)
    else:
        sig, dsig, sig_inv = (lambda t: t), (lambda t: 1), (lambda s_: s_)
    if scaling == "vp":
        sc = lambda t: 1 / (1 + sig(t) ** 2).sqrt()
        dsc = lambda t: -sig(t) * dsig(t) * (sc(t) ** 3)
    else:
        sc, dsc = (lambda t: 1), (lambda t: 0)
    ts = sig_inv(table.round(sig_steps))
    ts = torch.cat([ts, torch.zeros_like(ts[:1])])
    C = latents.shape[1]
    it = iter(noises)
    den = lambda x, t: denoise(table, pred_type, model_fn, x / sc(t), sig(t), C).to(torch.float64)
    slope = lambda x, t, d: (dsig(t) / sig(t) + dsc(t) / sc(t)) * x - dsig(t) * sc(t) / sig(t) * d
    x_next = latents.to(torch.float64) * (sig(ts[0]) * sc(ts[0]))
    for i, (t_cur, t_next) in enumerate(zip(ts[:-1], ts[1:])):
        x_cur = x_next
        gamma = min(S_churn / num_steps, np.sqrt(2) - 1) if S_min <= sig(t_cur) <= S_max else 0
        t_hat = sig_inv(table.round(sig(t_cur) + gamma * sig(t_cur)))
        x_hat = sc(t_hat) / sc(t_cur) * x_cur + (sig(t_hat) ** 2 - sig(t_cur) ** 2).clip(min=0).sqrt() * sc(t_hat) * S_noise * next(it).to(torch.float64)
        h = t_next - t_hat
        d_cur = slope(x_hat, t_hat, den(x_hat, t_hat))
        if solver == "euler" or i == num_steps - 1:
            x_next = x_hat + h * d_cur
            continue
        x_prime, t_prime = x_hat + alpha * h * d_cur, t_hat + alpha * h
        d_prime = slope(x_prime, t_prime, den(x_prime, t_prime))
        x_next = x_hat + h * ((1 - 1 / (2 * alpha)) * d_cur + 1 / (2 * alpha) * d_prime)
    return x_next
