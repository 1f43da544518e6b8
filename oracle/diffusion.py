"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's diffusion training objective.

Follows /root/reference/tools/gaussian_diffusion.py:
    schedules          get_named_beta_schedule :59-104, betas_for_alpha_bar :107-123
    tables             GaussianDiffusion.__init__ :166-205
    gather             _extract_into_tensor :1059-1072   (float64 table -> gather -> ONE rounding to float32)
    q_sample           :234-252      x_t = fl(fl(a*x0) + fl(s*eps))
    compute_target     :818-832
    loss weight        compute_mse_loss_weight :1092-1148 (table in SURVEY.md §A.1)
    training_losses    :834-930 (MSE branch, fixed variance) and FlowMatching.training_losses :1297-1340
    align loss         compute_align_loss :1007-1046
and tools/nn.py:86-90 (mean_flat).

numpy for the integer/fp32 elementwise arithmetic (bit-exact with torch's eager fp32 ops), torch only where a model
has to be differentiated.  Pinned against the executed reference by tests/golden/make_golden.py -> tests/golden/*.npz
(tests/test_oracle_golden.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
from __future__ import annotations

import math

import numpy as np

MEAN_TYPES = ("PREVIOUS_X", "START_X", "EPSILON", "VELOCITY", "VECTOR", "SCORE")


def named_beta_schedule(name, T, lambda_max=10.0, lambda_min=-10.0, max_beta=0.999):
    if name == "linear":
        scale = 1000 / T
        return np.linspace(scale * 0.0001, scale * 0.02, T, dtype=np.float64)
    if name == "cosine":
        fn = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    elif name == "linear_logsnr":
        fn = lambda t: 1.0 / (1.0 + math.exp(-(lambda_max + t * (lambda_min - lambda_max))))
    else:
        raise NotImplementedError(name)
    return np.array([min(1 - fn((i + 1) / T) / fn(i / T), max_beta) for i in range(T)], dtype=np.float64)


def tables(betas):
    """The float64 per-timestep tables of GaussianDiffusion.__init__."""
    betas = np.asarray(betas, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    return dict(
        betas=betas, alphas_cumprod=ac, alphas_cumprod_prev=ac_prev,
        sqrt_alphas_cumprod=np.sqrt(ac), sqrt_one_minus_alphas_cumprod=np.sqrt(1.0 - ac),
        posterior_mean_coef1=betas * np.sqrt(ac_prev) / (1.0 - ac),
        posterior_mean_coef2=(1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    )


def extract(table, t):
    """float64 table gathered at integer t, rounded once to float32, shaped [N,1,1,1]."""
    return np.asarray(table, dtype=np.float64)[np.asarray(t)].astype(np.float32).reshape(-1, 1, 1, 1)


def q_sample(tb, x0, t, eps):
    a, s = extract(tb["sqrt_alphas_cumprod"], t), extract(tb["sqrt_one_minus_alphas_cumprod"], t)
    x0 = np.asarray(x0, dtype=np.float32)
    eps = np.asarray(eps, dtype=np.float32)
    return (a * x0).astype(np.float32) + (s * eps).astype(np.float32)


def target(tb, mean_type, x0, t, eps):
    x0 = np.asarray(x0, dtype=np.float32)
    eps = np.asarray(eps, dtype=np.float32)
    if mean_type == "START_X":
        return x0
    if mean_type == "EPSILON":
        return eps
    a, s = extract(tb["sqrt_alphas_cumprod"], t), extract(tb["sqrt_one_minus_alphas_cumprod"], t)
    if mean_type == "VELOCITY":
        return (a * eps).astype(np.float32) - (s * x0).astype(np.float32)
    if mean_type == "PREVIOUS_X":
        c1, c2 = extract(tb["posterior_mean_coef1"], t), extract(tb["posterior_mean_coef2"], t)
        return (c1 * x0).astype(np.float32) + (c2 * q_sample(tb, x0, t, eps)).astype(np.float32)
    raise NotImplementedError(mean_type)


def loss_weight(mean_type, weight_type, alpha, sigma, p2_k=1.0, p2_gamma=1.0):
    """compute_mse_loss_weight on float32 vectors alpha, sigma (per sample or per timestep); fp32 op order of the
    reference.  Raises ValueError for the combinations the reference rejects."""
    alpha = np.asarray(alpha, dtype=np.float32)
    sigma = np.asarray(sigma, dtype=np.float32)
    one = np.float32(1.0)
    if weight_type == "constant":
        return np.ones_like(alpha)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = (alpha / sigma).astype(np.float32)
        snr = (q * q).astype(np.float32)
        w = None
        def kval(prefix):
            return np.float32(float(weight_type.split(prefix)[-1]))
        if mean_type == "EPSILON":
            if weight_type.startswith("min_snr_"):
                w = np.minimum(snr, kval("min_snr_")) / snr
            elif weight_type.startswith("max_snr_"):
                w = np.maximum(snr, kval("max_snr_")) / snr
            elif weight_type == "lambda":
                w = sigma.copy()
            elif weight_type == "debias":
                w = sigma / alpha
            elif weight_type == "p2":
                w = one / np.power((np.float32(p2_k) + snr).astype(np.float32), np.float32(p2_gamma))
            elif weight_type == "min_debias":
                w = np.minimum(sigma / alpha, one)
            elif weight_type == "max_debias":
                w = np.maximum(sigma / alpha, one)
        elif mean_type == "START_X":
            if weight_type == "trunc_snr":
                w = np.maximum(snr, one)
            elif weight_type == "snr":
                w = snr.copy()
            elif weight_type == "inv_snr":
                w = one / snr
            elif weight_type.startswith("min_snr_"):
                w = np.minimum(snr, kval("min_snr_"))
            elif weight_type.startswith("max_snr_"):
                w = np.maximum(snr, kval("max_snr_"))
            elif weight_type == "lambda":
                w = alpha.copy()
        elif mean_type == "VECTOR":
            if weight_type == "lambda":
                w = np.ones_like(alpha)
        elif mean_type == "VELOCITY":
            if weight_type.startswith("min_snr_"):
                w = np.minimum(snr, kval("min_snr_")) / (snr + one)
            elif weight_type == "lambda":
                w = alpha * sigma
    if w is None:
        raise ValueError(f"Invalid mse_loss_weight_type: {weight_type}")
    w = np.asarray(w, dtype=np.float32)
    w[snr == 0] = 1.0
    return w


def weight_lut(tb, mean_type, weight_type, p2_k=1.0, p2_gamma=1.0):
    a = tb["sqrt_alphas_cumprod"].astype(np.float32)
    s = tb["sqrt_one_minus_alphas_cumprod"].astype(np.float32)
    return loss_weight(mean_type, weight_type, a, s, p2_k, p2_gamma)


def sample_from_latent(latent, eps, latent_scale=1.0):
    """tools/trainer.py:21-25 with the draw made explicit: latent [N, 2C, ...] = (mean | std) -> (mean + std*eps)*scale,
    every fp32 operation rounded separately (torch elementwise order)."""
    latent = np.asarray(latent, dtype=np.float32)
    c = latent.shape[1] // 2
    mean, std = latent[:, :c], latent[:, c:]
    return ((mean + std * np.asarray(eps, dtype=np.float32)).astype(np.float32) * np.float32(latent_scale)).astype(np.float32)


def mean_flat(x):
    return x.reshape(x.shape[0], -1).mean(axis=1, dtype=np.float32)


def mse_terms(tb, mean_type, weight_type, x0, t, eps, model_output, p2_k=1.0, p2_gamma=1.0):
    """(mse [N], d mse_n / d model_output [N,...]) in float64-accumulated numpy (reference value up to fp32
    summation order): mse_n = w_n * mean((target - out)^2)."""
    tg = target(tb, mean_type, x0, t, eps).astype(np.float64)
    out = np.asarray(model_output, dtype=np.float64)
    a = extract(tb["sqrt_alphas_cumprod"], t).reshape(-1)
    s = extract(tb["sqrt_one_minus_alphas_cumprod"], t).reshape(-1)
    w = loss_weight(mean_type, weight_type, a, s, p2_k, p2_gamma).astype(np.float64)
    d = tg - out
    chw = d[0].size
    mse = w * (d.reshape(d.shape[0], -1) ** 2).mean(axis=1)
    grad = (-2.0 / chw) * w.reshape(-1, *([1] * (d.ndim - 1))) * d
    return mse, grad


# ----------------------------------------------------------------------------------------------------------
# torch versions (differentiable through a model) for the CPU training step used as the reported CPU baseline
# ----------------------------------------------------------------------------------------------------------
def training_losses_torch(tb, mean_type, weight_type, model_fn, x0, t, eps, *, num_timesteps=1000, rescale=True,
                          features=None, gamma=0.5, learn_align=False):
    """Reference :834-930 with torch tensors on CPU (MSE branch).  model_fn(x_t, t_scaled) -> out or (out, zs)."""
    import torch
    x_t = torch.from_numpy(q_sample(tb, x0.numpy(), t.numpy(), eps.numpy()))
    a = torch.from_numpy(extract(tb["sqrt_alphas_cumprod"], t.numpy()).reshape(-1))
    s = torch.from_numpy(extract(tb["sqrt_one_minus_alphas_cumprod"], t.numpy()).reshape(-1))
    w = torch.from_numpy(loss_weight(mean_type, weight_type, a.numpy(), s.numpy()))
    ts = t.float() * (1000.0 / num_timesteps) if rescale else t
    raw = model_fn(x_t, ts)
    out, zs = (raw[0], raw[1]) if isinstance(raw, tuple) else (raw, None)
    tg = torch.from_numpy(np.ascontiguousarray(target(tb, mean_type, x0.numpy(), t.numpy(), eps.numpy())))
    raw_mse = ((tg - out) ** 2).mean(dim=list(range(1, out.dim())))
    terms = {"mse": w * raw_mse}
    if learn_align:
        terms["align"] = torch.nn.functional.mse_loss(zs, features)
        terms["loss"] = terms["mse"] + gamma * terms["align"]
    else:
        terms["loss"] = terms["mse"]
    return terms
