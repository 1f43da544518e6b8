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
    reverse step       p_mean_variance :278-384, p_sample :455-506, ddim_sample :603-651, ddim_reverse_sample :653-689
                       and IntervalCFG.forward tools/sampler.py:32-48 (the guidance combine)
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
        alphas_cumprod_next=np.append(ac[1:], 0.0),
        sqrt_recip_alphas_cumprod=np.sqrt(1.0 / ac), sqrt_recipm1_alphas_cumprod=np.sqrt(1.0 / ac - 1),
        posterior_variance=(pv := betas * (1.0 - ac_prev) / (1.0 - ac)),
        posterior_log_variance_clipped=np.log(np.append(pv[1], pv[1:])),
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


# ---- reverse process (inference path) ----------------------------------------------------------------
def _bf16_round(a):
    """Round-to-nearest-even to bfloat16, returned as float32 (what a bf16 torch op does to its fp32 result)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = r.view(np.float32).copy()
    nan = np.isnan(a)
    out[nan] = np.nan
    return out.reshape(np.shape(a))


def p_mean_variance(tb, mean_type, var_type, model_output, x, t, clip_denoised=True, out_is_bf16=False):
    """gaussian_diffusion.py:278-384.  model_output: [N, C or 2C, H, W] float32 values (already widened if the model
    returned bf16: pass out_is_bf16=True so the variance branch keeps the reference's bf16 intermediates)."""
    f = np.float32
    x = np.asarray(x, dtype=f)
    model_output = np.asarray(model_output, dtype=f)
    C = x.shape[1]
    rb = _bf16_round if out_is_bf16 else (lambda a: a)
    if var_type in ("LEARNED", "LEARNED_RANGE"):
        assert model_output.shape[1] == 2 * C
        model_output, v = model_output[:, :C], model_output[:, C:]
        if var_type == "LEARNED":
            log_var = v
            var = rb(np.exp(v).astype(f))
        else:
            min_log = extract(tb["posterior_log_variance_clipped"], t)
            max_log = extract(np.log(tb["betas"]), t)
            frac = rb((rb((v + f(1)).astype(f)) / f(2)).astype(f))
            log_var = (frac * max_log).astype(f) + (rb((f(1) - frac).astype(f)) * min_log).astype(f)
            var = np.exp(log_var).astype(f)
    else:
        pv = tb["posterior_variance"]
        if var_type == "FIXED_LARGE":
            vt = np.append(pv[1], tb["betas"][1:])
            var_t, lv_t = vt, np.log(vt)
        else:
            var_t, lv_t = pv, tb["posterior_log_variance_clipped"]
        var = np.broadcast_to(extract(var_t, t), x.shape)
        log_var = np.broadcast_to(extract(lv_t, t), x.shape)

    def proc(a):
        return np.clip(a, f(-1), f(1)) if clip_denoised else a

    if mean_type == "PREVIOUS_X":
        c1, c2 = tb["posterior_mean_coef1"], tb["posterior_mean_coef2"]
        xs = proc((extract(1.0 / c1, t) * model_output).astype(f) - (extract(c2 / c1, t) * x).astype(f))
        mean = model_output
    else:
        if mean_type == "START_X":
            xs = proc(model_output)
        elif mean_type == "EPSILON":
            xs = proc((extract(tb["sqrt_recip_alphas_cumprod"], t) * x).astype(f)
                      - (extract(tb["sqrt_recipm1_alphas_cumprod"], t) * model_output).astype(f))
        elif mean_type == "VELOCITY":   # per-sample coefficients (the reference's :392-397 only runs for N == 1)
            xs = proc((extract(tb["sqrt_alphas_cumprod"], t) * x).astype(f)
                      - (extract(tb["sqrt_one_minus_alphas_cumprod"], t) * model_output).astype(f))
        else:
            raise NotImplementedError(mean_type)
        mean = (extract(tb["posterior_mean_coef1"], t) * xs).astype(f) + (extract(tb["posterior_mean_coef2"], t) * x).astype(f)
    return dict(mean=mean, variance=var, log_variance=log_var, pred_xstart=xs)


def _nonzero_mask(t):
    return (np.asarray(t) != 0).astype(np.float32).reshape(-1, 1, 1, 1)


def p_sample(tb, mean_type, var_type, model_output, x, t, noise, clip_denoised=True, out_is_bf16=False):
    """:455-506 without cond_fn."""
    f = np.float32
    out = p_mean_variance(tb, mean_type, var_type, model_output, x, t, clip_denoised, out_is_bf16)
    lv = out["log_variance"]
    if var_type == "LEARNED" and out_is_bf16:
        e = _bf16_round(np.exp(_bf16_round((f(0.5) * lv).astype(f))).astype(f))
    else:
        e = np.exp((f(0.5) * lv).astype(f)).astype(f)
    sample = out["mean"] + ((_nonzero_mask(t) * e).astype(f) * np.asarray(noise, dtype=f)).astype(f)
    return dict(sample=sample.astype(f), pred_xstart=out["pred_xstart"])


def _eps_from_xstart(tb, x, t, xs):
    f = np.float32
    return (((extract(tb["sqrt_recip_alphas_cumprod"], t) * x).astype(f) - xs).astype(f)
            / extract(tb["sqrt_recipm1_alphas_cumprod"], t)).astype(f)


def ddim_sample(tb, mean_type, var_type, model_output, x, t, noise, eta=0.0, clip_denoised=True, out_is_bf16=False):
    """:603-651 without cond_fn."""
    f = np.float32
    x = np.asarray(x, dtype=f)
    out = p_mean_variance(tb, mean_type, var_type, model_output, x, t, clip_denoised, out_is_bf16)
    xs = out["pred_xstart"]
    eps = _eps_from_xstart(tb, x, t, xs)
    ab, abp = extract(tb["alphas_cumprod"], t), extract(tb["alphas_cumprod_prev"], t)
    one = f(1)
    sigma = ((f(eta) * np.sqrt(((one - abp).astype(f) / (one - ab).astype(f)).astype(f))).astype(f)
             * np.sqrt((one - (ab / abp).astype(f)).astype(f))).astype(f)
    mean_pred = ((xs * np.sqrt(abp)).astype(f)
                 + (np.sqrt(((one - abp).astype(f) - (sigma * sigma).astype(f)).astype(f)) * eps).astype(f)).astype(f)
    sample = mean_pred + ((_nonzero_mask(t) * sigma).astype(f) * np.asarray(noise, dtype=f)).astype(f)
    return dict(sample=sample.astype(f), pred_xstart=xs)


def ddim_reverse_sample(tb, mean_type, var_type, model_output, x, t, clip_denoised=True, out_is_bf16=False):
    """:653-689."""
    f = np.float32
    x = np.asarray(x, dtype=f)
    out = p_mean_variance(tb, mean_type, var_type, model_output, x, t, clip_denoised, out_is_bf16)
    xs = out["pred_xstart"]
    eps = _eps_from_xstart(tb, x, t, xs)
    abn = extract(tb["alphas_cumprod_next"], t)
    mean_pred = (xs * np.sqrt(abn)).astype(f) + (np.sqrt((f(1) - abn).astype(f)) * eps).astype(f)
    return dict(sample=mean_pred.astype(f), pred_xstart=xs)


def cfg_combine(cond, uncond, scale, is_bf16=False):
    """tools/sampler.py:46-48: uncond + scale * (cond - uncond), each op rounded in the tensor dtype."""
    f = np.float32
    rb = _bf16_round if is_bf16 else (lambda a: a)
    cond, uncond = np.asarray(cond, dtype=f), np.asarray(uncond, dtype=f)
    d = rb((cond - uncond).astype(f))
    m = rb((f(scale) * d).astype(f))
    return rb((uncond + m).astype(f))


# ---- respacing (tools/respace.py:9-115) -----------------------------------------------------------------------------
def space_timesteps(num_timesteps, section_counts):
    """:9-62 -> sorted list of the kept base timesteps."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            n = int(section_counts[4:])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == n:
                    return sorted(range(0, num_timesteps, i))
            raise ValueError("no integer stride gives that many steps")
        section_counts = [int(x) for x in section_counts.split(",")]
    per, extra = num_timesteps // len(section_counts), num_timesteps % len(section_counts)
    start, steps = 0, []
    for i, c in enumerate(section_counts):
        size = per + (1 if i < extra else 0)
        if size < c:
            raise ValueError("section too small")
        frac = 1 if c <= 1 else (size - 1) / (c - 1)
        cur = 0.0
        for _ in range(c):
            steps.append(start + round(cur))
            cur += frac
        start += size
    return sorted(set(steps))


def spaced_betas(betas, kept):
    """:72-84: betas of the process restricted to `kept`, so that its alphas_cumprod equals the base one there."""
    ac = np.cumprod(1.0 - np.asarray(betas, dtype=np.float64))
    last, out = 1.0, []
    for i, a in enumerate(ac):
        if i in set(kept):
            out.append(1 - a / last)
            last = a
    return np.array(out)
