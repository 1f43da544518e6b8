"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's flow-matching objective and
its SDE sampler.

Follows /root/reference/tools/gaussian_diffusion.py, class FlowMatching:
    interpolant                       :1182-1203   (alpha_t, sigma_t, d_alpha_t, d_sigma_t) for linear / cosine / linear_logsnr
    convert_model_output_to_vector    :1206-1228
    convert_model_output_to_score     :1230-1257
    q_sample                          :1273-1277   x_t = fl(fl(a x0) + fl(s eps))
    compute_target                    :1280-1294   START_X / EPSILON / VELOCITY / VECTOR / SCORE
    training_losses                   :1297-1340   mse = w * mean_flat((target - out)^2), w from compute_mse_loss_weight
    compute_diffusion, sde_sample     :1366-1408   Euler-Maruyama / stochastic Heun with a final noise-free Euler step

numpy float32, one rounding per reference operation (torch's eager fp32 elementwise semantics: a Python scalar operand
is first rounded to float32).  Pinned against the executed reference by tests/golden/make_golden.py ->
tests/golden/flow_golden.npz (tests/test_oracle_flow.py).  cos / sin / sigmoid go through the host libm, which differs
from torch's vectorised CPU kernels and from the device's in the last bit: comparisons through them carry a 2-ulp
tolerance, everything else is exact.  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

from .diffusion import loss_weight

f32 = np.float32


def _sigmoid(x):
    return (f32(1.0) / (f32(1.0) + np.exp(-x).astype(f32))).astype(f32)


def interpolant(path_type, t):
    """-> (alpha, sigma, d_alpha, d_sigma), float32 arrays shaped like t (:1182-1203)."""
    t = np.asarray(t, dtype=f32)
    if path_type == "linear":
        return (f32(1.0) - t).astype(f32), t.copy(), np.full_like(t, -1.0), np.full_like(t, 1.0)
    if path_type == "cosine":
        arg = ((t * f32(np.pi)).astype(f32) / f32(2.0)).astype(f32)
        a, s = np.cos(arg).astype(f32), np.sin(arg).astype(f32)
        half_pi = f32(np.pi / 2)
        return a, s, (f32(-np.pi / 2) * s).astype(f32), (half_pi * a).astype(f32)
    if path_type == "linear_logsnr":
        lam = (f32(10.0) + (t * f32(-20.0)).astype(f32)).astype(f32)
        a, s = _sigmoid((f32(0.5) * lam).astype(f32)), _sigmoid((f32(-0.5) * lam).astype(f32))
        da = ((f32(-10.0) * a).astype(f32) * s).astype(f32)
        return a, s, da, (-da).astype(f32)
    raise NotImplementedError(path_type)


def _e(v, x):
    return np.asarray(v, dtype=f32).reshape(-1, *([1] * (np.ndim(x) - 1)))


def q_sample(path_type, x0, eps, t):
    a, s, _, _ = interpolant(path_type, t)
    x0, eps = np.asarray(x0, dtype=f32), np.asarray(eps, dtype=f32)
    return ((_e(a, x0) * x0).astype(f32) + (_e(s, x0) * eps).astype(f32)).astype(f32)


def target(path_type, mean_type, x0, eps, t):
    a, s, da, ds = interpolant(path_type, t)
    x0, eps = np.asarray(x0, dtype=f32), np.asarray(eps, dtype=f32)
    if mean_type == "START_X":
        return x0
    if mean_type == "EPSILON":
        return eps
    if mean_type == "VELOCITY":
        return ((_e(a, x0) * eps).astype(f32) - (_e(s, x0) * x0).astype(f32)).astype(f32)
    if mean_type == "VECTOR":
        return ((_e(da, x0) * x0).astype(f32) + (_e(ds, x0) * eps).astype(f32)).astype(f32)
    if mean_type == "SCORE":
        return ((-eps).astype(f32) / _e(s, x0)).astype(f32)
    raise NotImplementedError(mean_type)


def mse_terms(path_type, mean_type, weight_type, x0, eps, t, model_output):
    """(mse [N], d mse_n / d model_output) accumulated in float64 (the reference value up to fp32 summation order)."""
    a, s, _, _ = interpolant(path_type, t)
    w = loss_weight(mean_type, weight_type, a, s).astype(np.float64)
    d = target(path_type, mean_type, x0, eps, t).astype(np.float64) - np.asarray(model_output, dtype=np.float64)
    chw = d[0].size
    mse = w * (d.reshape(d.shape[0], -1) ** 2).mean(axis=1)
    grad = (-2.0 / chw) * w.reshape(-1, *([1] * (d.ndim - 1))) * d
    return mse, grad


def to_vector(path_type, mean_type, model_output, x_t, t):
    """convert_model_output_to_vector (:1206-1228); t per sample."""
    a, s, da, ds = (_e(v, x_t) for v in interpolant(path_type, t))
    mo, x_t = np.asarray(model_output, dtype=f32), np.asarray(x_t, dtype=f32)
    if mean_type == "VECTOR":
        return mo
    if mean_type == "START_X":
        xs = mo
        noise = ((x_t - (a * xs).astype(f32)).astype(f32) / s).astype(f32)
    elif mean_type == "EPSILON":
        noise = mo
        xs = ((x_t - (s * noise).astype(f32)).astype(f32) / a).astype(f32)
    elif mean_type == "VELOCITY":
        den = ((a * a).astype(f32) + (s * s).astype(f32)).astype(f32)
        xs = (((a * x_t).astype(f32) - (s * mo).astype(f32)).astype(f32) / den).astype(f32)
        noise = (((s * x_t).astype(f32) + (a * mo).astype(f32)).astype(f32) / den).astype(f32)
    else:
        raise NotImplementedError(mean_type)
    return ((da * xs).astype(f32) + (ds * noise).astype(f32)).astype(f32)


def to_score(path_type, mean_type, model_output, x_t, t):
    """convert_model_output_to_score (:1230-1257)."""
    a, s, da, ds = (_e(v, x_t) for v in interpolant(path_type, t))
    mo, x_t = np.asarray(model_output, dtype=f32), np.asarray(x_t, dtype=f32)
    if mean_type == "SCORE":
        return mo
    if mean_type == "START_X":
        return ((-(x_t - (a * mo).astype(f32)).astype(f32)).astype(f32) / (s * s).astype(f32)).astype(f32)
    if mean_type == "EPSILON":
        noise = mo
    elif mean_type == "VELOCITY":
        den = ((a * a).astype(f32) + (s * s).astype(f32)).astype(f32)
        noise = (((s * x_t).astype(f32) + (a * mo).astype(f32)).astype(f32) / den).astype(f32)
    elif mean_type == "VECTOR":
        den = ((s * da).astype(f32) - (a * ds).astype(f32)).astype(f32)
        noise = (((da * x_t).astype(f32) - (a * mo).astype(f32)).astype(f32) / den).astype(f32)
    else:
        raise NotImplementedError(mean_type)
    return ((-noise).astype(f32) / s).astype(f32)


def sde_times(num_steps):
    """float64 grid of sde_sample (:1377-1378): linspace(1, 0.04, n) then 0."""
    return np.append(np.linspace(1.0, 0.04, num_steps, dtype=np.float64), 0.0)


def diffusion_coef(path_type, t):
    _, s, _, ds = interpolant(path_type, t)
    return ((f32(2.0) * s).astype(f32) * ds).astype(f32)


def drift(path_type, mean_type, model_output, x, t):
    """compute_drift (:1371-1375): vector - 0.5 * diffusion * score."""
    diff = _e(diffusion_coef(path_type, t), x)
    v = to_vector(path_type, mean_type, model_output, x, t)
    sc = to_score(path_type, mean_type, model_output, x, t)
    return (v - ((f32(0.5) * diff).astype(f32) * sc).astype(f32)).astype(f32)


def sde_sample(path_type, mean_type, model_fn, start, noises, num_steps, solver):
    """sde_sample (:1370-1408).  model_fn(x, t[N]) -> prediction; noises: iterable of randn_like draws."""
    ts = sde_times(num_steps)
    x = np.asarray(start, dtype=f32)
    N = x.shape[0]
    it = iter(noises)
    for cur, nxt in zip(ts[:-2], ts[1:-1]):
        step = f32(nxt - cur)                      # 0-dim float64 tensor operand -> rounded to fp32 by the fp32 op
        sq = f32(np.sqrt(np.abs(nxt - cur)))
        tc = np.full(N, cur, dtype=np.float64).astype(f32)
        diff = _e(diffusion_coef(path_type, tc), x)
        d_cur = drift(path_type, mean_type, model_fn(x, tc), x, tc)
        noise_term = ((np.sqrt(diff).astype(f32) * np.asarray(next(it), dtype=f32)).astype(f32) * sq).astype(f32)
        pred = (((x + (d_cur * step).astype(f32)).astype(f32)) + noise_term).astype(f32)
        if solver == "euler":
            x = pred
        elif solver == "heun":
            tn = np.full(N, nxt, dtype=np.float64).astype(f32)
            d_next = drift(path_type, mean_type, model_fn(pred, tn), pred, tn)
            avg = (f32(0.5) * (d_cur + d_next).astype(f32)).astype(f32)
            x = (((x + (avg * step).astype(f32)).astype(f32)) + noise_term).astype(f32)
        else:
            raise ValueError(f"Unknown solver: {solver}")
    cur, nxt = ts[-2], ts[-1]
    tc = np.full(N, cur, dtype=np.float64).astype(f32)
    d_cur = drift(path_type, mean_type, model_fn(x, tc), x, tc)
    return (x + (d_cur * f32(nxt - cur)).astype(f32)).astype(f32)
