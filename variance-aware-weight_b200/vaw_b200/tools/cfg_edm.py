"""EDM-style sampling (the sampler the reference's shipped recipe uses: run.sh `--solver heun`) on the library.

Mirror of /root/reference/tools/cfg_edm.py:
    Net (:14-107)              preconditioning wrapper around the (guided) denoiser: sigma table `u` of the DDPM schedule,
                               `round_sigma` (nearest table entry), c_in / c_noise / c_skip / c_out for EPSILON / START_X /
                               VELOCITY predictions
    ablation_sampler (:109-210) Euler / Heun integration of the probability-flow ODE with the vp / ve / iddpm / edm time
                               discretisations, vp / ve / linear sigma(t) schedules, vp / none scalings and stochastic churn

What runs where: the time grid and every per-step coefficient are scalars - they are evaluated on the host with the same
float64 / float32 torch operations, in the same order, as the reference evaluates them on 0-dim tensors.  Everything
that touches the [N, C, H, W] state runs in two fused kernels per denoiser evaluation (csrc/sampling_kernels.cu):
    vaw_edm_pre    x_hat = a x_cur + c noise (float64) and the denoiser input c_in * float32(x_hat / s)
    vaw_edm_post   denoised = c_skip x + c_out F(x) (float32) -> float64, d = A x - B denoised, and the Euler step /
                   Heun predictor (plus the NEXT denoiser input) / Heun corrector
instead of ~25 elementwise launches per evaluation; each float64 / float32 operation is rounded separately in the
reference's order, so for a given denoiser output the trajectory is bit-identical to the eager code.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib as L

_P, _D, _F, _I, _LL = C.c_void_p, C.c_double, C.c_float, C.c_int, C.c_longlong
L.register("vaw_edm_pre", [_P, _P, _D, _D, _D, _F, _P, _P, _LL, _P])
L.register("vaw_edm_post", [_P, _I, _LL, _P, _D, _F, _F, _I, _D, _D, _I, _D, _P, _P, _D, _D, _P, _P, _D, _F, _P, _LL, _LL,
                            _P])

PRED = {"START_X": 2, "EPSILON": 3, "VELOCITY": 4}      # ModelMeanType values
MODE_EULER, MODE_PREDICT, MODE_CORRECT = 0, 1, 2


class Net(torch.nn.Module):
    """Reference cfg_edm.py:14-107, same constructor.  `forward(x, sigma, class_labels)` is kept for API parity (it runs
    the two kernels around one denoiser call); `ablation_sampler` drives the kernels directly."""

    def __init__(self, model, img_resolution, img_channels, pred_type="EPSILON", label_dim=0, amp=False, C_1=0.001,
                 C_2=0.008, M=1000, noise_schedule="linear", lambda_max=10.0, lambda_min=-10.0):
        super().__init__()
        if pred_type not in PRED:
            raise ValueError(f"Unsupported pred_type: {pred_type}")
        self.img_resolution, self.img_channels, self.label_dim = img_resolution, img_channels, label_dim
        self.amp, self.C_1, self.C_2, self.M = amp, C_1, C_2, M
        self.model, self.noise_schedule, self.pred_type = model, noise_schedule, pred_type
        self.lambda_max, self.lambda_min = lambda_max, lambda_min
        # u_{j-1} = sqrt((u_j^2 + 1) / max(abar_{j-1} / abar_j, C_1) - 1), float32 like the reference's buffer (:42-45)
        u = torch.zeros(M + 1)
        for j in range(M, 0, -1):
            ratio = self.alpha_bar(j - 1) / self.alpha_bar(j)
            u[j - 1] = ((u[j] ** 2 + 1) / ratio.clip(min=C_1) - 1).sqrt()
        self.register_buffer("u", u)
        self.sigma_min = float(u[M - 1])
        self.sigma_max = float(u[0])
        self._u_host = u.clone()           # the table lookups are host-side scalar work

    def alpha_bar(self, j):
        j = torch.as_tensor(j)
        if self.noise_schedule == "cosine":
            return (0.5 * np.pi * j / self.M / (self.C_2 + 1)).sin() ** 2
        if self.noise_schedule == "linear":
            betas = np.linspace(0.0001, 0.02, self.M + 1, dtype=np.float64)
            return np.cumprod(1.0 - betas, axis=0)[self.M - j]
        if self.noise_schedule == "linear_logsnr":
            t = (self.M - j) / self.M
            return torch.sigmoid(self.lambda_max + t * (self.lambda_min - self.lambda_max))
        raise NotImplementedError(f"unknown path type: {self.noise_schedule}")

    def round_sigma(self, sigma, return_index=False):
        """Nearest entry of the sigma table (float32 distances, first minimum), reference :103-107."""
        sigma = torch.as_tensor(sigma)
        u = self._u_host
        index = torch.cdist(sigma.detach().cpu().to(torch.float32).reshape(1, -1, 1), u.reshape(1, -1, 1)).argmin(2)
        result = index if return_index else u[index.flatten()].to(sigma.dtype)
        return result.reshape(sigma.shape).to(sigma.device)

    # ---- scalar preconditioning coefficients (reference :51-79), float32 arithmetic on the host ----------------
    def coefficients(self, sigma):
        """sigma: python float / 0-dim tensor (float64).  -> dict of python floats and the integer c_noise."""
        s32 = torch.as_tensor(sigma, dtype=torch.float64).to(torch.float32).reshape(1)
        c_noise = self.M - 1 - self.round_sigma(s32, return_index=True).to(torch.float32)
        c_in = 1 / (s32 ** 2 + 1).sqrt()
        if self.pred_type == "EPSILON":
            c_skip, c_out = torch.ones(1), -s32
        elif self.pred_type == "START_X":
            c_skip, c_out = torch.zeros(1), torch.ones(1)
        else:
            c_skip, c_out = c_in ** 2, -s32 * c_in
        return dict(c_in=float(c_in), c_skip=float(c_skip), c_out=float(c_out), c_noise=int(c_noise.int()))

    def _denoiser(self, x_in, c_noise, class_labels, model_kwargs):
        t = torch.full((x_in.shape[0],), c_noise, dtype=torch.int32, device=x_in.device)
        t._vaw_host_value = float(c_noise)      # lets IntervalCFG test its interval without a device round trip
        raw = self.model(x_in, t, y=class_labels, **model_kwargs)
        out = raw[0] if isinstance(raw, tuple) else raw
        if out.dtype not in (torch.float32, torch.bfloat16):
            out = out.float()
        return out.contiguous()

    def forward(self, x, sigma, class_labels=None, force_fp32=False, **model_kwargs):
        """denoised = c_skip x + c_out F(c_in x, c_noise) as float32 (reference :51-79)."""
        L.require_cuda(x)
        x64 = x.to(torch.float64).contiguous()
        k = self.coefficients(sigma)
        x_in = torch.empty(x64.shape, dtype=torch.float32, device=x.device)
        L.call("vaw_edm_pre", x64.data_ptr(), None, 1.0, 0.0, 1.0, k["c_in"], None, x_in.data_ptr(), x64.numel(),
               L.stream_ptr())
        out = self._denoiser(x_in, k["c_noise"], class_labels, model_kwargs)
        den = torch.empty_like(x_in)
        chw = x64[0].numel()
        L.call("vaw_edm_post", out.data_ptr(), L.BF16 if out.dtype == torch.bfloat16 else L.F32, out[0].numel(),
               x64.data_ptr(), 1.0, k["c_skip"], k["c_out"], PRED[self.pred_type], 0.0, 0.0, -1, 0.0, None, None, 0.0, 0.0,
               None, None, 1.0, 0.0, den.data_ptr(), x64.shape[0], chw, L.stream_ptr())
        return den


def _schedule(net, num_steps, sigma_min, sigma_max, rho, discretization, schedule, scaling, epsilon_s, C_1, C_2, M):
    """Time grid and the sigma(t) / s(t) function families of the reference (:118-185), float64 0-dim tensors."""
    vp_sigma = lambda beta_d, beta_min: lambda t: (np.e ** (0.5 * beta_d * (t ** 2) + beta_min * t) - 1) ** 0.5
    vp_sigma_inv = lambda beta_d, beta_min: lambda sigma: ((beta_min ** 2 + 2 * beta_d * (sigma ** 2 + 1).log()).sqrt() - beta_min) / beta_d
    if sigma_min is None:
        sigma_min = {"vp": vp_sigma(19.9, 0.1)(epsilon_s), "ve": 0.02, "iddpm": 0.002, "edm": 0.002}[discretization]
    if sigma_max is None:
        sigma_max = {"vp": vp_sigma(19.9, 0.1)(1), "ve": 100, "iddpm": 81, "edm": 80}[discretization]
    sigma_min = max(sigma_min, net.sigma_min)
    sigma_max = min(sigma_max, net.sigma_max)
    vp_beta_d = 2 * (np.log(sigma_min ** 2 + 1) / epsilon_s - np.log(sigma_max ** 2 + 1)) / (epsilon_s - 1)
    vp_beta_min = np.log(sigma_max ** 2 + 1) - 0.5 * vp_beta_d
    idx = torch.arange(num_steps, dtype=torch.float64)
    if discretization == "vp":
        sigma_steps = vp_sigma(vp_beta_d, vp_beta_min)(1 + idx / (num_steps - 1) * (epsilon_s - 1))
    elif discretization == "ve":
        sigma_steps = ((sigma_max ** 2) * ((sigma_min ** 2 / sigma_max ** 2) ** (idx / (num_steps - 1)))).sqrt()
    elif discretization == "iddpm":
        u = torch.zeros(M + 1, dtype=torch.float64)
        abar = lambda j: (0.5 * np.pi * j / M / (C_2 + 1)).sin() ** 2
        for j in torch.arange(M, 0, -1):
            u[j - 1] = ((u[j] ** 2 + 1) / (abar(j - 1) / abar(j)).clip(min=C_1) - 1).sqrt()
        kept = u[torch.logical_and(u >= sigma_min, u <= sigma_max)]
        sigma_steps = kept[((len(kept) - 1) / (num_steps - 1) * idx).round().to(torch.int64)]
    else:
        sigma_steps = (sigma_max ** (1 / rho) + idx / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    if schedule == "vp":
        sigma = vp_sigma(vp_beta_d, vp_beta_min)
        sigma_deriv = lambda t: 0.5 * (vp_beta_min + vp_beta_d * t) * (sigma(t) + 1 / sigma(t))
        sigma_inv = vp_sigma_inv(vp_beta_d, vp_beta_min)
    elif schedule == "ve":
        sigma, sigma_deriv, sigma_inv = (lambda t: t.sqrt()), (lambda t: 0.5 / t.sqrt()), (lambda s_: s_ ** 2)
    else:
        sigma, sigma_deriv, sigma_inv = (lambda t: t), (lambda t: 1), (lambda s_: s_)
    if scaling == "vp":
        s = lambda t: 1 / (1 + sigma(t) ** 2).sqrt()
        s_deriv = lambda t: -sigma(t) * sigma_deriv(t) * (s(t) ** 3)
    else:
        s, s_deriv = (lambda t: 1), (lambda t: 0)
    t_steps = sigma_inv(net.round_sigma(sigma_steps))
    t_steps = torch.cat([t_steps, torch.zeros_like(t_steps[:1])])
    return t_steps, sigma, sigma_deriv, sigma_inv, s, s_deriv


@torch.no_grad()
def ablation_sampler(net, latents, class_labels=None, randn_like=torch.randn_like, num_steps=18, sigma_min=None,
                     sigma_max=None, rho=7, solver="heun", discretization="edm", schedule="linear", scaling="none",
                     epsilon_s=1e-3, C_1=0.001, C_2=0.008, M=1000, alpha=1, S_churn=0, S_min=0, S_max=float("inf"),
                     S_noise=1, **model_kwargs):
    """Reference cfg_edm.py:109-210, same signature and result (float64 [N, C, H, W])."""
    assert solver in ["euler", "heun"]
    assert discretization in ["vp", "ve", "iddpm", "edm"]
    assert schedule in ["vp", "ve", "linear"]
    assert scaling in ["vp", "none"]
    L.require_cuda(latents)
    dev = latents.device
    t_steps, sigma, sigma_deriv, sigma_inv, s, s_deriv = _schedule(
        net, num_steps, sigma_min, sigma_max, rho, discretization, schedule, scaling, epsilon_s, C_1, C_2, M)
    f = float
    pred = PRED[net.pred_type]
    N, chw, n = latents.shape[0], latents[0].numel(), latents.numel()
    t_next = t_steps[0]
    x_next = latents.to(torch.float64) * f(sigma(t_next) * s(t_next))
    x_hat, x_prime, d_cur = (torch.empty_like(x_next) for _ in range(3))
    x_in = torch.empty(x_next.shape, dtype=torch.float32, device=dev)

    def post(out, x_src, t_src, k, mode, h_coef, d_prev=None, c1=0.0, c2=0.0, x_out=None, d_out=None, nxt=None):
        A = f(sigma_deriv(t_src) / sigma(t_src) + s_deriv(t_src) / s(t_src))
        Bc = f(sigma_deriv(t_src) * s(t_src) / sigma(t_src))
        L.call("vaw_edm_post", out.data_ptr(), L.BF16 if out.dtype == torch.bfloat16 else L.F32, out[0].numel(),
               x_src.data_ptr(), f(s(t_src)), k["c_skip"], k["c_out"], pred, A, Bc, mode, f(h_coef), x_hat.data_ptr(),
               L.ptr(d_prev), c1, c2, L.ptr(x_out), L.ptr(d_out), nxt[0] if nxt else 1.0, nxt[1] if nxt else 0.0,
               x_in.data_ptr() if nxt else None, N, chw, L.stream_ptr())

    for i, (t_cur, t_next) in enumerate(zip(t_steps[:-1], t_steps[1:])):
        x_cur = x_next
        gamma = min(S_churn / num_steps, np.sqrt(2) - 1) if S_min <= sigma(t_cur) <= S_max else 0
        t_hat = sigma_inv(net.round_sigma(sigma(t_cur) + gamma * sigma(t_cur)))
        a = s(t_hat) / s(t_cur)
        c = (sigma(t_hat) ** 2 - sigma(t_cur) ** 2).clip(min=0).sqrt() * s(t_hat) * S_noise
        noise = randn_like(x_cur)          # drawn every step, like the reference (keeps the generator stream aligned)
        k_hat = net.coefficients(sigma(t_hat))
        L.call("vaw_edm_pre", x_cur.data_ptr(), noise.to(torch.float64).contiguous().data_ptr(), f(a), f(c),
               f(s(t_hat)), k_hat["c_in"], x_hat.data_ptr(), x_in.data_ptr(), n, L.stream_ptr())
        h = t_next - t_hat
        out = net._denoiser(x_in, k_hat["c_noise"], class_labels, model_kwargs)
        x_next = torch.empty_like(x_hat)
        if solver == "euler" or i == num_steps - 1:
            post(out, x_hat, t_hat, k_hat, MODE_EULER, h, x_out=x_next)
            continue
        t_prime = t_hat + alpha * h
        k_prime = net.coefficients(sigma(t_prime))
        # predictor: d_cur, x_prime = x_hat + (alpha h) d_cur and the next denoiser input c_in' * float32(x_prime / s')
        post(out, x_hat, t_hat, k_hat, MODE_PREDICT, alpha * h, x_out=x_prime, d_out=d_cur,
             nxt=(f(s(t_prime)), k_prime["c_in"]))
        out = net._denoiser(x_in, k_prime["c_noise"], class_labels, model_kwargs)
        # corrector: x_next = x_hat + h ((1 - 1/(2 alpha)) d_cur + 1/(2 alpha) d_prime)
        post(out, x_prime, t_prime, k_prime, MODE_CORRECT, h, d_prev=d_cur, c1=1 - 1 / (2 * alpha), c2=1 / (2 * alpha),
             x_out=x_next)
    return x_next
