"""Diffusion training objective behind the reference's API, executed by the sm_100a kernels.

Host-side mirror of /root/reference/tools/gaussian_diffusion.py for the TRAINING path only:
    ModelMeanType / ModelVarType / LossType            (:21-56)
    get_named_beta_schedule / betas_for_alpha_bar      (:59-123)
    GaussianDiffusion(args=, betas=, ...)              (:126-205)  q_sample (:234) sample_t (:810) compute_target (:818)
                                                                   training_losses (:834-930) _scale_timesteps (:417)
    FlowMatching(args=, model_mean_type=)              (:1151-1340)
    compute_mse_loss_weight (:1092-1148), compute_align_loss (:1007-1046), _extract_into_tensor (:1059-1072)
    create_gaussian_diffusion(**kw)                    alias asked for by the north-star text

training_losses launches K1 (fused q_sample + target), the denoiser, and K2 (fused weighted-MSE forward+backward)
through the C ABI (include/vaw_b200.h); with a learned variance or a KL loss type the variational-bound term
(:775-808, :862-906) is one more fused pass (vaw_vb_terms).
"""
from __future__ import annotations

import enum
import math
from types import SimpleNamespace

import numpy as np
import torch as th

from .. import _lib as L


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()
    VELOCITY = enum.auto()
    VECTOR = enum.auto()
    SCORE = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()

    def is_vb(self):
        return self in (LossType.KL, LossType.RESCALED_KL)


# ---------------------------------------------------------------------------------------------------------
# schedules (float64 on the host, evaluated with the same scalar libm calls as the reference)
# ---------------------------------------------------------------------------------------------------------
def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    T = num_diffusion_timesteps
    out = np.empty(T, dtype=np.float64)
    for i in range(T):
        lo, hi = i / T, (i + 1) / T
        out[i] = min(1 - alpha_bar(hi) / alpha_bar(lo), max_beta)
    return out


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps, lambda_max=10.0, lambda_min=-10.0):
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(num_diffusion_timesteps,
                                   lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    if schedule_name == "linear_logsnr":
        def alpha_bar(t):
            logsnr = lambda_max + t * (lambda_min - lambda_max)
            return 1.0 / (1.0 + math.exp(-logsnr))
        return betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


# ---------------------------------------------------------------------------------------------------------
# module-level helpers with the reference's signatures
# ---------------------------------------------------------------------------------------------------------
def _extract_into_tensor(arr, timesteps, broadcast_shape):
    """float64 table -> float32 values gathered at `timesteps`, expanded to `broadcast_shape` (reference :1059)."""
    res = th.from_numpy(np.asarray(arr)).to(device=timesteps.device)[timesteps].float()
    while res.dim() < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape)


def _parse_weight_type(name):
    """weight_type string -> (kind code, k)."""
    simple = {"constant": L.W_CONSTANT, "lambda": L.W_LAMBDA, "debias": L.W_DEBIAS, "min_debias": L.W_MIN_DEBIAS,
              "max_debias": L.W_MAX_DEBIAS, "p2": L.W_P2, "trunc_snr": L.W_TRUNC_SNR, "snr": L.W_SNR,
              "inv_snr": L.W_INV_SNR}
    if name in simple:
        return simple[name], 0.0
    if name.startswith("min_snr_"):
        return L.W_MIN_SNR, float(name.split("min_snr_")[-1])
    if name.startswith("max_snr_"):
        return L.W_MAX_SNR, float(name.split("max_snr_")[-1])
    raise ValueError(f"Invalid mse_loss_weight_type: {name}")


def compute_mse_loss_weight(model_mean_type, mse_loss_weight_type, t, alpha, sigma, p2_k=1.0, p2_gamma=1.0):
    """Per-sample loss weight from fp32 alpha/sigma tensors, reference semantics (:1092-1148, table in SURVEY §A.1).
    Elementwise on [N] values; the hot path uses the per-timestep LUT (vaw_loss_weight_lut) instead."""
    if mse_loss_weight_type == "constant":
        return th.ones_like(t)
    snr = (alpha / sigma) ** 2
    kind, k = _parse_weight_type(mse_loss_weight_type)
    name = model_mean_type.name
    w = None
    if name == "EPSILON":
        if kind == L.W_MIN_SNR:
            w = th.clamp(snr, max=k) / snr
        elif kind == L.W_MAX_SNR:
            w = th.clamp(snr, min=k) / snr
        elif kind == L.W_LAMBDA:
            w = sigma.clone()
        elif kind == L.W_DEBIAS:
            w = sigma / alpha
        elif kind == L.W_P2:
            w = 1 / (p2_k + snr) ** p2_gamma
        elif kind == L.W_MIN_DEBIAS:
            w = th.clamp(sigma / alpha, max=1.0)
        elif kind == L.W_MAX_DEBIAS:
            w = th.clamp(sigma / alpha, min=1.0)
    elif name == "START_X":
        if kind == L.W_TRUNC_SNR:
            w = th.clamp(snr, min=1.0)
        elif kind == L.W_SNR:
            w = snr.clone()
        elif kind == L.W_INV_SNR:
            w = 1.0 / snr
        elif kind == L.W_MIN_SNR:
            w = th.clamp(snr, max=k)
        elif kind == L.W_MAX_SNR:
            w = th.clamp(snr, min=k)
        elif kind == L.W_LAMBDA:
            w = alpha.clone()
    elif name == "VECTOR":
        if kind == L.W_LAMBDA:
            w = th.ones_like(t)
    elif name == "VELOCITY":
        if kind == L.W_MIN_SNR:
            w = th.clamp(snr, max=k) / (snr + 1)
        elif kind == L.W_LAMBDA:
            w = alpha * sigma
    if w is None:
        raise ValueError(f"Invalid mse_loss_weight_type: {mse_loss_weight_type}")
    w[snr == 0] = 1.0
    return w


class _AlignMSE(th.autograd.Function):
    """REPA alignment loss, type 'mse' (reference :1011-1013): one fused pass gives the scalar and d zs."""

    @staticmethod
    def forward(ctx, zs, feat):
        L.require_cuda(zs, feat)
        if zs.shape != feat.shape:     # F.mse_loss would broadcast-or-raise; a flat kernel would read out of bounds
            raise ValueError(f"align loss: projector output {tuple(zs.shape)} and teacher features {tuple(feat.shape)} "
                             "must have the same shape")
        zs_c, feat_c = zs.contiguous(), feat.detach().contiguous()
        dmap = {th.float32: L.F32, th.bfloat16: L.BF16}
        if zs_c.dtype not in dmap:
            zs_c = zs_c.float()
        if feat_c.dtype not in dmap:
            feat_c = feat_c.float()
        dz = th.empty_like(zs_c)
        part = th.empty(1024, dtype=th.float32, device=zs.device)
        loss = th.empty((), dtype=th.float32, device=zs.device)
        L.call("vaw_align_mse", zs_c.data_ptr(), dmap[zs_c.dtype], feat_c.data_ptr(), dmap[feat_c.dtype],
               dz.data_ptr(), 1.0, zs_c.numel(), part.data_ptr(), loss.data_ptr(), L.stream_ptr())
        ctx.save_for_backward(dz)
        ctx.in_dtype = zs.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        out = th.empty_like(dz)
        s = g.reshape(1).float().contiguous()
        L.call("vaw_scale_rows", dz.data_ptr(), s.data_ptr(), out.data_ptr(),
               L.F32 if dz.dtype == th.float32 else L.BF16, 1, dz.numel(), L.stream_ptr())
        return out.to(ctx.in_dtype), None


L.register("vaw_align_mse", [L.C.c_void_p, L.C.c_int, L.C.c_void_p, L.C.c_int, L.C.c_void_p, L.C.c_float,
                             L.C.c_longlong, L.C.c_void_p, L.C.c_void_p, L.C.c_void_p])
L.register("vaw_align_mse_bwd", [L.C.c_void_p, L.C.c_int, L.C.c_void_p, L.C.c_int, L.C.c_void_p, L.C.c_void_p,
                                 L.C.c_longlong, L.C.c_void_p])
L.register("vaw_align_rowwise", [L.C.c_void_p, L.C.c_int, L.C.c_void_p, L.C.c_int, L.C.c_int, L.C.c_void_p, L.C.c_float,
                                 L.C.c_longlong, L.C.c_int, L.C.c_void_p, L.C.c_void_p, L.C.c_void_p])


class _AlignMSEFused(th.autograd.Function):
    """'mse' alignment loss whose VALUE was already accumulated in the epilogue of the last projector GEMM
    (vaw_dit_forward_align / VAW_EPI_ALIGN_MSE): this node only carries it into the graph; the backward is one pass
    dzs = g * 2 (zs - feat) / n with the upstream gradient read from the device."""

    @staticmethod
    def forward(ctx, zs, feat, loss):
        ctx.save_for_backward(zs, feat)
        return loss.clone()

    @staticmethod
    def backward(ctx, g):
        zs, feat = ctx.saved_tensors
        dz = th.empty_like(zs)
        gd_ = g.reshape(1).float().contiguous()
        L.call("vaw_align_mse_bwd", zs.data_ptr(), L.BF16, feat.data_ptr(), L.BF16, gd_.data_ptr(), dz.data_ptr(),
               zs.numel(), L.stream_ptr())
        return dz, None, None


class _AlignRowwise(th.autograd.Function):
    """'cosine' (reference :1008-1009) and 'mse_l2' (:1014-1017): one warp per token row, loss and d zs in one pass."""

    @staticmethod
    def forward(ctx, zs, feat, kind):
        L.require_cuda(zs, feat)
        if zs.shape != feat.shape:
            raise ValueError(f"align loss: projector output {tuple(zs.shape)} and teacher features {tuple(feat.shape)} "
                             "must have the same shape")
        dmap = {th.float32: L.F32, th.bfloat16: L.BF16}
        zs_c, feat_c = zs.contiguous(), feat.detach().contiguous()
        if zs_c.dtype not in dmap:
            zs_c = zs_c.float()
        if feat_c.dtype not in dmap:
            feat_c = feat_c.float()
        D = zs_c.shape[-1]
        rows = zs_c.numel() // D
        dz = th.empty_like(zs_c)
        part = th.empty(rows, dtype=th.float32, device=zs.device)
        loss = th.empty((), dtype=th.float32, device=zs.device)
        L.call("vaw_align_rowwise", zs_c.data_ptr(), dmap[zs_c.dtype], feat_c.data_ptr(), dmap[feat_c.dtype], kind,
               dz.data_ptr(), 1.0, rows, D, part.data_ptr(), loss.data_ptr(), L.stream_ptr())
        ctx.save_for_backward(dz)
        ctx.in_dtype = zs.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        out = th.empty_like(dz)
        s = g.reshape(1).float().contiguous()
        L.call("vaw_scale_rows", dz.data_ptr(), s.data_ptr(), out.data_ptr(),
               L.F32 if dz.dtype == th.float32 else L.BF16, 1, dz.numel(), L.stream_ptr())
        return out.to(ctx.in_dtype), None, None


def compute_align_loss(target, output, type, temperature=0.1):
    """Projection-alignment loss (reference :1007-1046).  'mse' (the default and the benchmarked type), 'cosine' and
    'mse_l2' run library kernels (value + gradient in one pass); 'nt_xent' needs the [N*T, N*T] similarity matrix
    (16384^2 logits at the benchmark batch - SURVEY a9 marks it impractical) and stays a plain tensor expression."""
    import torch.nn.functional as F
    if type == "mse":
        fused = getattr(output, "_vaw_align", None)
        if fused is not None and fused[1] is target:
            return _AlignMSEFused.apply(output, fused[2], fused[0])
        return _AlignMSE.apply(output, target)
    if type == "cosine":
        return _AlignRowwise.apply(output, target, 0)
    if type == "mse_l2":
        return _AlignRowwise.apply(output, target, 1)
    if type == "nt_xent":
        assert temperature > 0, "temperature must be > 0"
        n, tt, d = target.shape
        tg = F.normalize(target.reshape(n * tt, d).float(), dim=1)
        ou = F.normalize(output.reshape(n * tt, d).float(), dim=1)
        logits = ou @ tg.T / temperature
        labels = th.arange(n * tt, device=logits.device)
        return 0.5 * (F.cross_entropy(logits, labels) + F.cross_entropy(logits.T, labels))
    raise ValueError(f"Unknown align loss type: {type}.")


# ---------------------------------------------------------------------------------------------------------
# K2 as an autograd node
# ---------------------------------------------------------------------------------------------------------
class _WeightedMSE(th.autograd.Function):
    """terms['mse'] = w_t * mean((target - out)^2) with the gradient w.r.t. `out` produced in the same pass."""

    @staticmethod
    def forward(ctx, out, x0, noise, t, coef, mean_code):
        # coef = (tab_alpha, tab_sigma, tab_c0, tab_c1, w_tab); tables are [T] when t is given, per-sample [N] otherwise
        # `out` may carry the variance channels behind the mean channels ([N, 2C, H, W], reference :889-891): the mean
        # half is read in place through the row stride and the gradient of the variance half is zero
        L.require_cuda(out, x0, noise)
        N = out.shape[0]
        chw = x0[0].numel() if N else 1
        o = out.contiguous()
        if o.dtype not in (th.float32, th.bfloat16):
            o = o.float()
        stride = o[0].numel() if N else chw
        code = L.F32 if o.dtype == th.float32 else L.BF16
        mse = th.empty(N, dtype=th.float32, device=out.device)
        need_grad = out.requires_grad
        g = None
        if need_grad:
            g = th.empty_like(o) if stride == chw else th.zeros_like(o)
        ta, ts, c0, c1, wt = coef
        L.call("vaw_wmse_fwd_bwd_strided", o.data_ptr(), code, stride, x0.data_ptr(), noise.data_ptr(), L.ptr(t),
               ta.data_ptr(), ts.data_ptr(), L.ptr(c0), L.ptr(c1), L.ptr(wt), mse.data_ptr(), None, L.ptr(g), stride,
               None, 1.0, mean_code, N, chw, L.stream_ptr())
        ctx.g, ctx.in_dtype, ctx.code = g, out.dtype, code
        return mse

    @staticmethod
    def backward(ctx, gm):
        g = ctx.g
        if g is None:
            return None, None, None, None, None, None
        out = th.empty_like(g)
        s = gm.float().contiguous()
        L.call("vaw_scale_rows", g.data_ptr(), s.data_ptr(), out.data_ptr(), ctx.code, g.shape[0], g[0].numel(),
               L.stream_ptr())
        return out.to(ctx.in_dtype), None, None, None, None, None


L.register("vaw_wmse_fwd_bwd_strided", [L.C.c_void_p, L.C.c_int, L.C.c_longlong] + [L.C.c_void_p] * 11 +
           [L.C.c_longlong, L.C.c_void_p, L.C.c_float, L.C.c_int, L.C.c_longlong, L.C.c_longlong, L.C.c_void_p])
L.register("vaw_flow_sde_step", [L.C.c_void_p, L.C.c_int, L.C.c_void_p, L.C.c_void_p, L.C.c_int, L.C.c_int] +
           [L.C.c_void_p] * 3 + [L.C.c_float] * 3 + [L.C.c_void_p, L.C.c_void_p, L.C.c_longlong, L.C.c_void_p])
L.register("vaw_vb_terms", [L.C.c_void_p, L.C.c_int, L.C.c_longlong] + [L.C.c_void_p] * 4 + [L.C.c_int] +
           [L.C.c_void_p] * 2 + [L.C.c_longlong, L.C.c_void_p, L.C.c_float, L.C.c_int, L.C.c_int, L.C.c_int, L.C.c_float,
                                 L.C.c_longlong, L.C.c_longlong, L.C.c_void_p])


_U64 = L.C.c_ulonglong
L.register("vaw_philox_offset_increment", [L.C.c_longlong, L.C.c_void_p])
L.register("vaw_qsample_philox", [L.C.c_void_p, L.C.c_void_p, L.C.c_float, _U64, _U64, _U64] + [L.C.c_void_p] * 9 +
           [L.C.c_int, L.C.c_longlong, L.C.c_longlong, L.C.c_void_p])
L.register("vaw_randint_philox", [_U64, _U64, L.C.c_longlong, L.C.c_longlong, L.C.c_void_p, L.C.c_longlong, L.C.c_void_p])
L.register("vaw_wmse_fwd_bwd_philox", [L.C.c_void_p, L.C.c_int, L.C.c_longlong, L.C.c_void_p, _U64, _U64] +
           [L.C.c_void_p] * 8 + [L.C.c_longlong, L.C.c_float, L.C.c_int, L.C.c_longlong, L.C.c_longlong, L.C.c_void_p])


class DeferredLatent:
    """What `vaw_b200.tools.trainer.sample_from_latent(latent, scale, defer=True)` returns: the 8-channel (mean | std)
    latent and its scale, NOT yet sampled.  `training_losses` accepts it as `x_start` and draws eps1 inside K1 (same
    generator stream as the reference's `mean + std * randn_like(mean)`, trainer.py:21-25), so neither eps1 nor - for the
    eps objective - x_start itself is ever written to HBM.  `.tensor()` materialises it the ordinary way."""

    def __init__(self, latent, latent_scale=1.0):
        if latent.dim() < 2 or latent.shape[1] % 2:
            raise ValueError("latent must be [N, 2C, ...] (mean | std along the channels)")
        self.latent, self.latent_scale = latent, float(latent_scale)
        self.shape = th.Size((latent.shape[0], latent.shape[1] // 2, *latent.shape[2:]))
        self.device, self.is_cuda, self.dtype = latent.device, latent.is_cuda, th.float32

    def tensor(self):
        from .trainer import sample_from_latent
        return sample_from_latent(self.latent, self.latent_scale)


class _PhiloxDraws:
    """Bookkeeping of the device generator for draws made inside our kernels: hands out (seed, offset) pairs and advances
    torch's generator by exactly what ATen's own kernels would have consumed (vaw_philox_offset_increment)."""

    def __init__(self, device):
        idx = device.index if device.index is not None else th.cuda.current_device()
        self.gen = th.cuda.default_generators[idx]
        self.seed = int(self.gen.initial_seed())

    def take(self, numel):
        off = int(self.gen.get_offset())
        inc = _U64()
        L.call("vaw_philox_offset_increment", int(numel), L.C.byref(inc))
        self.gen.set_offset(off + int(inc.value))
        return off


class _WeightedMSEPhilox(th.autograd.Function):
    """K2 for a step whose noise was drawn inside K1: the eps each element needs is re-drawn from (seed, offset)."""

    @staticmethod
    def forward(ctx, out, x0, t, coef, mean_code, seed, offset, chw):
        L.require_cuda(out)
        N = out.shape[0]
        o = out.contiguous()
        if o.dtype not in (th.float32, th.bfloat16):
            o = o.float()
        stride = o[0].numel() if N else chw
        code = L.F32 if o.dtype == th.float32 else L.BF16
        mse = th.empty(N, dtype=th.float32, device=out.device)
        g = None
        if out.requires_grad:
            g = th.empty_like(o) if stride == chw else th.zeros_like(o)
        ta, ts, c0, c1, wt = coef
        L.call("vaw_wmse_fwd_bwd_philox", o.data_ptr(), code, stride, L.ptr(x0), seed, offset, L.ptr(t), ta.data_ptr(),
               ts.data_ptr(), L.ptr(c0), L.ptr(c1), L.ptr(wt), mse.data_ptr(), L.ptr(g), stride, 1.0, mean_code, N, chw,
               L.stream_ptr())
        ctx.g, ctx.in_dtype, ctx.code = g, out.dtype, code
        return mse

    @staticmethod
    def backward(ctx, gm):
        g = ctx.g
        if g is None:
            return (None,) * 8
        res = th.empty_like(g)
        s = gm.float().contiguous()
        L.call("vaw_scale_rows", g.data_ptr(), s.data_ptr(), res.data_ptr(), ctx.code, g.shape[0], g[0].numel(),
               L.stream_ptr())
        return (res.to(ctx.in_dtype),) + (None,) * 7


class _VbTerm(th.autograd.Function):
    """terms['vb'] = _vb_terms_bpd(...)["output"] (reference :775-808) with d vb_n / d out from the same pass.
    detach_mean=True is the learned-variance term of the MSE objective (:896-906: the bound must not move the mean
    prediction); False is the KL objective (:865-872)."""

    @staticmethod
    def forward(ctx, out, x0, x_t, t, tab, T, mean_code, var_code, detach_mean, out_scale):
        L.require_cuda(out, x0, x_t, t)
        N = out.shape[0]
        chw = x0[0].numel() if N else 1
        o = out.contiguous()
        if o.dtype not in (th.float32, th.bfloat16):
            o = o.float()
        stride = o[0].numel() if N else chw
        code = L.F32 if o.dtype == th.float32 else L.BF16
        vb = th.empty(N, dtype=th.float32, device=out.device)
        g = None
        if out.requires_grad:
            g = th.zeros_like(o) if detach_mean else th.empty_like(o)
        if N:
            L.call("vaw_vb_terms", o.data_ptr(), code, stride, x0.data_ptr(), x_t.data_ptr(), t.data_ptr(),
                   tab.data_ptr(), T, vb.data_ptr(), L.ptr(g), stride, None, 1.0, mean_code, var_code,
                   1 if detach_mean else 0, float(out_scale), N, chw, L.stream_ptr())
        ctx.g, ctx.in_dtype, ctx.code = g, out.dtype, code
        return vb

    @staticmethod
    def backward(ctx, gv):
        g = ctx.g
        if g is None:
            return (None,) * 10
        res = th.empty_like(g)
        s = gv.float().contiguous()
        L.call("vaw_scale_rows", g.data_ptr(), s.data_ptr(), res.data_ptr(), ctx.code, g.shape[0], g[0].numel(),
               L.stream_ptr())
        return (res.to(ctx.in_dtype),) + (None,) * 9


def _f32(x0):
    return x0 if (x0.dtype == th.float32 and x0.is_contiguous()) else x0.float().contiguous()


class GaussianDiffusion:
    """Training-side GaussianDiffusion (reference :126-930).  Keyword-only constructor incl. the `args` namespace."""

    def __init__(self, *, args, betas, model_mean_type, model_var_type, loss_type, rescale_timesteps=False,
                 device="cuda"):
        self.args = args
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.loss_type = loss_type
        self.rescale_timesteps = rescale_timesteps
        self.mse_loss_weight_type = args.weight_type
        self.gamma = args.gamma
        self.learn_sigma = args.learn_sigma
        self.p2_gamma = args.p2_gamma
        self.p2_k = args.p2_k

        betas = np.array(betas, dtype=np.float64)
        self.betas = betas
        assert betas.ndim == 1, "betas must be 1-D"
        assert (betas >= 0).all() and (betas <= 1).all()
        self.num_timesteps = int(betas.shape[0])

        # float64 tables (reference :178-205)
        self.alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(self.alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(self.alphas) / (1.0 - self.alphas_cumprod)

        self._dev_tables = {}  # device -> (alpha, sigma, c0, c1, w_lut) fp32 tensors

    # ---- device tables: uploaded once per device instead of once per _extract call -----------------------
    def weight_lut(self):
        """float32 [T] host LUT of compute_mse_loss_weight over every timestep (vaw_loss_weight_lut)."""
        kind, k = _parse_weight_type(self.mse_loss_weight_type)
        lut = np.empty(self.num_timesteps, dtype=np.float32)
        a = np.ascontiguousarray(self.sqrt_alphas_cumprod)
        s = np.ascontiguousarray(self.sqrt_one_minus_alphas_cumprod)
        try:
            L.call("vaw_loss_weight_lut", a.ctypes.data, s.ctypes.data, self.num_timesteps,
                   self.model_mean_type.value, kind, float(k), float(self.p2_k), float(self.p2_gamma), lut.ctypes.data)
        except L.VawError as e:
            raise ValueError(f"Invalid mse_loss_weight_type: {self.mse_loss_weight_type}") from e
        return lut

    def _tables(self, device):
        key = str(device)
        tb = self._dev_tables.get(key)
        if tb is None:
            def up(a):
                return th.from_numpy(np.asarray(a, dtype=np.float64)).to(device).float().contiguous()
            ta, ts = up(self.sqrt_alphas_cumprod), up(self.sqrt_one_minus_alphas_cumprod)
            c0 = c1 = None
            if self.model_mean_type == ModelMeanType.PREVIOUS_X:
                c0, c1 = up(self.posterior_mean_coef1), up(self.posterior_mean_coef2)
            wt = th.from_numpy(self.weight_lut()).to(device)
            tb = (ta, ts, c0, c1, wt)
            self._dev_tables[key] = tb
        return tb

    # ---- reference API -------------------------------------------------------------------------------
    def _scale_timesteps(self, t):
        if self.rescale_timesteps:
            out = t.float() * (1000.0 / self.num_timesteps)
            hint = getattr(t, "_vaw_host_value", None)
            if hint is not None:
                out._vaw_host_value = float(hint) * (1000.0 / self.num_timesteps)
            return out
        return t

    def sample_t(self, x_start):
        if self.args.time_dist[0] == "uniform":
            return th.randint(0, self.num_timesteps, (x_start.shape[0],), device=x_start.device)
        raise NotImplementedError(f"Unknown time_dist: {self.args.time_dist}")

    def _k1(self, x_start, t, noise, want_target):
        L.require_cuda(x_start, t, noise)
        x0, eps = _f32(x_start), _f32(noise)
        ta, ts, c0, c1, _ = self._tables(x0.device)
        t64 = t.to(th.int64).contiguous()
        x_t = th.empty_like(x0)
        target = th.empty_like(x0) if want_target else None
        N = x0.shape[0]
        if x0.numel() == 0:
            return x_t, target
        L.call("vaw_qsample_target", x0.data_ptr(), eps.data_ptr(), t64.data_ptr(), ta.data_ptr(), ts.data_ptr(),
               L.ptr(c0), L.ptr(c1), x_t.data_ptr(), L.ptr(target), self.model_mean_type.value, N,
               x0[0].numel() if N else 1, L.stream_ptr())
        return x_t, target

    def q_sample(self, x_start, t, noise=None):
        if noise is None:
            noise = th.randn_like(x_start)
        assert noise.shape == x_start.shape
        return self._k1(x_start, t, noise, False)[0]

    def compute_target(self, x_start, noise, t, alpha=None, sigma=None):
        if self.model_mean_type == ModelMeanType.START_X:
            return x_start
        if self.model_mean_type == ModelMeanType.EPSILON:
            return noise
        return self._k1(x_start, t, noise, True)[1]

    def training_losses(self, model, x_start, features=None, t=None, model_kwargs=None, noise=None):
        """Reference :834-930.  Returns {"mse": [N], "loss": [N], ("align": scalar)} (fp32)."""
        if model_kwargs is None:
            model_kwargs = {}
        deferred = x_start if isinstance(x_start, DeferredLatent) else None
        philox = None
        if deferred is not None and (noise is not None or not deferred.is_cuda or th.cuda.is_current_stream_capturing()):
            x_start, deferred = deferred.tensor(), None       # explicit noise / capture: the ordinary two-step path
        if (noise is None and x_start.is_cuda and self.args.time_dist[0] == "uniform"
                and not th.cuda.is_current_stream_capturing()):
            # The noise (and the timesteps, and a deferred latent's eps1) are drawn INSIDE the kernels from the device
            # generator, in the reference's order (eps1, noise, t) and bit-identical to randn_like / randint
            philox = _PhiloxDraws(x_start.device)
            numel = int(np.prod(x_start.shape))
            off_latent = philox.take(numel) if deferred is not None else 0
            off_noise = philox.take(numel)
            N = x_start.shape[0]
            if t is None:
                t = th.empty(N, dtype=th.int64, device=x_start.device)
                if N:
                    L.call("vaw_randint_philox", philox.seed, philox.take(N), 0, self.num_timesteps, t.data_ptr(), N,
                           L.stream_ptr())
        else:
            if noise is None:
                noise = th.randn_like(x_start)  # RNG order as in the reference: noise first ...
            if t is None:
                t = self.sample_t(x_start)      # ... then the timesteps (:849-852)
        t64 = t.to(th.int64).contiguous()
        if philox is not None:
            learned_ = self.model_var_type in (ModelVarType.LEARNED, ModelVarType.LEARNED_RANGE)
            needs_x0 = self.model_mean_type not in (ModelMeanType.EPSILON, ModelMeanType.SCORE) or learned_ or self.loss_type.is_vb()
            ta, ts, c0, c1, _ = self._tables(x_start.device)
            chw = int(np.prod(x_start.shape[1:]))
            x_t = th.empty(tuple(x_start.shape), dtype=th.float32, device=x_start.device)
            if deferred is not None:
                lat = _f32(deferred.latent)
                x0 = th.empty_like(x_t) if needs_x0 else None
                src = (None, lat.data_ptr(), deferred.latent_scale)
            else:
                x0 = _f32(x_start)
                src = (x0.data_ptr(), None, 1.0)
            if x_t.numel():
                L.call("vaw_qsample_philox", src[0], src[1], src[2], philox.seed, off_latent, off_noise, t64.data_ptr(),
                       ta.data_ptr(), ts.data_ptr(), L.ptr(c0), L.ptr(c1),
                       x0.data_ptr() if (deferred is not None and x0 is not None) else None, None, x_t.data_ptr(), None,
                       self.model_mean_type.value, x_t.shape[0], chw, L.stream_ptr())
            eps = None
        else:
            assert noise.shape == x_start.shape
            x0, eps = _f32(x_start), _f32(noise)
            x_t, _ = self._k1(x0, t64, eps, False)

        if (self.args.learn_align and self.args.align_type == "mse" and th.is_tensor(features)
                and features.dtype == th.bfloat16 and getattr(getattr(model, "module", model), "supports_fused_align", False)):
            # REPA with an engine-backed DiT: the alignment loss is accumulated in the epilogue of the GEMM that produces
            # zs (north-star piece 5); compute_align_loss below then finds the value attached to zs
            model_kwargs = dict(model_kwargs, align_target=features)
        raw_output = model(x_t, self._scale_timesteps(t64), **model_kwargs)
        sec_out = None
        if isinstance(raw_output, tuple):
            model_output = raw_output[0]
            sec_out = raw_output[1] if len(raw_output) > 1 else None
        else:
            model_output = raw_output
        learned = self.model_var_type in (ModelVarType.LEARNED, ModelVarType.LEARNED_RANGE)
        B, C = x_t.shape[:2]
        if learned:
            assert model_output.shape == (B, C * 2, *x_t.shape[2:])
        else:
            assert model_output.shape == x_t.shape

        terms = {}
        if self.loss_type.is_vb():
            # LossType.KL / RESCALED_KL (:862-875): the bound itself is the loss, gradient through mean AND variance
            if self.model_mean_type.value > ModelMeanType.VELOCITY.value:
                raise NotImplementedError(self.model_mean_type)
            scale = float(self.num_timesteps) if self.loss_type == LossType.RESCALED_KL else 1.0
            terms["loss"] = _VbTerm.apply(model_output, x0, x_t, t64, self._reverse_table(x0.device), self.num_timesteps,
                                          self.model_mean_type.value, self.model_var_type.value, False, scale)
            return terms
        if self.loss_type not in (LossType.MSE, LossType.RESCALED_MSE):
            raise NotImplementedError(self.loss_type)
        if learned:
            # learn the variance with the variational bound without letting it move the mean prediction (:886-906)
            scale = self.num_timesteps / 1000.0 if self.loss_type == LossType.RESCALED_MSE else 1.0
            terms["vb"] = _VbTerm.apply(model_output, x0, x_t, t64, self._reverse_table(x0.device), self.num_timesteps,
                                        self.model_mean_type.value, self.model_var_type.value, True, scale)
        if philox is not None:
            terms["mse"] = _WeightedMSEPhilox.apply(model_output, x0, t64, self._tables(x_t.device),
                                                    self.model_mean_type.value, philox.seed, off_noise,
                                                    int(np.prod(x_t.shape[1:])))
        else:
            terms["mse"] = _WeightedMSE.apply(model_output, x0, eps, t64, self._tables(x0.device),
                                              self.model_mean_type.value)
        if self.args.learn_align:
            assert self.gamma > 0, "Gamma must be greater than 0 for align loss"
            terms["align"] = compute_align_loss(features, sec_out, self.args.align_type)
        if "vb" in terms:                       # :921-926, same precedence as the reference
            terms["loss"] = terms["mse"] + terms["vb"]
        elif self.args.learn_align:
            terms["loss"] = terms["mse"] + self.gamma * terms["align"]
        else:
            terms["loss"] = terms["mse"]
        return terms


    # ---- reverse process: one fused kernel per step instead of ~25 elementwise launches ---------
    def unpack_model_output(self, raw_output):
        """Reference :208-215: models may return (pred, aux, ...); sampling needs the prediction only."""
        return raw_output[0] if isinstance(raw_output, tuple) else raw_output

    def _reverse_table(self, device):
        """[VAW_RT_ROWS, T] fp32 device table (include/vaw_b200.h), every float64 row rounded once like
        _extract_into_tensor (:1059-1072)."""
        key = ("reverse", str(device))
        tb = self._dev_tables.get(key)
        if tb is None:
            zero = np.zeros_like(self.betas)
            if self.model_var_type == ModelVarType.FIXED_LARGE:      # :326-331
                var = np.append(self.posterior_variance[1], self.betas[1:])
                logvar = np.log(var)
            elif self.model_var_type == ModelVarType.FIXED_SMALL:
                var, logvar = self.posterior_variance, self.posterior_log_variance_clipped
            else:                                                     # LEARNED_RANGE: min_log (:319-321)
                var, logvar = zero, self.posterior_log_variance_clipped
            with np.errstate(divide="ignore", invalid="ignore"):
                rows = [self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod, self.sqrt_alphas_cumprod,
                        self.sqrt_one_minus_alphas_cumprod, 1.0 / self.posterior_mean_coef1,
                        self.posterior_mean_coef2 / self.posterior_mean_coef1, self.posterior_mean_coef1,
                        self.posterior_mean_coef2, logvar, np.log(self.betas), var, self.alphas_cumprod,
                        self.alphas_cumprod_prev, self.alphas_cumprod_next, self.posterior_log_variance_clipped]
            tb = th.from_numpy(np.stack([np.asarray(r, dtype=np.float64) for r in rows])).to(device).float().contiguous()
            self._dev_tables[key] = tb
        return tb

    def _model_output(self, model, x, t, model_kwargs):
        if model_kwargs is None:
            model_kwargs = {}
        B, C = x.shape[:2]
        assert t.shape == (B,)
        out = self.unpack_model_output(model(x, self._scale_timesteps(t), **model_kwargs))
        learned = self.model_var_type in (ModelVarType.LEARNED, ModelVarType.LEARNED_RANGE)
        assert out.shape == (B, C * 2 if learned else C, *x.shape[2:])
        if out.dtype not in (th.float32, th.bfloat16):
            out = out.float()
        return out.contiguous()

    def _reverse(self, mode, model, x, t, *, clip_denoised, denoised_fn, cond_fn, model_kwargs, eta=0.0, noise=None,
                 want=("sample", "pred_xstart")):
        if denoised_fn is not None or cond_fn is not None:
            raise NotImplementedError("denoised_fn / cond_fn hooks are outside the fused reverse step")
        if self.model_mean_type not in (ModelMeanType.PREVIOUS_X, ModelMeanType.START_X, ModelMeanType.EPSILON,
                                        ModelMeanType.VELOCITY):
            raise NotImplementedError(self.model_mean_type)
        L.require_cuda(x, t)
        xf = _f32(x)
        t64 = t.to(th.int64).contiguous()
        out = self._model_output(model, xf, t64, model_kwargs)
        need_noise = mode == L.RS_DDPM or (mode == L.RS_DDIM and eta != 0.0)
        if mode in (L.RS_DDPM, L.RS_DDIM) and noise is None:
            noise = th.randn_like(xf)        # drawn even when unused, so the RNG stream matches the reference's
        nz = _f32(noise) if need_noise else None
        if nz is not None:
            L.require_cuda(nz)
            assert nz.shape == xf.shape
        res = {k: th.empty_like(xf) for k in want}
        N = xf.shape[0]
        if xf.numel():
            tab = self._reverse_table(xf.device)
            L.call("vaw_reverse_step", out.data_ptr(), L.BF16 if out.dtype == th.bfloat16 else L.F32,
                   out[0].numel(), xf.data_ptr(), L.ptr(nz), t64.data_ptr(),
                   tab.data_ptr(), self.num_timesteps, L.ptr(res.get("sample")), L.ptr(res.get("pred_xstart")),
                   L.ptr(res.get("mean")), L.ptr(res.get("log_variance")), L.ptr(res.get("variance")),
                   self.model_mean_type.value, self.model_var_type.value, mode, float(eta), 1 if clip_denoised else 0,
                   N, xf[0].numel(), L.stream_ptr())
        return res

    def p_mean_variance(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None):
        """Reference :278-384 -> {"mean", "variance", "log_variance", "pred_xstart"} (fp32, shape of x)."""
        return self._reverse(L.RS_MOMENTS, model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                             cond_fn=None, model_kwargs=model_kwargs,
                             want=("mean", "variance", "log_variance", "pred_xstart"))

    def p_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None):
        """Reference :455-506 -> {"sample", "pred_xstart"}."""
        return self._reverse(L.RS_DDPM, model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                             cond_fn=cond_fn, model_kwargs=model_kwargs)

    def ddim_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None, eta=0.0):
        """Reference :603-651 -> {"sample", "pred_xstart"}."""
        return self._reverse(L.RS_DDIM, model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                             cond_fn=cond_fn, model_kwargs=model_kwargs, eta=eta)

    def ddim_reverse_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None, eta=0.0):
        """Reference :653-689 (deterministic reverse ODE)."""
        assert eta == 0.0, "Reverse ODE only for deterministic path"
        return self._reverse(L.RS_DDIM_REVERSE, model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                             cond_fn=None, model_kwargs=model_kwargs)

    def _loop(self, step, model, shape, noise, device, progress, **kw):
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        img = noise if noise is not None else th.randn(*shape, device=device)
        indices = list(range(self.num_timesteps))[::-1]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        for i in indices:
            t = th.full((shape[0],), i, dtype=th.int64, device=device)
            t._vaw_host_value = i     # known on the host: IntervalCFG tests its interval without a device round trip
            with th.no_grad():
                out = step(model, img, t, **kw)
                yield out
                img = out["sample"]

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                  model_kwargs=None, device=None, progress=False):
        """Reference :555-601."""
        yield from self._loop(self.p_sample, model, shape, noise, device, progress, clip_denoised=clip_denoised,
                              denoised_fn=denoised_fn, cond_fn=cond_fn, model_kwargs=model_kwargs)

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                      model_kwargs=None, device=None, progress=False):
        """Reference :508-553."""
        final = None
        for final in self.p_sample_loop_progressive(model, shape, noise, clip_denoised, denoised_fn, cond_fn,
                                                    model_kwargs, device, progress):
            pass
        return final["sample"]

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None,
                                     cond_fn=None, model_kwargs=None, device=None, progress=False, eta=0.0):
        """Reference :725-774."""
        yield from self._loop(self.ddim_sample, model, shape, noise, device, progress, clip_denoised=clip_denoised,
                              denoised_fn=denoised_fn, cond_fn=cond_fn, model_kwargs=model_kwargs, eta=eta)

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, eta=0.0):
        """Reference :691-723."""
        final = None
        for final in self.ddim_sample_loop_progressive(model, shape, noise, clip_denoised, denoised_fn, cond_fn,
                                                       model_kwargs, device, progress, eta):
            pass
        return final["sample"]


class FlowMatching:
    """Flow-matching objective (reference :1151-1340): continuous t in (0,1), analytic interpolant; same K1/K2 with
    per-sample coefficient arrays instead of per-timestep tables."""

    def __init__(self, *, args, model_mean_type, device="cuda"):
        self.args = args
        self.model_mean_type = model_mean_type
        self.mse_loss_weight_type = args.weight_type
        self.path_type = args.path_type
        self.sampler_type = getattr(args, "sampler_type", None)
        self.p2_gamma = args.p2_gamma
        self.p2_k = args.p2_k
        self.gamma = args.gamma
        self.learn_sigma = args.learn_sigma

    def interpolant(self, t):
        if self.path_type == "linear":
            return 1 - t, t, th.full_like(t, -1.0), th.full_like(t, 1.0)
        if self.path_type == "cosine":
            a, s = th.cos(t * np.pi / 2), th.sin(t * np.pi / 2)
            return a, s, -np.pi / 2 * s, np.pi / 2 * a
        if self.path_type == "linear_logsnr":
            lam = 10 + t * (-10.0 - 10)
            a, s = th.sigmoid(0.5 * lam), th.sigmoid(-0.5 * lam)
            da = -10.0 * a * s
            return a, s, da, -da
        raise NotImplementedError()

    def sample_t(self, x_start):
        kind = self.args.time_dist[0]
        if kind == "uniform":
            return th.rand(x_start.shape[0], device=x_start.device)
        if kind == "lognorm":
            mu, sigma = float(self.args.time_dist[-2]), float(self.args.time_dist[-1])
            return th.sigmoid(th.randn(x_start.shape[0], device=x_start.device) * sigma + mu)
        raise NotImplementedError(f"Unknown time_dist: {self.args.time_dist}")

    def _coef(self, t):
        a, s, da, ds = self.interpolant(t.float())
        c0 = c1 = None
        if self.model_mean_type == ModelMeanType.VECTOR:
            c0, c1 = da.contiguous(), ds.contiguous()
        return a.contiguous(), s.contiguous(), c0, c1

    def q_sample(self, x_start, noise, t):
        L.require_cuda(x_start, noise, t)
        x0, eps = _f32(x_start), _f32(noise)
        a, s, c0, c1 = self._coef(t)
        x_t = th.empty_like(x0)
        L.call("vaw_qsample_target", x0.data_ptr(), eps.data_ptr(), None, a.data_ptr(), s.data_ptr(), L.ptr(c0),
               L.ptr(c1), x_t.data_ptr(), None, self.model_mean_type.value, x0.shape[0], x0[0].numel(), L.stream_ptr())
        return x_t

    def compute_target(self, x_start, noise, t, alpha_t=None, sigma_t=None, d_alpha_t=None, d_sigma_t=None):
        if self.model_mean_type == ModelMeanType.START_X:
            return x_start
        if self.model_mean_type == ModelMeanType.EPSILON:
            return noise
        x0, eps = _f32(x_start), _f32(noise)
        a, s, c0, c1 = self._coef(t)
        x_t, tgt = th.empty_like(x0), th.empty_like(x0)
        L.call("vaw_qsample_target", x0.data_ptr(), eps.data_ptr(), None, a.data_ptr(), s.data_ptr(), L.ptr(c0),
               L.ptr(c1), x_t.data_ptr(), tgt.data_ptr(), self.model_mean_type.value, x0.shape[0], x0[0].numel(),
               L.stream_ptr())
        return tgt

    # ---- sampling (reference :1343-1418) ---------------------------------------------------------------------
    def expand_t_like_x(self, t, x):
        if t.dim() == 0:
            t = t.expand(x.shape[0])
        return t.view(t.size(0), *([1] * (x.dim() - 1))).to(x)

    def forward_model(self, model, sample_tensor, time_tensor, **model_kwargs):
        raw_output = model(sample_tensor, time_tensor.view(sample_tensor.shape[0]), **model_kwargs)
        return raw_output[0] if isinstance(raw_output, tuple) else raw_output

    def compute_diffusion(self, time_tensor):
        _, sigma_t, _, d_sigma_t = self.interpolant(time_tensor)
        return 2 * sigma_t * d_sigma_t

    def _step_coef(self, time_scalar):
        """HOST float32 coefficients of one sampling time (the reference evaluates the same torch ops on [N,1,1,1] copies
        of the scalar): alpha, sigma, d_alpha, d_sigma, 2 sigma d_sigma."""
        t32 = time_scalar.detach().cpu().to(th.float32).reshape(1)
        a, s, da, ds = self.interpolant(t32)
        return np.array([float(a), float(s), float(da), float(ds), float(2 * s * ds)], dtype=np.float32)

    def sde_sample(self, model, noise, device, num_steps=50, solver="heun", **model_kwargs):
        """Reference :1370-1408 (Euler-Maruyama / stochastic Heun, last step noise-free): one fused kernel per drift
        evaluation (vaw_flow_sde_step) instead of ~30 elementwise launches."""
        if solver not in ("euler", "heun"):
            raise ValueError(f"Unknown solver: {solver}")
        if self.model_mean_type.value < ModelMeanType.START_X.value or self.model_mean_type.value > ModelMeanType.VECTOR.value:
            raise NotImplementedError("Unsupported model_mean_type for vector")
        L.require_cuda(noise)
        timesteps = th.cat([th.linspace(1.0, 0.04, num_steps, dtype=th.float64), th.tensor([0.0], dtype=th.float64)])
        x = _f32(noise)
        N, n, mt = x.shape[0], x.numel(), self.model_mean_type.value

        def drift_step(x_eval, t_scalar, mode, x_base, step, sq, rnd=None, d_prev=None, want_drift=False, nscale=0.0):
            tt = th.full((N,), float(t_scalar.to(th.float32)), dtype=th.float32, device=x.device)
            mo = self.forward_model(model, x_eval, tt, **model_kwargs)
            if mo.dtype not in (th.float32, th.bfloat16):
                mo = mo.float()
            mo = mo.contiguous()
            coef = self._step_coef(t_scalar)
            x_out = th.empty_like(x)
            d_out = th.empty_like(x) if want_drift else None
            L.call("vaw_flow_sde_step", mo.data_ptr(), L.BF16 if mo.dtype == th.bfloat16 else L.F32, x_eval.data_ptr(),
                   coef.ctypes.data, mt, mode, x_base.data_ptr(), L.ptr(d_prev), L.ptr(rnd), float(step), float(nscale),
                   float(sq), x_out.data_ptr(), L.ptr(d_out), n, L.stream_ptr())
            return x_out, d_out

        with th.no_grad():
            for cur, nxt in zip(timesteps[:-2], timesteps[1:-1]):
                # the reference multiplies fp32 tensors by 0-dim float64 tensors: the scalar is rounded to fp32 first
                step = (nxt - cur).to(th.float32)
                sq = th.sqrt(th.abs(nxt - cur)).to(th.float32)
                rnd = th.randn_like(x)
                # th.sqrt(diffusion) of the CURRENT time scales the noise term of both Heun stages (:1387); a negative
                # coefficient (cosine path at t = 1 in fp32) gives NaN, like the reference
                nscale = float(th.sqrt(th.tensor(self._step_coef(cur)[4])))
                if solver == "euler":
                    x, _ = drift_step(x, cur, 0, x, step, sq, rnd, nscale=nscale)
                else:
                    pred, d_cur = drift_step(x, cur, 1, x, step, sq, rnd, want_drift=True, nscale=nscale)
                    x, _ = drift_step(pred, nxt, 2, x, step, sq, rnd, d_prev=d_cur, nscale=nscale)
            cur, nxt = timesteps[-2], timesteps[-1]
            x, _ = drift_step(x, cur, 0, x, (nxt - cur).to(th.float32), 0.0)
        return x

    def ode_sample(self, model, noise, device, num_steps=50, solver="dopri5", **model_kwargs):
        """Reference :1355-1363 integrates the probability-flow ODE with torchdiffeq's adaptive solvers and reads
        self.rtol / self.atol, which its constructor never sets: the call raises AttributeError in the reference itself.
        torchdiffeq is not part of this path; use sampler_type='sde'."""
        raise NotImplementedError("FlowMatching.ode_sample needs torchdiffeq (and fails in the reference: self.rtol / "
                                  "self.atol are never set); use sampler_type='sde'")

    def sample(self, model, noise, device, num_steps=50, solver="heun", **model_kwargs):
        """Reference :1411-1418."""
        if self.sampler_type == "ode":
            return self.ode_sample(model, noise, device, num_steps, solver=solver, **model_kwargs)
        if self.sampler_type == "sde":
            return self.sde_sample(model, noise, device, num_steps, solver=solver, **model_kwargs)
        raise NotImplementedError(f"Unsupported sampler_type: {self.sampler_type}")

    def training_losses(self, model, x_start, features=None, t=None, model_kwargs=None, noise=None):
        if model_kwargs is None:
            model_kwargs = {}
        if noise is None:
            noise = th.randn_like(x_start)
        if t is None:
            t = self.sample_t(x_start)
        x0, eps = _f32(x_start), _f32(noise)
        a, s, c0, c1 = self._coef(t)
        w = compute_mse_loss_weight(self.model_mean_type, self.mse_loss_weight_type, t, a, s, self.p2_k, self.p2_gamma)
        w = w.float().contiguous()
        x_t = self.q_sample(x0, eps, t)
        raw_output = model(x_t, t, **model_kwargs)
        sec_out = None
        if isinstance(raw_output, tuple):
            model_output = raw_output[0]
            sec_out = raw_output[1] if len(raw_output) > 1 else None
        else:
            model_output = raw_output
        assert model_output.shape == x_start.shape
        terms = {"mse": _WeightedMSE.apply(model_output, x0, eps, None, (a, s, c0, c1, w), self.model_mean_type.value)}
        if self.args.learn_align:
            assert self.gamma > 0, "Gamma must be greater than 0 for align loss"
            terms["align"] = compute_align_loss(features, sec_out, self.args.align_type)
            terms["loss"] = terms["mse"] + self.gamma * terms["align"]
        else:
            terms["loss"] = terms["mse"]
        return terms


def default_args(**over):
    """The subset of the reference's argparse namespace (main.py:36-135) that the objective reads."""
    d = dict(weight_type="lambda", gamma=0.5, learn_sigma=False, p2_gamma=1.0, p2_k=1.0, time_dist=["uniform"],
             learn_align=False, align_type="mse", amp=False, path_type="linear", sampler_type="ode")
    d.update(over)
    return SimpleNamespace(**d)


def create_gaussian_diffusion(*, steps=1000, noise_schedule="linear", mean_type="epsilon", var_type="fixed_large",
                              loss_type="mse", rescale_timesteps=True, device="cuda", **arg_overrides):
    """Alias named by the north-star text: builds the `args` namespace and calls the reference-shaped constructor
    (what main.py:224-245 build_diffusion does)."""
    args = default_args(**arg_overrides)
    return GaussianDiffusion(args=args, betas=get_named_beta_schedule(noise_schedule, steps),
                             model_mean_type=ModelMeanType[mean_type.upper()],
                             model_var_type=ModelVarType[var_type.upper()], loss_type=LossType[loss_type.upper()],
                             rescale_timesteps=rescale_timesteps, device=device)
