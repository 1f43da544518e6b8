"""Generate the golden fixtures in tests/golden/ by EXECUTING the reference (/root/reference, read-only) on CPU.

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

The reference is imported with the small import stubs in oracle/ref_stubs (timm / diffusers / torchdiffeq are not
installed; see oracle/ref_stubs/README.md).  Everything written here is an input/output pair of reference code:
    diffusion_golden.npz   schedules, tables, compute_mse_loss_weight over all t, q_sample / compute_target,
                           training_losses (+ autograd gradient w.r.t. the model output)
    sampler_golden.npz     LossSecondMomentResampler.weights / sample / update_with_all_losses, UniformSampler.sample
    unet_golden.npz        two tiny UNets (both attention orders / conditioning styles): weights, forward, losses, grads
    unet_shapes.json       state-dict names and shapes of every UNet factory (built on the meta device)
    reverse_golden.npz     p_mean_variance / p_sample / ddim_sample / ddim_reverse_sample over every mean / variance type
                           (fp32 and bf16 model outputs, pinned noise), IntervalCFG combine
    vit_golden.npz         the MoCo-v3 ViT teacher (tiny), preprocess_raw_image, ViT-B/16 position embedding and names
    dit_golden.npz         a tiny DiT (with REPA projector): weights, forward outputs, full training_losses + grads
    vb_golden.npz          training_losses with a learned variance (LEARNED / LEARNED_RANGE x MSE / RESCALED_MSE / KL /
                           RESCALED_KL): terms + the gradient w.r.t. the 2C-channel model output
    edm_golden.npz         tools/cfg_edm.py: Net (sigma table u, round_sigma, preconditioning) and ablation_sampler (Euler /
                           Heun x every discretisation / schedule / scaling, churn with pinned noise) around a toy denoiser
    flow_golden.npz        FlowMatching (tools/gaussian_diffusion.py:1151-1418): interpolant, q_sample, compute_target,
                           training_losses (+ gradient) over every path type x prediction type, the vector / score
                           conversions, and sde_sample (euler / heun) with pinned noise
"""
import os
import sys
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle", "ref_stubs"), "/root/reference"]

import numpy as np
import torch

sys.path.insert(0, HERE)
from fill import fill_by_name, grad_digest   # noqa: E402

import tools.gaussian_diffusion as rgd   # noqa: E402  (reference)
import tools.resample as rrs             # noqa: E402  (reference)
import models.dit as rdit                # noqa: E402  (reference)
import models.uvit as ruvit              # noqa: E402  (reference)
import models.unet as runet              # noqa: E402  (reference)


def ref_args(**kw):
    d = dict(weight_type="lambda", gamma=0.5, learn_sigma=False, p2_gamma=1.0, p2_k=1.0, time_dist=["uniform"],
             learn_align=False, align_type="mse", amp=False)
    d.update(kw)
    return SimpleNamespace(**d)


def make_diffusion(schedule="linear", mean="EPSILON", **kw):
    return rgd.GaussianDiffusion(args=ref_args(**kw), betas=rgd.get_named_beta_schedule(schedule, 1000),
                                 model_mean_type=rgd.ModelMeanType[mean], model_var_type=rgd.ModelVarType.FIXED_LARGE,
                                 loss_type=rgd.LossType.MSE, rescale_timesteps=True, device="cpu")


WEIGHT_CASES = [
    ("EPSILON", "constant"), ("EPSILON", "lambda"), ("EPSILON", "min_snr_5"), ("EPSILON", "max_snr_1"),
    ("EPSILON", "debias"), ("EPSILON", "p2"), ("EPSILON", "min_debias"), ("EPSILON", "max_debias"),
    ("START_X", "trunc_snr"), ("START_X", "snr"), ("START_X", "inv_snr"), ("START_X", "min_snr_5"),
    ("START_X", "max_snr_2"), ("START_X", "lambda"), ("VELOCITY", "min_snr_5"), ("VELOCITY", "lambda"),
]


def diffusion_golden():
    out = {}
    for sched in ("linear", "cosine", "linear_logsnr"):
        d = make_diffusion(sched)
        out[f"betas_{sched}"] = d.betas
        out[f"sqrt_ac_{sched}"] = d.sqrt_alphas_cumprod
        out[f"sqrt_1mac_{sched}"] = d.sqrt_one_minus_alphas_cumprod
        out[f"pmc1_{sched}"] = d.posterior_mean_coef1
        out[f"pmc2_{sched}"] = d.posterior_mean_coef2
    t_all = torch.arange(1000)
    for sched in ("linear", "cosine"):
        d = make_diffusion(sched)
        alpha = rgd._extract_into_tensor(d.sqrt_alphas_cumprod, t_all, t_all.shape)
        sigma = rgd._extract_into_tensor(d.sqrt_one_minus_alphas_cumprod, t_all, t_all.shape)
        for mean, wt in WEIGHT_CASES:
            w = rgd.compute_mse_loss_weight(rgd.ModelMeanType[mean], wt, t_all, alpha.clone(), sigma.clone(), 1.0, 1.0)
            out[f"w_{sched}_{mean}_{wt}"] = w.float().numpy()
    g = torch.Generator().manual_seed(7)
    x0 = torch.randn(6, 3, 8, 8, generator=g).clamp(-1, 1)
    eps = torch.randn(6, 3, 8, 8, generator=g)
    t = torch.tensor([0, 1, 500, 998, 999, 37])
    model_out = torch.randn(6, 3, 8, 8, generator=g)
    out.update(x0=x0.numpy(), eps=eps.numpy(), t=t.numpy(), model_out=model_out.numpy())
    for mean, wt in (("EPSILON", "lambda"), ("START_X", "lambda"), ("VELOCITY", "min_snr_5"), ("PREVIOUS_X", "constant"),
                     ("EPSILON", "min_snr_5")):
        d = make_diffusion("linear", mean, weight_type=wt)
        out[f"xt_{mean}"] = d.q_sample(x0, t, eps).numpy()
        out[f"target_{mean}"] = d.compute_target(x0, eps, t).numpy()
        mo = model_out.clone().requires_grad_(True)
        terms = d.training_losses(lambda x, ts, **k: mo, x0, None, t=t, noise=eps)
        terms["loss"].mean().backward()
        out[f"mse_{mean}_{wt}"] = terms["mse"].detach().float().numpy()
        out[f"grad_{mean}_{wt}"] = mo.grad.numpy()
    # sample_from_latent (tools/trainer.py:21-25) with its randn_like pinned to a recorded tensor
    import tools.trainer as rtr   # noqa: E402  (reference)
    latent = torch.randn(5, 8, 4, 4, generator=g)
    latent[:, 4:] = latent[:, 4:].abs() * 0.3
    lat_eps = torch.randn(5, 4, 4, 4, generator=g)
    orig = torch.randn_like
    torch.randn_like = lambda *a, **k: lat_eps.clone()
    try:
        lat_out = rtr.sample_from_latent(latent, 0.18215)
    finally:
        torch.randn_like = orig
    out.update(latent=latent.numpy(), latent_eps=lat_eps.numpy(), latent_out=lat_out.numpy())
    np.savez_compressed(os.path.join(HERE, "diffusion_golden.npz"), **out)
    print("diffusion_golden.npz", len(out), "arrays")


def sampler_golden():
    out = {}
    d = make_diffusion("linear")
    rng = np.random.RandomState(11)
    # warmed-up sampler
    s = rrs.LossSecondMomentResampler(d)
    hist = np.abs(rng.randn(1000, 10)) * (0.02 + rng.rand(1000, 1))
    s._loss_history[:] = hist
    s._loss_counts[:] = 10
    out["hist"] = hist
    out["weights"] = s.weights()
    np.random.seed(2024)
    idx, w = s.sample(48, "cpu")
    out["sample_idx"], out["sample_w"] = idx.numpy(), w.numpy()
    out["next_uniform_after_sample"] = np.array([np.random.random_sample()])
    # cold sampler -> uniform branch
    s2 = rrs.LossSecondMomentResampler(d)
    out["weights_cold"] = s2.weights()
    np.random.seed(77)
    idx2, w2 = s2.sample(32, "cpu")
    out["cold_idx"], out["cold_w"] = idx2.numpy(), w2.numpy()
    # update sequence with duplicates, small T range so that rows wrap around
    ts = rng.randint(0, 25, size=400)
    losses = rng.rand(400).astype(np.float32)
    s3 = rrs.LossSecondMomentResampler(d)
    s3.update_with_all_losses(ts.tolist(), [float(x) for x in losses])
    out["upd_ts"], out["upd_losses"] = ts, losses
    out["upd_hist"], out["upd_counts"] = s3._loss_history.copy(), s3._loss_counts.copy()
    # uniform sampler
    u = rrs.UniformSampler(d)
    np.random.seed(5)
    ui, uw = u.sample(40, "cpu")
    out["uni_idx"], out["uni_w"] = ui.numpy(), uw.numpy()
    np.savez_compressed(os.path.join(HERE, "sampler_golden.npz"), **out)
    print("sampler_golden.npz", len(out), "arrays")


def dit_golden():
    torch.manual_seed(3)
    cfg = dict(image_size=8, patch_size=2, in_channels=4, hidden_size=64, depth=2, num_heads=1,
               class_dropout_prob=0.0, num_classes=10, learn_sigma=False, learn_align=True, encoder_depth=1,
               z_dims=16, projector_dim=32)
    m = rdit.DiT(**cfg)
    with torch.no_grad():
        for p in m.parameters():   # de-zero the adaLN-Zero / final-layer tensors
            if p.requires_grad and p.abs().sum() == 0:
                p.normal_(0, 0.02)
    m.train()
    g = torch.Generator().manual_seed(9)
    x0 = torch.randn(3, 4, 8, 8, generator=g)
    eps = torch.randn(3, 4, 8, 8, generator=g)
    t = torch.tensor([3, 500, 990])
    y = torch.tensor([1, 7, 4])
    feats = torch.randn(3, 16, 16, generator=g)
    out = {f"param::{k}": v.detach().numpy() for k, v in m.state_dict().items()}
    out.update(x0=x0.numpy(), eps=eps.numpy(), t=t.numpy(), y=y.numpy(), feats=feats.numpy())
    d = make_diffusion("cosine", "EPSILON", weight_type="lambda", learn_align=True, gamma=0.5)
    x_t = d.q_sample(x0, t, eps)
    o, zs = m(x_t, d._scale_timesteps(t), y)
    out["fwd_out"], out["fwd_zs"] = o.detach().numpy(), zs.detach().numpy()
    terms = d.training_losses(m, x0, feats, t=t, model_kwargs={"y": y}, noise=eps)
    terms["loss"].mean().backward()
    out["mse"], out["align"], out["loss"] = (terms[k].detach().numpy() for k in ("mse", "align", "loss"))
    for k, p in m.named_parameters():
        if p.grad is not None:
            out[f"grad::{k}"] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "dit_golden.npz"), **out)
    print("dit_golden.npz", len(out), "arrays")


def uvit_golden():
    torch.manual_seed(4)
    m = ruvit.UViT(image_size=8, patch_size=2, in_channels=4, embed_dim=64, depth=3, num_heads=1, mlp_ratio=4,
                   num_classes=10, class_dropout_prob=0.0)
    m.train()
    g = torch.Generator().manual_seed(10)
    x0 = torch.randn(3, 4, 8, 8, generator=g)
    eps = torch.randn(3, 4, 8, 8, generator=g)
    t = torch.tensor([7, 450, 985])
    y = torch.tensor([2, 9, 0])
    out = {f"param::{k}": v.detach().numpy() for k, v in m.state_dict().items()}
    out.update(x0=x0.numpy(), eps=eps.numpy(), t=t.numpy(), y=y.numpy())
    d = make_diffusion("linear", "EPSILON", weight_type="lambda")
    x_t = d.q_sample(x0, t, eps)
    out["fwd_out"] = m(x_t, d._scale_timesteps(t), y).detach().numpy()
    terms = d.training_losses(m, x0, None, t=t, model_kwargs={"y": y}, noise=eps)
    terms["loss"].mean().backward()
    out["mse"], out["loss"] = terms["mse"].detach().numpy(), terms["loss"].detach().numpy()
    for k, p in m.named_parameters():
        if p.grad is not None:
            out[f"grad::{k}"] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "uvit_golden.npz"), **out)
    print("uvit_golden.npz", len(out), "arrays")


UNET_CASES = {
    # the shipped style: scale-shift norm, ResBlock up/down, new attention order, class-conditional
    "a": dict(image_size=8, num_channels=32, num_res_blocks=1, channel_mult="1,2", in_channels=3, num_classes=10,
              class_cond=True, attention_resolutions="4", num_heads=2, use_scale_shift_norm=True, resblock_updown=True,
              use_new_attention_order=True),
    # the other branches: additive conditioning, conv down/up-sampling, legacy attention order, unconditional
    "b": dict(image_size=8, num_channels=32, num_res_blocks=1, channel_mult="1,2", in_channels=3, num_classes=10,
              class_cond=False, attention_resolutions="4,8", num_heads=1, num_head_channels=16,
              use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False),
}


def unet_golden():
    out = {}
    for tag, cfg in UNET_CASES.items():
        torch.manual_seed(5)
        m = runet.create_unet_model(**cfg)
        fill_by_name(m)   # name-keyed deterministic weights (tests/golden/fill.py): the tests refill their own model
        m.train()
        g = torch.Generator().manual_seed(12)
        x0 = torch.randn(3, 3, 8, 8, generator=g)
        eps = torch.randn(3, 3, 8, 8, generator=g)
        t = torch.tensor([5, 420, 977])
        y = torch.tensor([3, 8, 1])
        kw = {"y": y} if cfg["class_cond"] else {}
        out.update({f"{tag}::x0": x0.numpy(), f"{tag}::eps": eps.numpy(), f"{tag}::t": t.numpy(), f"{tag}::y": y.numpy()})
        d = make_diffusion("linear", "EPSILON", weight_type="min_snr_5")
        x_t = d.q_sample(x0, t, eps)
        out[f"{tag}::fwd_out"] = m(x_t, d._scale_timesteps(t), **kw).detach().numpy()
        terms = d.training_losses(m, x0, None, t=t, model_kwargs=kw, noise=eps)
        terms["loss"].mean().backward()
        out[f"{tag}::mse"], out[f"{tag}::loss"] = terms["mse"].detach().numpy(), terms["loss"].detach().numpy()
        for k, p in m.named_parameters():
            if p.grad is not None:
                out[f"{tag}::grad::{k}"] = grad_digest(p.grad)
    # parameter names/shapes of every shipped factory (checkpoint compatibility), no values
    shapes = {}
    for name, fn in runet.UNet_models.items():
        with torch.device("meta"):
            mm = fn()
        shapes[name] = {k: list(v.shape) for k, v in mm.state_dict().items()}
    import json
    with open(os.path.join(HERE, "unet_shapes.json"), "w") as f:
        json.dump(shapes, f, separators=(",", ":"))
    np.savez_compressed(os.path.join(HERE, "unet_golden.npz"), **out)
    print("unet_golden.npz", len(out), "arrays; unet_shapes.json", {k: len(v) for k, v in shapes.items()})


REVERSE_CASES = [  # (mean type, variance type, model-output dtype, clip_denoised)
    ("EPSILON", "FIXED_LARGE", "f32", True), ("EPSILON", "FIXED_SMALL", "f32", False),
    ("EPSILON", "FIXED_LARGE", "bf16", True), ("START_X", "FIXED_LARGE", "f32", True),
    ("PREVIOUS_X", "FIXED_SMALL", "f32", True), ("EPSILON", "LEARNED_RANGE", "f32", True),
    ("EPSILON", "LEARNED_RANGE", "bf16", True), ("EPSILON", "LEARNED", "f32", True),
    ("START_X", "LEARNED", "bf16", False),
]


RESPACE_CASES = [(1000, "ddim10"), (1000, "ddim50"), (1000, "ddim250"), (300, "10,15,20"), (1000, [1000]),
                 (1000, "250"), (100, "3,1,7"), (1000, [1, 1, 1]), (999, "37,2")]


def reverse_golden():
    import tools.sampler as rsm   # noqa: E402  (reference)
    out = {}
    g = torch.Generator().manual_seed(21)
    x = torch.randn(6, 3, 8, 8, generator=g) * 1.5
    mo2 = torch.randn(6, 6, 8, 8, generator=g)
    noise = torch.randn(6, 3, 8, 8, generator=g)
    t = torch.tensor([0, 1, 500, 998, 999, 37])
    out.update(x=x.numpy(), model_out2=mo2.numpy(), noise=noise.numpy(), t=t.numpy())
    orig = torch.randn_like
    torch.randn_like = lambda a, **k: noise[:a.shape[0]].clone()
    try:
        for mean, var, dt, clip in REVERSE_CASES:
            d = rgd.GaussianDiffusion(args=ref_args(), betas=rgd.get_named_beta_schedule("linear", 1000),
                                      model_mean_type=rgd.ModelMeanType[mean], model_var_type=rgd.ModelVarType[var],
                                      loss_type=rgd.LossType.MSE, rescale_timesteps=True, device="cpu")
            mo = mo2 if var.startswith("LEARNED") else mo2[:, :3].contiguous()
            if dt == "bf16":
                mo = mo.bfloat16()
            model = lambda xx, ts, **k: mo
            key = f"{mean}_{var}_{dt}_{int(clip)}"
            pmv = d.p_mean_variance(model, x, t, clip_denoised=clip)
            for k in ("mean", "variance", "log_variance", "pred_xstart"):
                out[f"pmv_{k}::{key}"] = pmv[k].float().numpy()
            out[f"p_sample::{key}"] = d.p_sample(model, x, t, clip_denoised=clip)["sample"].float().numpy()
            out[f"ddim0::{key}"] = d.ddim_sample(model, x, t, clip_denoised=clip, eta=0.0)["sample"].float().numpy()
            out[f"ddim7::{key}"] = d.ddim_sample(model, x, t, clip_denoised=clip, eta=0.7)["sample"].float().numpy()
            out[f"ddimrev::{key}"] = d.ddim_reverse_sample(model, x, t, clip_denoised=clip)["sample"].float().numpy()
        # VELOCITY: the reference's _predict_xstart_from_v broadcasts its coefficient over the last axis, so it only
        # means "per sample" for a batch of one
        d = make_diffusion("cosine", "VELOCITY")
        for i in (1, 2, 4):
            model = lambda xx, ts, **k: mo2[i:i + 1, :3]
            out[f"vel_ddim0::{i}"] = d.ddim_sample(model, x[i:i + 1], t[i:i + 1], eta=0.0)["sample"].numpy()
            out[f"vel_xs::{i}"] = d.p_mean_variance(model, x[i:i + 1], t[i:i + 1])["pred_xstart"].numpy()
    finally:
        torch.randn_like = orig
    # IntervalCFG (tools/sampler.py:10-48): doubled batch through a recorded "model", guidance inside / outside the interval
    y = torch.tensor([1, 2, 3])
    both = torch.randn(6, 3, 8, 8, generator=g)
    calls = []

    class Rec(torch.nn.Module):
        def __init__(self, dt):
            super().__init__()
            self.dt = dt

        def forward(self, xx, tt, **kw):
            calls.append((tuple(xx.shape), kw["y"].tolist()))
            return (both if xx.shape[0] == 6 else both[:3]).to(self.dt)

    for dt, name in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        cfg = rsm.IntervalCFG(Rec(dt), num_classes=10, guidance_scale=2.5, interval=(100.0, 600.0))
        out[f"cfg_in::{name}"] = cfg(x[:3], torch.tensor([300.0] * 3), y=y).float().numpy()
        out[f"cfg_out::{name}"] = cfg(x[:3], torch.tensor([800.0] * 3), y=y).float().numpy()
    out["cfg_both"] = both.numpy()
    # respacing (tools/respace.py): kept timesteps, rebuilt betas, and what the wrapped model is handed
    import tools.respace as rrp   # noqa: E402  (reference)
    for n, spec in RESPACE_CASES:
        out[f"space::{n}::{spec}"] = np.array(sorted(rrp.space_timesteps(n, spec)))
    for spec in ("ddim10", "ddim50", "10,15,20"):
        sd = rrp.SpacedDiffusion(use_timesteps=rrp.space_timesteps(1000, spec), args=ref_args(),
                                 betas=rgd.get_named_beta_schedule("cosine", 1000),
                                 model_mean_type=rgd.ModelMeanType.EPSILON, model_var_type=rgd.ModelVarType.FIXED_LARGE,
                                 loss_type=rgd.LossType.MSE, rescale_timesteps=True, device="cpu")
        out[f"spaced_betas::{spec}"] = sd.betas
        out[f"spaced_map::{spec}"] = np.array(sd.timestep_map)
        seen = []
        sd.p_mean_variance(lambda xx, ts, **k: (seen.append(ts.clone()), xx)[1], x, torch.tensor([0, 1, 2, 3, 5, 9]))
        out[f"spaced_model_t::{spec}"] = seen[0].numpy()
    out["cfg_labels_seen"] = np.array(calls[0][1])
    np.savez_compressed(os.path.join(HERE, "reverse_golden.npz"), **out)
    print("reverse_golden.npz", len(out), "arrays")


def vit_golden():
    """The reference's VisionTransformerMoCo (encoders/mocov3_vit.py) + preprocess_raw_image (tools/align_utils.py) on a
    tiny configuration.  timm's VisionTransformer base is the restatement in oracle/ref_stubs (timm is not installed),
    so this pins the MoCo-specific parts: position embedding, parameter names, preprocessing, cls-token drop."""
    from functools import partial
    import encoders.mocov3_vit as rvit      # noqa: E402  (reference)
    import tools.align_utils as rau         # noqa: E402  (reference)
    torch.manual_seed(6)
    m = rvit.VisionTransformerMoCo(img_size=32, patch_size=8, embed_dim=128, depth=2, num_heads=2, mlp_ratio=4,
                                   qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=0)
    pos = m.pos_embed.detach().clone()
    fill_by_name(m)                      # name-keyed deterministic weights (tests refill their own model) ...
    with torch.no_grad():
        m.pos_embed.copy_(pos)           # ... except the fixed sin-cos table, which is the reference's own
    m.eval()
    g = torch.Generator().manual_seed(14)
    raw = torch.randint(0, 256, (3, 3, 32, 32), generator=g).float()
    out = {"pos_embed": pos.numpy(), "names": np.array(sorted(k for k in m.state_dict() if not k.startswith("head")))}
    out["raw"] = raw.numpy()
    out["pre"] = rau.preprocess_raw_image(raw, "mocov3-vit-b").numpy()
    with torch.no_grad():
        out["features"] = m.forward_features(torch.from_numpy(out["pre"]))[:, 1:].numpy()
    big = rvit.vit_base(num_classes=0)
    out["pos_embed_base_sub"] = big.pos_embed.detach().numpy().astype(np.float32)[:, ::8, ::16]
    out["base_names"] = np.array(sorted(k for k in big.state_dict() if not k.startswith("head")))
    out["base_shapes"] = np.array([str(tuple(big.state_dict()[k].shape)) for k in out["base_names"]])
    np.savez_compressed(os.path.join(HERE, "vit_golden.npz"), **out)
    print("vit_golden.npz", len(out), "arrays")


VB_CASES = [("EPSILON", "LEARNED_RANGE", "MSE", "lambda"), ("EPSILON", "LEARNED", "RESCALED_MSE", "min_snr_5"),
            ("START_X", "LEARNED_RANGE", "KL", "constant"), ("EPSILON", "LEARNED_RANGE", "RESCALED_KL", "constant"),
            ("PREVIOUS_X", "LEARNED", "KL", "constant"), ("START_X", "LEARNED_RANGE", "RESCALED_MSE", "lambda")]


def vb_golden():
    out = {}
    g = torch.Generator().manual_seed(41)
    x0 = torch.randn(6, 3, 8, 8, generator=g).clamp(-1, 1)
    x0[:, :, 0, :4] = -1.0          # pixels at the ends of the range take the one-sided branches of the decoder NLL
    x0[:, :, 1, :4] = 1.0
    eps = torch.randn(6, 3, 8, 8, generator=g)
    t = torch.tensor([0, 0, 1, 500, 998, 999])
    mo = torch.randn(6, 6, 8, 8, generator=g)
    mo[:, 3:] *= 0.5
    out.update(x0=x0.numpy(), eps=eps.numpy(), t=t.numpy(), model_out=mo.numpy())
    for sched in ("linear", "cosine"):
        for mean, var, loss, wt in VB_CASES:
            d = rgd.GaussianDiffusion(args=ref_args(weight_type=wt, learn_sigma=True),
                                      betas=rgd.get_named_beta_schedule(sched, 1000),
                                      model_mean_type=rgd.ModelMeanType[mean], model_var_type=rgd.ModelVarType[var],
                                      loss_type=rgd.LossType[loss], rescale_timesteps=True, device="cpu")
            m = mo.clone().requires_grad_(True)
            terms = d.training_losses(lambda x, ts, **k: m, x0, None, t=t, noise=eps)
            terms["loss"].mean().backward()
            key = f"{sched}::{mean}::{var}::{loss}::{wt}"
            for k in ("mse", "vb", "loss"):
                if k in terms:
                    out[f"{k}::{key}"] = terms[k].detach().float().numpy()
            out[f"grad::{key}"] = m.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "vb_golden.npz"), **out)
    print("vb_golden.npz", len(out), "arrays")


EDM_CASES = [  # (tag, Net kwargs, sampler kwargs)
    ("heun_edm_eps", dict(pred_type="EPSILON", noise_schedule="linear"), dict(solver="heun")),
    ("euler_edm_eps", dict(pred_type="EPSILON", noise_schedule="linear"), dict(solver="euler")),
    ("heun_edm_x0_cos", dict(pred_type="START_X", noise_schedule="cosine"), dict(solver="heun")),
    ("heun_edm_v_logsnr", dict(pred_type="VELOCITY", noise_schedule="linear_logsnr"), dict(solver="heun")),
    ("heun_vp", dict(pred_type="EPSILON", noise_schedule="cosine"), dict(solver="heun", discretization="vp", schedule="vp", scaling="vp")),
    ("heun_ve", dict(pred_type="EPSILON", noise_schedule="linear"), dict(solver="heun", discretization="ve", schedule="ve", scaling="none")),
    ("euler_iddpm", dict(pred_type="EPSILON", noise_schedule="cosine"), dict(solver="euler", discretization="iddpm", schedule="linear", scaling="none")),
    ("heun_churn", dict(pred_type="EPSILON", noise_schedule="linear"), dict(solver="heun", S_churn=20, S_min=0.05, S_max=50, S_noise=1.003)),
    ("heun_alpha", dict(pred_type="VELOCITY", noise_schedule="cosine"), dict(solver="heun", alpha=0.5)),
]


class _ToyDenoiser(torch.nn.Module):
    """Stands in for the (guided) denoiser inside Net: smooth, depends on x, the integer timestep and the label."""

    def forward(self, x, t, y=None, **kw):
        tt = torch.sin(t.float() / 100.0).view(-1, 1, 1, 1)
        yy = (y.float().view(-1, 1, 1, 1) / 10.0) if y is not None else 0.0
        return (0.3 * x + 0.05 * tt + 0.01 * yy).to(x.dtype)


def edm_golden():
    import warnings
    warnings.filterwarnings("ignore")
    import tools.cfg_edm as redm   # noqa: E402  (reference)
    out = {}
    g = torch.Generator().manual_seed(51)
    latents = torch.randn(4, 3, 8, 8, generator=g)
    labels = torch.tensor([1, 7, 3, 9])
    noises = torch.randn(12, 4, 3, 8, 8, generator=g, dtype=torch.float64)
    out.update(latents=latents.numpy(), labels=labels.numpy(), noises=noises.numpy())
    for sched in ("linear", "cosine", "linear_logsnr"):
        net = redm.Net(_ToyDenoiser(), img_resolution=8, img_channels=3, noise_schedule=sched)
        out[f"u::{sched}"] = net.u.numpy()
        out[f"sigma_minmax::{sched}"] = np.array([net.sigma_min, net.sigma_max])
        probe = torch.tensor([0.002, 0.01, 0.5, 1.0, 7.3, 80.0, 155.0], dtype=torch.float64)
        out[f"round_idx::{sched}"] = net.round_sigma(probe, return_index=True).numpy()
        out[f"round_val::{sched}"] = net.round_sigma(probe).numpy()
        x = torch.randn(4, 3, 8, 8, generator=g, dtype=torch.float64)
        out[f"net_x::{sched}"] = x.numpy()
        for pred in ("EPSILON", "START_X", "VELOCITY"):
            n2 = redm.Net(_ToyDenoiser(), img_resolution=8, img_channels=3, noise_schedule=sched, pred_type=pred)
            out[f"net_out::{sched}::{pred}"] = n2(x, torch.tensor(2.5, dtype=torch.float64), labels).numpy()
    for tag, nkw, skw in EDM_CASES:
        net = redm.Net(_ToyDenoiser(), img_resolution=8, img_channels=3, **nkw)
        it = iter(noises)
        res = redm.ablation_sampler(net, latents, class_labels=labels, randn_like=lambda a: next(it).to(a.dtype),
                                    num_steps=7, **skw)
        assert res.dtype == torch.float64
        out[f"sample::{tag}"] = res.numpy()
    np.savez_compressed(os.path.join(HERE, "edm_golden.npz"), **out)
    print("edm_golden.npz", len(out), "arrays")


FLOW_CASES = [("START_X", "lambda"), ("EPSILON", "lambda"), ("EPSILON", "min_snr_5"), ("VELOCITY", "lambda"),
              ("VELOCITY", "min_snr_5"), ("VECTOR", "lambda"), ("VECTOR", "constant"), ("SCORE", "constant")]


SDE_CASES = [("linear", "VECTOR"), ("linear", "VELOCITY"), ("linear", "START_X"), ("linear_logsnr", "VECTOR"),
             ("linear_logsnr", "VELOCITY"), ("linear_logsnr", "START_X"), ("linear_logsnr", "EPSILON"), ("cosine", "VECTOR")]


def flow_golden():
    out = {}
    g = torch.Generator().manual_seed(31)
    x0 = torch.randn(6, 3, 8, 8, generator=g).clamp(-1, 1)
    eps = torch.randn(6, 3, 8, 8, generator=g)
    model_out = torch.randn(6, 3, 8, 8, generator=g)
    t = torch.tensor([0.01, 0.2, 0.5, 0.77, 0.9, 0.99])
    out.update(x0=x0.numpy(), eps=eps.numpy(), model_out=model_out.numpy(), t=t.numpy())
    for path in ("linear", "cosine", "linear_logsnr"):
        for mean, wt in FLOW_CASES:
            fm = rgd.FlowMatching(args=ref_args(weight_type=wt, path_type=path, sampler_type="sde"),
                                  model_mean_type=rgd.ModelMeanType[mean], device="cpu")
            key = f"{path}::{mean}::{wt}"
            if wt in ("lambda", "constant") and mean in ("START_X", "VECTOR", "SCORE"):
                a, s_, da, ds = fm.interpolant(t)
                out[f"interp::{path}"] = torch.stack([a, s_, da, ds]).numpy()
                out[f"xt::{path}"] = fm.q_sample(x0, eps, t).numpy()
            out[f"target::{path}::{mean}"] = fm.compute_target(x0, eps, t).numpy()
            mo = model_out.clone().requires_grad_(True)
            terms = fm.training_losses(lambda x, ts, **k: mo, x0, None, t=t, noise=eps)
            terms["loss"].float().mean().backward()
            out[f"mse::{key}"] = terms["mse"].detach().float().numpy()
            out[f"grad::{key}"] = mo.grad.numpy()
        # model-output conversions used by the samplers (:1206-1257), t broadcast like sde_sample does
        tt = t.view(-1, 1, 1, 1)
        for mean in ("START_X", "EPSILON", "VELOCITY", "VECTOR", "SCORE"):
            fm = rgd.FlowMatching(args=ref_args(path_type=path, sampler_type="sde"),
                                  model_mean_type=rgd.ModelMeanType[mean], device="cpu")
            if mean != "SCORE":
                out[f"to_vector::{path}::{mean}"] = fm.convert_model_output_to_vector(model_out, x0, tt).numpy()
            out[f"to_score::{path}::{mean}"] = fm.convert_model_output_to_score(model_out, x0, tt).numpy()
    # sde_sample (:1370-1408) with the randn_like draws pinned and a linear toy denoiser; ode_sample (:1355-1363) needs
    # torchdiffeq and reads self.rtol / self.atol, which the class never sets, so it cannot be executed
    noises = torch.randn(16, 4, 3, 8, 8, generator=g)
    start = torch.randn(4, 3, 8, 8, generator=g)
    out.update(sde_noises=noises.numpy(), sde_start=start.numpy())
    orig = torch.randn_like
    # (the cosine path makes the reference return NaN: cos(fp32(pi/2)) < 0 puts a negative number under the sqrt of the
    # diffusion coefficient at t = 1; EPSILON on the linear path divides by alpha(1) = 0 — one such case is recorded)
    for path, mean in SDE_CASES:
        if True:
            for solver in ("euler", "heun"):
                it = iter(noises)
                torch.randn_like = lambda a, **k: next(it).to(a.dtype)
                try:
                    a_ = ref_args(path_type=path, sampler_type="sde")
                    fm = rgd.FlowMatching(args=a_, model_mean_type=rgd.ModelMeanType[mean], device="cpu")
                    toy = lambda x, tm, **k: (0.25 * x - 0.1 * tm.view(-1, 1, 1, 1).to(x.dtype)).float()
                    res = fm.sample(toy, start, "cpu", num_steps=6, solver=solver)
                finally:
                    torch.randn_like = orig
                out[f"sde::{path}::{mean}::{solver}"] = res.numpy()
    np.savez_compressed(os.path.join(HERE, "flow_golden.npz"), **out)
    print("flow_golden.npz", len(out), "arrays")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "edm":
        edm_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "vb":
        vb_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "flow":
        flow_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "vit":
        vit_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "reverse":
        reverse_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "diffusion":
        diffusion_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "unet":
        unet_golden()
        sys.exit(0)
    unet_golden()
    uvit_golden()
    diffusion_golden()
    sampler_golden()
    reverse_golden()
    vit_golden()
    dit_golden()
    flow_golden()
    vb_golden()
    edm_golden()
