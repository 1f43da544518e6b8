"""ORACLE (test infrastructure, not product code): one CPU training step of the reference's path, end to end.

Composes the restatements in oracle/diffusion.py, oracle/resample.py and oracle/dit.py the way the reference's
trainer composes the originals (tools/trainer.py:55-58,104-135 with the sampler wired in the upstream way,
SURVEY.md D3):

    t, w   = sampler.sample(B)                                  resample.py:43-59
    terms  = diffusion.training_losses(model, x, t=t, y=...)    gaussian_diffusion.py:834-930
    sampler.update_with_all_losses(t, terms["loss"])            resample.py:151-159
    (terms["loss"] * w).mean().backward(); AdamW.step()         trainer.py:107-133, main.py:354

Used (a) as the checker of the GPU step in tests/ and smoke(), (b) as the timed CPU baseline / `--impl reference`
arm of bench.py (kind "port": the reference itself cannot travel to the GPU box).  fp32, torch CPU kernels, all host
threads.  Never imported by the product.
"""
from __future__ import annotations

import numpy as np
import torch

from . import diffusion as odiff
from . import resample as ors
from .dit import dit_forward

DIT_CONFIGS = {  # models/dit.py:361-375
    "DiT-S": dict(hidden=384, depth=12, heads=6),
    "DiT-B": dict(hidden=768, depth=12, heads=12),
    "DiT-L": dict(hidden=1024, depth=24, heads=16),
    "DiT-XL": dict(hidden=1152, depth=28, heads=16),
}


def init_dit_state(hidden, depth, heads, *, image_size=32, patch_size=2, in_channels=4, num_classes=1000, seed=0,
                   learn_align=False, z_dims=768, projector_dim=2048, mlp_ratio=4):
    """Random DiT weights with the reference's names/shapes (values N(0, 0.02); the CPU step's cost does not depend
    on them and parity tests copy real weights in)."""
    g = torch.Generator().manual_seed(seed)
    D, T = hidden, (image_size // patch_size) ** 2
    ppc = patch_size * patch_size * in_channels
    shapes = {
        "pos_embed": (1, T, D), "x_embedder.proj.weight": (D, in_channels, patch_size, patch_size),
        "x_embedder.proj.bias": (D,), "t_embedder.mlp.0.weight": (D, 256), "t_embedder.mlp.0.bias": (D,),
        "t_embedder.mlp.2.weight": (D, D), "t_embedder.mlp.2.bias": (D,),
        "y_embedder.embedding_table.weight": (num_classes, D),
        "final_layer.linear.weight": (ppc, D), "final_layer.linear.bias": (ppc,),
        "final_layer.adaLN_modulation.1.weight": (2 * D, D), "final_layer.adaLN_modulation.1.bias": (2 * D,),
    }
    for i in range(depth):
        p = f"blocks.{i}."
        shapes.update({p + "attn.qkv.weight": (3 * D, D), p + "attn.qkv.bias": (3 * D,),
                       p + "attn.proj.weight": (D, D), p + "attn.proj.bias": (D,),
                       p + "mlp.fc1.weight": (mlp_ratio * D, D), p + "mlp.fc1.bias": (mlp_ratio * D,),
                       p + "mlp.fc2.weight": (D, mlp_ratio * D), p + "mlp.fc2.bias": (D,),
                       p + "adaLN_modulation.1.weight": (6 * D, D), p + "adaLN_modulation.1.bias": (6 * D,)})
    if learn_align:
        shapes.update({"projectors.0.weight": (projector_dim, D), "projectors.0.bias": (projector_dim,),
                       "projectors.2.weight": (projector_dim, projector_dim), "projectors.2.bias": (projector_dim,),
                       "projectors.4.weight": (z_dims, projector_dim), "projectors.4.bias": (z_dims,)})
    return {k: (torch.randn(s, generator=g) * 0.02) for k, s in shapes.items()}


class OracleTrainer:
    def __init__(self, model="DiT-XL", *, schedule="cosine", mean_type="EPSILON", weight_type="lambda",
                 sampler="loss-second-moment", lr=1e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0, seed=0,
                 patch_size=2, state=None, prefill_history=True):
        cfg = DIT_CONFIGS[model]
        self.cfg, self.patch_size = cfg, patch_size
        self.tb = odiff.tables(odiff.named_beta_schedule(schedule, 1000))
        self.mean_type, self.weight_type = mean_type, weight_type
        sd = state if state is not None else init_dit_state(cfg["hidden"], cfg["depth"], cfg["heads"], seed=seed)
        self.sd = {k: v.clone().float().requires_grad_(k != "pos_embed") for k, v in sd.items()}
        self.opt = torch.optim.AdamW([v for k, v in self.sd.items() if v.requires_grad], lr=lr, betas=betas, eps=eps,
                                     weight_decay=weight_decay)
        self.sampler = sampler
        self.hist = np.zeros((1000, 10), dtype=np.float64)
        self.counts = np.zeros(1000, dtype=int)
        if sampler == "loss-second-moment" and prefill_history:
            self.hist, self.counts = synthetic_history(seed)

    def weights(self):
        if self.sampler == "uniform":
            return np.ones(1000, dtype=np.float64)
        return ors.second_moment_weights(self.hist, self.counts)

    def step(self, x0, y, noise=None, history_losses=None):
        """x0 [B,C,H,W] fp32 CPU tensor, y [B] int64.  Returns (loss scalar, terms, t, importance weights).
        history_losses: optional per-sample losses to record in the sampler history instead of this step's own
        (lets a parity test keep the two sampler states identical when the compared model runs in bf16)."""
        B = x0.shape[0]
        idx, iw = ors.sample(self.weights(), B)
        t = torch.from_numpy(idx)
        if noise is None:
            noise = torch.randn_like(x0)
        model_fn = lambda xt, ts: dit_forward(self.sd, xt, ts, y, patch_size=self.patch_size,
                                              num_heads=self.cfg["heads"], depth=self.cfg["depth"])
        terms = odiff.training_losses_torch(self.tb, self.mean_type, self.weight_type, model_fn, x0, t, noise)
        if self.sampler == "loss-second-moment":
            rec = terms["loss"].detach().tolist() if history_losses is None else list(history_losses)
            ors.update_history(self.hist, self.counts, idx.tolist(), rec)
        loss = (terms["loss"] * torch.from_numpy(iw)).mean()
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return loss.detach(), terms, t, iw


def synthetic_history(seed=0, T=1000, H=10):
    """Deterministic warmed-up loss history (every timestep has H entries) so that benchmarks and tests exercise
    the non-uniform branch of LossSecondMomentResampler.weights (resample.py:145-149)."""
    rng = np.random.RandomState(1234 + seed)
    base = 0.02 + 0.5 * np.exp(-np.arange(T) / 300.0)
    hist = np.abs(base[:, None] * (1.0 + 0.1 * rng.randn(T, H))).astype(np.float32).astype(np.float64)
    return hist, np.full(T, H, dtype=int)
