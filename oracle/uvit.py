"""ORACLE (test infrastructure, not product code): plain-PyTorch restatement of the reference U-ViT forward.

Follows /root/reference/models/uvit.py — timestep_embedding (:21-39), unpatchify (:47-52), Attention (:55-93, the
'flash' branch: `.float()` then F.scaled_dot_product_attention), Block (:96-121), PatchEmbed (:124-136),
UViT.forward (:220-250) — and tools/timm.py:96-112 (Mlp).  Functional over a state_dict with the reference's
parameter names; pinned against the executed reference by tests/golden/make_golden.py -> uvit_golden.npz.
Only tests/, smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def timestep_embedding(timesteps, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _block(sd, pre, x, skip, num_heads):
    if skip is not None:
        x = F.linear(torch.cat([x, skip], dim=-1), sd[pre + "skip_linear.weight"], sd[pre + "skip_linear.bias"])
    B, Lt, Cd = x.shape
    D = Cd
    h = F.layer_norm(x, (D,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])
    qkv = F.linear(h, sd[pre + "attn.qkv.weight"])
    qkv = qkv.reshape(B, Lt, 3, num_heads, D // num_heads).permute(2, 0, 3, 1, 4).float()
    o = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
    o = o.permute(0, 2, 1, 3).reshape(B, Lt, D)
    x = x + F.linear(o, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])
    h = F.layer_norm(x, (D,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])
    h = F.linear(F.gelu(F.linear(h, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])),
                 sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
    return x + h


def uvit_forward(sd, x, timesteps, y, *, patch_size, num_heads, depth, conv=True):
    """sd: reference-named state dict; x [N,C,H,W]; timesteps [N]; y [N] long or None.  Returns [N,C,H,W]."""
    p = patch_size
    D = sd["pos_embed"].shape[-1]
    C = x.shape[1]
    h = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=p).flatten(2).transpose(1, 2)
    Lp = h.shape[1]
    tok = timestep_embedding(timesteps, D).unsqueeze(1)
    h = torch.cat((tok, h), dim=1)
    extras = 1
    if y is not None:
        h = torch.cat((F.embedding(y, sd["label_emb.weight"]).unsqueeze(1), h), dim=1)
        extras = 2
    h = h + sd["pos_embed"]
    n_half = depth // 2
    skips = []
    for i in range(n_half):
        h = _block(sd, f"in_blocks.{i}.", h, None, num_heads)
        skips.append(h)
    h = _block(sd, "mid_block.", h, None, num_heads)
    for i in range(n_half):
        h = _block(sd, f"out_blocks.{i}.", h, skips.pop(), num_heads)
    h = F.layer_norm(h, (D,), sd["norm.weight"], sd["norm.bias"])
    h = F.linear(h, sd["decoder_pred.weight"], sd["decoder_pred.bias"])[:, extras:, :]
    g = int(Lp ** 0.5)
    # 'B (h w) (p1 p2 C) -> B C (h p1) (w p2)'
    img = h.reshape(-1, g, g, p, p, C).permute(0, 5, 1, 3, 2, 4).reshape(-1, C, g * p, g * p)
    if conv:
        img = F.conv2d(img, sd["final_layer.weight"], sd["final_layer.bias"], padding=1)
    return img
