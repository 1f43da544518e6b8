"""ORACLE (test infrastructure, not product code): plain-PyTorch restatement of the reference DiT forward.

Follows /root/reference/models/dit.py — modulate (:24-25), TimestepEmbedder (:41-79), LabelEmbedder (:82-110),
DiTBlock (:118-137), FinalLayer (:140-155), DiT.forward (:258-280), unpatchify (:243-256) — and the three timm 0.9.2
modules it imports (:17): PatchEmbed, Attention, Mlp, restated from timm's published semantics (SURVEY.md §A.3;
timm is a pip dependency pinned in the reference's requirements.txt:110 and absent from /root/reference, so DiT
parity at that boundary is "unpinned" by reference tests).

Written functionally over a state_dict (name -> tensor with the reference's parameter names) so the same weights can
be fed to the reference module (in tests/golden/make_golden.py, which pins this file against the executed reference)
and to the CUDA engine.  Runs on CPU or GPU, in fp32 or under torch.autocast(bf16) (the reference's AMP path with
dtype bf16).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def timestep_embedding(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _modulate(x, shift, scale):
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def _attention(x, w_qkv, b_qkv, w_proj, b_proj, num_heads):
    B, N, C = x.shape
    hd = C // num_heads
    qkv = F.linear(x, w_qkv, b_qkv).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0)
    o = o.transpose(1, 2).reshape(B, N, C)
    return F.linear(o, w_proj, b_proj)


def dit_forward(sd, x, t, y, *, patch_size, num_heads, depth, learn_align=False, encoder_depth=0):
    """sd: state dict with the reference's names.  x [N,C,H,W], t [N] float (already scaled), y [N] long or None.
    Returns (out [N,C_out,H,W], zs [N,T,z] or None)."""
    p = patch_size
    D = sd["pos_embed"].shape[-1]
    # PatchEmbed: Conv2d(k=s=p) -> flatten -> transpose
    h = F.conv2d(x, sd["x_embedder.proj.weight"], sd["x_embedder.proj.bias"], stride=p)
    h = h.flatten(2).transpose(1, 2) + sd["pos_embed"]
    # conditioning
    tf = timestep_embedding(t, 256)
    te = F.linear(F.silu(F.linear(tf, sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])),
                  sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])
    c = te
    if y is not None:
        c = te + F.embedding(y, sd["y_embedder.embedding_table.weight"])
    zs = None
    for i in range(depth):
        pre = f"blocks.{i}."
        mod = F.linear(F.silu(c), sd[pre + "adaLN_modulation.1.weight"], sd[pre + "adaLN_modulation.1.bias"])
        sh1, sc1, g1, sh2, sc2, g2 = mod.chunk(6, dim=1)
        a = _attention(_modulate(F.layer_norm(h, (D,), eps=1e-6), sh1, sc1), sd[pre + "attn.qkv.weight"],
                       sd[pre + "attn.qkv.bias"], sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"], num_heads)
        h = h + g1.unsqueeze(1) * a
        m = _modulate(F.layer_norm(h, (D,), eps=1e-6), sh2, sc2)
        m = F.linear(F.gelu(F.linear(m, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"]), approximate="tanh"),
                     sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
        h = h + g2.unsqueeze(1) * m
        if learn_align and (i + 1) == encoder_depth:
            z = F.silu(F.linear(h, sd["projectors.0.weight"], sd["projectors.0.bias"]))
            z = F.silu(F.linear(z, sd["projectors.2.weight"], sd["projectors.2.bias"]))
            zs = F.linear(z, sd["projectors.4.weight"], sd["projectors.4.bias"])
    mod = F.linear(F.silu(c), sd["final_layer.adaLN_modulation.1.weight"], sd["final_layer.adaLN_modulation.1.bias"])
    sh, sc = mod.chunk(2, dim=1)
    h = F.linear(_modulate(F.layer_norm(h, (D,), eps=1e-6), sh, sc), sd["final_layer.linear.weight"],
                 sd["final_layer.linear.bias"])
    # unpatchify: (N, T, p*p*C) -> (N, C, H, W)
    N, T, _ = h.shape
    g = int(T ** 0.5)
    co = h.shape[-1] // (p * p)
    h = h.reshape(N, g, g, p, p, co)
    out = torch.einsum("nhwpqc->nchpwq", h).reshape(N, co, g * p, g * p)
    return out, zs
