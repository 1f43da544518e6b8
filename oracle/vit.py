"""ORACLE (test infrastructure, not product code): the frozen MoCo-v3 ViT teacher of the REPA loss, restated in torch.

Follows /root/reference/encoders/mocov3_vit.py:52-106 (VisionTransformerMoCo: fixed 2-D sin-cos position embedding with a
zero cls slot), :159-165 (vit_base: img 256, patch 16, D 768, depth 12, 12 heads, qkv_bias, LayerNorm eps 1e-6) and
/root/reference/tools/align_utils.py:19-50 (preprocess_raw_image, get_feature -> features[:, 1:]).  The forward itself
lives in timm==0.9.2 (requirements.txt:110), which is NOT vendored under /root/reference: VisionTransformer.
forward_features is restated here from its published definition (patch-embed conv -> [cls; tokens] + pos_embed ->
pre-LN blocks x + attn(norm1 x), x + mlp(norm2 x) with erf GELU -> final LayerNorm).  PARITY UNPINNED for that part: the
reference holds no test or golden vector at this boundary; tests/golden/vit_golden.npz pins this file against the
reference's own VisionTransformerMoCo executed over the same restatement (oracle/ref_stubs/timm), i.e. the position
embedding, the parameter layout and preprocess_raw_image are pinned, timm's block arithmetic is not.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def preprocess_raw_image(x):
    """align_utils.py:26-28 (mocov3): x / 255 then torchvision Normalize."""
    x = x / 255.
    mean = torch.as_tensor(IMAGENET_DEFAULT_MEAN, dtype=x.dtype, device=x.device).view(-1, 1, 1)
    std = torch.as_tensor(IMAGENET_DEFAULT_STD, dtype=x.dtype, device=x.device).view(-1, 1, 1)
    return (x - mean) / std


def sincos_pos_embed(h, w, dim, temperature=10000.):
    """mocov3_vit.py:81-97."""
    gw, gh = torch.meshgrid(torch.arange(w, dtype=torch.float32), torch.arange(h, dtype=torch.float32), indexing="ij")
    pd = dim // 4
    omega = 1. / (temperature ** (torch.arange(pd, dtype=torch.float32) / pd))
    ow = torch.einsum("m,d->md", gw.flatten(), omega)
    oh = torch.einsum("m,d->md", gh.flatten(), omega)
    pe = torch.cat([torch.sin(ow), torch.cos(ow), torch.sin(oh), torch.cos(oh)], dim=1)[None]
    return torch.cat([torch.zeros(1, 1, dim), pe], dim=1)


def vit_forward_features(sd, x, *, patch_size, num_heads, depth, eps=1e-6):
    """sd: state dict with timm's names; x: [N, C, H, W] normalised pixels.  Returns [N, 1 + L, D]."""
    D = sd["cls_token"].shape[-1]
    h = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch_size)
    h = h.flatten(2).transpose(1, 2)
    h = torch.cat([sd["cls_token"].expand(h.shape[0], -1, -1).to(h.dtype), h], dim=1) + sd["pos_embed"]
    for i in range(depth):
        p = f"blocks.{i}."
        a = F.layer_norm(h, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
        B, T, _ = a.shape
        qkv = F.linear(a, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
        q, k, v = qkv.reshape(B, T, 3, num_heads, D // num_heads).permute(2, 0, 3, 1, 4).unbind(0)
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, T, D)
        h = h + F.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        m = F.layer_norm(h, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
        m = F.gelu(F.linear(m, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
        h = h + F.linear(m, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return F.layer_norm(h, (D,), sd["norm.weight"], sd["norm.bias"], eps)


def get_feature(sd, images, *, patch_size, num_heads, depth):
    """align_utils.py:43-50 for mocov3: raw pixels -> patch-token features (cls dropped)."""
    return vit_forward_features(sd, preprocess_raw_image(images), patch_size=patch_size, num_heads=num_heads,
                                depth=depth)[:, 1:]
