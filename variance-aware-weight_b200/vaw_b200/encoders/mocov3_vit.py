"""MoCo-v3 ViT teacher of the REPA alignment loss, forward only, on the B200 library.

Mirrors the reference's encoders/mocov3_vit.py:52-106,138-165 (`VisionTransformerMoCo`, `vit_small/base/large`: fixed 2-D
sin-cos position embedding, cls token, img_size 256 / patch 16) over timm 0.9.2's `VisionTransformer.forward_features`
(patch-embed conv -> cls + pos -> pre-LN blocks with erf-GELU MLPs -> final LayerNorm; that class is a third-party
dependency absent from /root/reference, so parity for it is anchored on the restatement in oracle/vit.py), and
tools/align_utils.py:19-50 (`preprocess_raw_image`, `get_feature`: features[:, 1:] drops the cls token).

Same parameter names as timm (`cls_token`, `pos_embed`, `patch_embed.proj.*`, `blocks.N.{norm1,attn.qkv,attn.proj,norm2,
mlp.fc1,mlp.fc2}.*`, `norm.*`), so a MoCo-v3 checkpoint loads with `load_state_dict`.  The forward is a sequence of
library launches: vaw_patchify_norm -> tcgen05 GEMM (+bias +pos through the residual-table epilogue) -> vaw_vit_assemble
-> per block { vaw_ln_fwd, GEMM qkv, vaw_attn_fwd, GEMM proj + residual, vaw_ln_fwd, GEMM fc1 + GELU(erf), GEMM fc2 +
residual } -> vaw_ln_fwd.  No autograd (the teacher is frozen, align_utils.py:45 `torch.no_grad()`), no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from .. import _lib as L

_P = C.c_void_p
L.register("vaw_ln_fwd", [_P] * 3 + [C.c_longlong, C.c_int] + [_P] * 5 + [C.c_int, C.c_int, C.c_float, _P])
L.register("vaw_attn_fwd", [_P] * 3 + [C.c_int] * 4 + [_P])
L.register("vaw_cast_f32_bf16", [_P, _P, C.c_longlong, _P])
L.register("vaw_patchify_in", [_P] * 2 + [C.c_int] * 5 + [_P])
L.register("vaw_patchify_norm", [_P] * 4 + [C.c_int] * 5 + [_P])
L.register("vaw_vit_assemble", [_P] * 4 + [C.c_int] * 3 + [_P])

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)   # timm.data constants used by tools/align_utils.py:3,28
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


class _Attn(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attn(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, hidden)


class _PatchEmbed(nn.Module):
    def __init__(self, in_chans, dim, patch):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, dim, kernel_size=patch, stride=patch)


def sincos_2d(h, w, dim, temperature=10000.0):
    """Reference encoders/mocov3_vit.py:81-97 (note its meshgrid is 'ij' over (w, h))."""
    grid_w, grid_h = torch.meshgrid(torch.arange(w, dtype=torch.float32), torch.arange(h, dtype=torch.float32),
                                    indexing="ij")
    assert dim % 4 == 0, "Embed dimension must be divisible by 4 for 2D sin-cos position embedding"
    pos_dim = dim // 4
    omega = 1.0 / (temperature ** (torch.arange(pos_dim, dtype=torch.float32) / pos_dim))
    out_w = grid_w.flatten()[:, None] * omega[None, :]
    out_h = grid_h.flatten()[:, None] * omega[None, :]
    pos = torch.cat([torch.sin(out_w), torch.cos(out_w), torch.sin(out_h), torch.cos(out_h)], dim=1)[None]
    return torch.cat([torch.zeros(1, 1, dim), pos], dim=1)


class VisionTransformerMoCo(nn.Module):
    def __init__(self, img_size=256, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4,
                 **unused):
        super().__init__()
        assert img_size % patch_size == 0 and embed_dim % num_heads == 0
        self.img_size, self.patch_size, self.in_chans = img_size, patch_size, in_chans
        self.embed_dim, self.depth, self.num_heads = embed_dim, depth, num_heads
        self.grid = img_size // patch_size
        self.num_patches = self.grid * self.grid
        self.hidden = int(embed_dim * mlp_ratio)
        self.patch_embed = _PatchEmbed(in_chans, embed_dim, patch_size)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(sincos_2d(self.grid, self.grid, embed_dim), requires_grad=False)
        self.blocks = nn.ModuleList([_Block(embed_dim, self.hidden) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self._init_weights()
        self.requires_grad_(False)
        self._shadow = None      # bf16 copies of the GEMM weights, rebuilt when the parameters change
        self._shadow_key = None

    def _init_weights(self):
        """Reference :58-78."""
        for name, m in self.named_modules():
            if isinstance(m, nn.Linear):
                if "qkv" in name:
                    val = math.sqrt(6.0 / float(m.weight.shape[0] // 3 + m.weight.shape[1]))
                    nn.init.uniform_(m.weight, -val, val)
                else:
                    nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)
        nn.init.normal_(self.cls_token, std=1e-6)
        val = math.sqrt(6.0 / float(3 * self.patch_size * self.patch_size + self.embed_dim))
        nn.init.uniform_(self.patch_embed.proj.weight, -val, val)
        nn.init.zeros_(self.patch_embed.proj.bias)

    # ---- bf16 weight shadows -------------------------------------------------------------------------------
    def _gemm_weights(self):
        ws = [self.patch_embed.proj.weight]
        for b in self.blocks:
            ws += [b.attn.qkv.weight, b.attn.proj.weight, b.mlp.fc1.weight, b.mlp.fc2.weight]
        return ws

    def _shadows(self):
        ws = self._gemm_weights()
        key = tuple((w.data_ptr(), w._version) for w in ws)
        if self._shadow is None or key != self._shadow_key:
            sh = []
            for w in ws:
                w32 = w.detach().float().contiguous()
                s = torch.empty(w32.shape, dtype=torch.bfloat16, device=w32.device)
                L.call("vaw_cast_f32_bf16", w32.data_ptr(), s.data_ptr(), w32.numel(), L.stream_ptr())
                sh.append(s)
            self._shadow, self._shadow_key = sh, key
        return self._shadow

    # ---- launches ------------------------------------------------------------------------------------------
    @staticmethod
    def _gemm(A, W, M, N, K, epi, out=None, out2=None, bias=None, resid=None, resid_mod=0):
        g = L.GemmArgs()
        g.A, g.B, g.lda, g.ldb = A.data_ptr(), W.data_ptr(), K, K
        g.M, g.N, g.K, g.epilogue = M, N, K, epi
        g.out, g.out2, g.bias, g.resid = L.ptr(out), L.ptr(out2), L.ptr(bias), L.ptr(resid)
        g.rows_per_sample, g.resid_mod = 1, resid_mod
        L.call("vaw_gemm_bf16", C.byref(g), L.stream_ptr())

    @staticmethod
    def _ln(x, norm, y, stats, M, D):
        L.call("vaw_ln_fwd", x.data_ptr(), None, None, 0, 1, norm.weight.data_ptr(), norm.bias.data_ptr(), y.data_ptr(),
               stats[0].data_ptr(), stats[1].data_ptr(), M, D, float(norm.eps), L.stream_ptr())

    @torch.no_grad()
    def forward_features(self, x, raw_pixels=False):
        """x: [B, C, H, W] fp32, already normalised (the reference's encoder.forward_features), or raw 0..255 pixels with
        raw_pixels=True (preprocess_raw_image fused into the patchify).  Returns bf16 [B, 1 + L, D]."""
        L.require_cuda(x)
        B = x.shape[0]
        assert x.shape[1:] == (self.in_chans, self.img_size, self.img_size), f"unexpected input {tuple(x.shape)}"
        dev, D, Lp, Hd = x.device, self.embed_dim, self.num_patches, self.hidden
        T = Lp + 1
        M, Kp = B * T, self.in_chans * self.patch_size ** 2
        sh = self._shadows()
        x = x.float().contiguous()
        bf = dict(dtype=torch.bfloat16, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        patches = torch.empty(B * Lp, Kp, **bf)
        if raw_pixels:
            assert self.in_chans == 3
            mean, std = torch.tensor(IMAGENET_DEFAULT_MEAN, **f32), torch.tensor(IMAGENET_DEFAULT_STD, **f32)
            L.call("vaw_patchify_norm", x.data_ptr(), mean.data_ptr(), std.data_ptr(), patches.data_ptr(), B,
                   self.in_chans, self.img_size, self.img_size, self.patch_size, L.stream_ptr())
        else:
            L.call("vaw_patchify_in", x.data_ptr(), patches.data_ptr(), B, self.in_chans, self.img_size, self.img_size,
                   self.patch_size, L.stream_ptr())
        pos = self.pos_embed.detach().float().contiguous()          # [1, 1 + L, D]
        tok = torch.empty(B * Lp, D, **f32)
        self._gemm(patches, sh[0], B * Lp, D, Kp, L.EPI_RES, out2=tok, bias=self.patch_embed.proj.bias,
                   resid=pos[0, 1:], resid_mod=Lp)
        xa, xb = torch.empty(M, D, **f32), torch.empty(M, D, **f32)
        L.call("vaw_vit_assemble", tok.data_ptr(), self.cls_token.data_ptr(), pos.data_ptr(), xa.data_ptr(), B, Lp, D,
               L.stream_ptr())
        xn = torch.empty(M, D, **bf)
        qkv = torch.empty(M, 3 * D, **bf)
        ao = torch.empty(M, D, **bf)
        lse = torch.empty(B * self.num_heads * T, **f32)
        h_pre, h_act = torch.empty(M, Hd, **bf), torch.empty(M, Hd, **bf)
        stats = (torch.empty(M, **f32), torch.empty(M, **f32))
        for i, blk in enumerate(self.blocks):
            wq, wp, w1, w2 = sh[1 + 4 * i: 5 + 4 * i]
            self._ln(xa, blk.norm1, xn, stats, M, D)
            self._gemm(xn, wq, M, 3 * D, D, L.EPI_BF16, out=qkv, bias=blk.attn.qkv.bias)
            L.call("vaw_attn_fwd", qkv.data_ptr(), ao.data_ptr(), lse.data_ptr(), B, T, self.num_heads,
                   D // self.num_heads, L.stream_ptr())
            self._gemm(ao, wp, M, D, D, L.EPI_RES, out2=xb, bias=blk.attn.proj.bias, resid=xa)
            self._ln(xb, blk.norm2, xn, stats, M, D)
            self._gemm(xn, w1, M, Hd, D, L.EPI_GELU_ERF, out=h_pre, out2=h_act, bias=blk.mlp.fc1.bias)
            self._gemm(h_act, w2, M, D, Hd, L.EPI_RES, out2=xa, bias=blk.mlp.fc2.bias, resid=xb)
        out = torch.empty(B, T, D, **bf)
        self._ln(xa, self.norm, out, stats, M, D)
        return out

    def forward(self, x):
        return self.forward_features(x)


def vit_small(**kw):
    return VisionTransformerMoCo(img_size=256, patch_size=16, embed_dim=384, depth=12, num_heads=12, mlp_ratio=4, **kw)


def vit_base(**kw):
    return VisionTransformerMoCo(img_size=256, patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, **kw)


def vit_large(**kw):
    return VisionTransformerMoCo(img_size=256, patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4, **kw)


def get_feature(args, images, encoder):
    """Reference tools/align_utils.py:43-50 for the mocov3 encoders: raw 0..255 pixels -> [N, L, D] features (cls
    token dropped)."""
    if "mocov3" not in args.enc_type:
        raise NotImplementedError(f"{args.enc_type}: only the MoCo-v3 ViT teacher runs on the B200 library")
    return encoder.forward_features(images, raw_pixels=True)[:, 1:]
