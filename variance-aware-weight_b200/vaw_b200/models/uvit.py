"""U-ViT with the reference's constructor and parameter names, executed by libvaw_b200.so.

Mirrors /root/reference/models/uvit.py:139-250 (class UViT) and :258-276 (UViT_S / S_D / M / L / H).
`forward(x, timesteps, y=None)` returns a tensor like the reference (:220-250).  Same architecture as
vaw_b200.models.dit: a flat fp32 parameter buffer + bf16 shadow + flat gradient buffer, forward and backward as one
C call each (csrc/uvit_engine.cu).  CUDA only, no PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from .. import _lib as L
from ._flat import FlatEngineModule, Named, ParamHolder


class UViTCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("B", "T", "D", "H", "depth", "hidden", "C", "P", "img_h", "img_w", "extras",
                                        "table_rows", "conv")]


L.register("vaw_uvit_param_layout", [C.POINTER(UViTCfg), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p])
L.register("vaw_uvit_workspace_bytes", [C.POINTER(UViTCfg), C.c_void_p])
L.register("vaw_uvit_forward", [C.POINTER(UViTCfg)] + [C.c_void_p] * 8)
L.register("vaw_uvit_backward", [C.POINTER(UViTCfg)] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p])
L.register("vaw_uvit_backward_ev", [C.POINTER(UViTCfg)] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p, C.c_void_p])
L.register("vaw_cast_f32_bf16", [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p])


def _trunc_normal_(t, std=0.02):
    return nn.init.trunc_normal_(t, std=std, a=-2.0, b=2.0)  # tools/timm.py:44-93 defaults


class _Block(nn.Module):
    def __init__(self, dim, hidden, skip):
        super().__init__()
        self.norm1 = ParamHolder((dim,), (dim,))
        self.attn = Named()
        self.attn.qkv = ParamHolder((3 * dim, dim))            # qkv_bias=False (uvit.py:62,141)
        self.attn.proj = ParamHolder((dim, dim), (dim,))
        self.norm2 = ParamHolder((dim,), (dim,))
        self.mlp = Named()
        self.mlp.fc1 = ParamHolder((hidden, dim), (hidden,))
        self.mlp.fc2 = ParamHolder((dim, hidden), (dim,))
        self.skip_linear = ParamHolder((dim, 2 * dim), (dim,)) if skip else None


class UViT(FlatEngineModule):
    def __init__(self, image_size=224, patch_size=16, in_channels=3, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.,
                 qkv_bias=False, qk_scale=None, norm_layer=nn.LayerNorm, mlp_time_embed=False, num_classes=-1,
                 use_checkpoint=False, conv=True, skip=True, class_dropout_prob=0.0):
        super().__init__()
        if mlp_time_embed or qkv_bias or qk_scale is not None or not skip or norm_layer is not nn.LayerNorm:
            raise NotImplementedError("vaw_b200 UViT covers the reference's shipped configurations "
                                      "(mlp_time_embed=False, qkv_bias=False, skip=True, nn.LayerNorm)")
        # the reference builds depth//2 in-blocks + mid + depth//2 out-blocks (uvit.py:167-185)
        self.num_features = self.embed_dim = embed_dim
        self.num_classes = num_classes
        self.in_channels = in_channels
        self.patch_size = patch_size
        self.image_size = image_size
        self.num_heads = num_heads
        self.class_dropout_prob = class_dropout_prob
        self.hidden = int(embed_dim * mlp_ratio)
        self.conv = bool(conv)
        D = embed_dim
        n_half = depth // 2
        self.num_blocks = 2 * n_half + 1
        num_patches = (image_size // patch_size) ** 2
        self.patch_embed = Named()
        self.patch_embed.proj = ParamHolder((D, in_channels, patch_size, patch_size), (D,))
        self.patch_embed.patch_size = patch_size
        if num_classes > 0:
            self.label_emb = ParamHolder((num_classes + int(class_dropout_prob > 0), D))
            self.extras = 2
        else:
            self.label_emb = None
            self.extras = 1
        self.pos_embed = nn.Parameter(torch.zeros(1, self.extras + num_patches, D))
        self.in_blocks = nn.ModuleList([_Block(D, self.hidden, False) for _ in range(n_half)])
        self.mid_block = _Block(D, self.hidden, False)
        self.out_blocks = nn.ModuleList([_Block(D, self.hidden, True) for _ in range(n_half)])
        self.norm = ParamHolder((D,), (D,))
        self.patch_dim = patch_size ** 2 * in_channels
        self.decoder_pred = ParamHolder((self.patch_dim, D), (self.patch_dim,))
        self.final_layer = ParamHolder((in_channels, in_channels, 3, 3), (in_channels,)) if conv else None
        self._cfg_static = dict(T=self.extras + num_patches, D=D, H=num_heads, depth=self.num_blocks, hidden=self.hidden,
                                C=in_channels, P=patch_size, img_h=image_size, img_w=image_size, extras=self.extras,
                                table_rows=(num_classes + int(class_dropout_prob > 0)) if num_classes > 0 else 0,
                                conv=int(self.conv))
        self._init_weights()

    # ------------------------------------------------------------------------------------------------
    def _init_weights(self):
        """Reference scheme (uvit.py:194-204): trunc_normal(0.02) for Linear weights and pos_embed, zero biases,
        LayerNorm weight 1 / bias 0; Conv2d and Embedding keep PyTorch's default initialisation."""
        with torch.no_grad():
            _trunc_normal_(self.pos_embed)
            for blk in self._blocks():
                for lin in (blk.attn.qkv, blk.attn.proj, blk.mlp.fc1, blk.mlp.fc2, blk.skip_linear):
                    if lin is not None:
                        _trunc_normal_(lin.weight)
                        if hasattr(lin, "bias"):
                            nn.init.zeros_(lin.bias)
                for ln in (blk.norm1, blk.norm2):
                    nn.init.ones_(ln.weight)
                    nn.init.zeros_(ln.bias)
            nn.init.ones_(self.norm.weight)
            nn.init.zeros_(self.norm.bias)
            _trunc_normal_(self.decoder_pred.weight)
            nn.init.zeros_(self.decoder_pred.bias)
            ref_conv = nn.Conv2d(self.in_channels, self.embed_dim, self.patch_size, self.patch_size)
            self.patch_embed.proj.weight.copy_(ref_conv.weight)
            self.patch_embed.proj.bias.copy_(ref_conv.bias)
            if self.final_layer is not None:
                ref_fin = nn.Conv2d(self.in_channels, self.in_channels, 3, padding=1)
                self.final_layer.weight.copy_(ref_fin.weight)
                self.final_layer.bias.copy_(ref_fin.bias)
            if self.label_emb is not None:
                nn.init.normal_(self.label_emb.weight)

    def _blocks(self):
        return list(self.in_blocks) + [self.mid_block] + list(self.out_blocks)

    def no_weight_decay(self):
        return {"pos_embed"}

    def _cfg(self, batch):
        return UViTCfg(B=batch, **self._cfg_static)

    def _layout(self):
        cfg = self._cfg(1)
        cap = 10 + 13 * self.num_blocks
        off = (C.c_longlong * cap)()
        num = (C.c_longlong * cap)()
        n = C.c_int()
        total = C.c_longlong()
        L.call("vaw_uvit_param_layout", C.byref(cfg), off, num, cap, C.byref(n), C.byref(total))
        return list(off)[: n.value], list(num)[: n.value], total.value

    def _slots(self):
        off, num, total = self._layout()
        head = [self.patch_embed.proj.weight, self.patch_embed.proj.bias,
                self.label_emb.weight if self.label_emb is not None else None, self.pos_embed, self.norm.weight,
                self.norm.bias, self.decoder_pred.weight, self.decoder_pred.bias,
                self.final_layer.weight if self.final_layer is not None else None,
                self.final_layer.bias if self.final_layer is not None else None]
        slots = []
        for i, p in enumerate(head):
            if p is not None and p.numel() > 0:
                assert p.numel() == num[i], (i, tuple(p.shape), num[i])
                slots.append((p, off[i]))
        for bi, b in enumerate(self._blocks()):
            base = 10 + 13 * bi
            ps = [b.norm1.weight, b.norm1.bias, b.attn.qkv.weight, b.attn.proj.weight, b.attn.proj.bias, b.norm2.weight,
                  b.norm2.bias, b.mlp.fc1.weight, b.mlp.fc1.bias, b.mlp.fc2.weight, b.mlp.fc2.bias,
                  b.skip_linear.weight if b.skip_linear is not None else None,
                  b.skip_linear.bias if b.skip_linear is not None else None]
            for j, p in enumerate(ps):
                if p is not None:
                    assert p.numel() == num[base + j], (bi, j, tuple(p.shape), num[base + j])
                    slots.append((p, off[base + j]))
        return slots, total

    def _workspace_bytes(self, batch):
        cfg = self._cfg(batch)
        nbytes = C.c_longlong()
        L.call("vaw_uvit_workspace_bytes", C.byref(cfg), C.byref(nbytes))
        return nbytes.value

    @property
    def depth(self):
        return self.num_blocks

    def block_grad_ranges(self):
        """Element ranges of the flat gradient buffer that become final with each block (in-blocks, mid, out-blocks in
        engine order: its thirteen tensors are contiguous), the range that only becomes final at the end of backward
        (patch embedding, label table, pos_embed, final norm, decoder_pred, 3x3 conv), and the buffer length.  The long
        skip connections only carry activation gradients, so blocks complete strictly in reverse order."""
        off, num, total = self._layout()
        per_block = []
        for bi in range(self.num_blocks):
            base = 10 + 13 * bi
            end = off[base + 13] if base + 13 < len(off) else total
            per_block.append([(off[base], end)])
        tail = [(0, off[10])]
        return per_block, tail, total

    def token_drop(self, labels, train=True):
        """Label dropout for classifier-free guidance (uvit.py:206-218)."""
        if train and self.class_dropout_prob > 0:
            drop = torch.rand(labels.shape[0], device=labels.device) < self.class_dropout_prob
            labels = torch.where(drop, self.num_classes, labels)
        return labels

    def forward(self, x, timesteps, y=None, **kwargs):
        if not x.is_cuda:
            raise L.VawError("vaw_b200.models.UViT runs on CUDA only (no CPU fallback)")
        if x.shape[1:] != (self.in_channels, self.image_size, self.image_size):
            raise AssertionError(f"expected input [N,{self.in_channels},{self.image_size},{self.image_size}], got {tuple(x.shape)}")
        if self.extras == 2:
            if y is None:
                raise ValueError("class-conditional UViT needs labels y")
            y = self.token_drop(y, self.training).to(torch.int64).contiguous()
        else:
            y = None
        self._ensure_flat(x.device)
        self._ensure_workspace(x.shape[0], x.device)
        return _UViTFunction.apply(self, x.float().contiguous(), timesteps.float().contiguous(), y, self._flat)


class _UViTFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, t, y, flat):
        B = x.shape[0]
        cfg = model._cfg(B)
        model._refresh_shadow()
        out = torch.empty_like(x)
        L.call("vaw_uvit_forward", C.byref(cfg), flat.data_ptr(), model._shadow.data_ptr(), model._ws.data_ptr(),
               x.data_ptr(), t.data_ptr(), L.ptr(y), out.data_ptr(), L.stream_ptr())
        model._fwd_serial += 1
        ctx.model, ctx.serial, ctx.batch, ctx.y = model, model._fwd_serial, B, y
        return out

    @staticmethod
    def backward(ctx, dout):
        model = ctx.model
        if ctx.serial != model._fwd_serial:
            raise L.VawError("UViT backward after a newer forward: the activation workspace was overwritten "
                             "(run forward/backward pairs in order)")
        cfg = model._cfg(ctx.batch)
        dout = dout.float().contiguous()
        fresh = model._bind_grads()
        ev = None
        if model._events is not None:
            ev = (C.c_void_p * len(model._events))(*[e.cuda_event for e in model._events])
        L.call("vaw_uvit_backward_ev", C.byref(cfg), model._flat.data_ptr(), model._shadow.data_ptr(),
               model._gflat.data_ptr(), model._ws.data_ptr(), dout.data_ptr(), L.ptr(ctx.y), 0 if fresh else 1, ev,
               L.stream_ptr())
        if model._post_backward is not None:
            model._post_backward()
        return None, None, None, None, None


def UViT_S(image_size, patch_size, in_channels, num_classes, class_dropout_prob, **kwargs):
    return UViT(image_size=image_size, patch_size=patch_size, in_channels=in_channels, embed_dim=512, depth=13,
                num_heads=8, mlp_ratio=4, num_classes=num_classes, class_dropout_prob=class_dropout_prob, **kwargs)


def UViT_S_D(image_size, patch_size, in_channels, num_classes, class_dropout_prob, **kwargs):
    return UViT(image_size=image_size, patch_size=patch_size, in_channels=in_channels, embed_dim=512, depth=17,
                num_heads=8, mlp_ratio=4, num_classes=num_classes, class_dropout_prob=class_dropout_prob, **kwargs)


def UViT_M(image_size, patch_size, in_channels, num_classes, class_dropout_prob, **kwargs):
    return UViT(image_size=image_size, patch_size=patch_size, in_channels=in_channels, embed_dim=768, depth=17,
                num_heads=12, mlp_ratio=4, num_classes=num_classes, class_dropout_prob=class_dropout_prob, **kwargs)


def UViT_L(image_size, patch_size, in_channels, num_classes, class_dropout_prob, **kwargs):
    return UViT(image_size=image_size, patch_size=patch_size, in_channels=in_channels, embed_dim=1024, depth=21,
                num_heads=16, mlp_ratio=4, num_classes=num_classes, class_dropout_prob=class_dropout_prob, **kwargs)


def UViT_H(image_size, patch_size, in_channels, num_classes, class_dropout_prob, **kwargs):
    return UViT(image_size=image_size, patch_size=patch_size, in_channels=in_channels, embed_dim=1152, depth=29,
                num_heads=16, mlp_ratio=4, num_classes=num_classes, class_dropout_prob=class_dropout_prob, **kwargs)


UViT_models = {"UViT-S": UViT_S, "UViT-S-D": UViT_S_D, "UViT-M": UViT_M, "UViT-L": UViT_L, "UViT-H": UViT_H}
