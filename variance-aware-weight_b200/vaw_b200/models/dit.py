"""DiT with the reference's constructor, parameter names and return convention, executed by libvaw_b200.so.

Mirrors /root/reference/models/dit.py:157-280 (class DiT) and :361-382 (DiT_S/B/L/XL).  `forward(x, t, y)` returns
the tuple `(x, zs)` like the reference (:258-280).  The nn.Module is only a *view* of the model: all parameters
live in one flat fp32 buffer (plus a bf16 shadow for the tensor cores and a flat fp32 gradient buffer); forward
and backward are one C call each (csrc/dit_engine.cu).  There is no PyTorch fallback: without a CUDA device and
the built library, forward raises.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from .. import _lib as L


class DiTCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "B", "T", "D", "H", "depth", "hidden", "C_in", "C_out", "P", "img_h", "img_w", "table_rows", "freq_dim",
        "learn_align", "encoder_depth", "proj_dim", "z_dim")]


L.register("vaw_dit_param_layout", [C.POINTER(DiTCfg), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p])
L.register("vaw_dit_workspace_bytes", [C.POINTER(DiTCfg), C.c_void_p])
L.register("vaw_dit_forward", [C.POINTER(DiTCfg)] + [C.c_void_p] * 9)
L.register("vaw_dit_forward_align", [C.POINTER(DiTCfg)] + [C.c_void_p] * 11)
L.register("vaw_dit_forward_ev", [C.POINTER(DiTCfg)] + [C.c_void_p] * 10)
L.register("vaw_dit_infer_workspace_bytes", [C.POINTER(DiTCfg), C.c_void_p])
L.register("vaw_dit_forward_infer", [C.POINTER(DiTCfg)] + [C.c_void_p] * 9)
L.register("vaw_dit_backward", [C.POINTER(DiTCfg)] + [C.c_void_p] * 7 + [C.c_int, C.c_void_p, C.c_void_p])
L.register("vaw_cast_f32_bf16", [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p])


from ._flat import FlatEngineModule, Named as _Named, ParamHolder as _ParamHolder, Slot as _Slot


def _sincos_1d(dim, pos):
    omega = np.arange(dim // 2, dtype=np.float64) / (dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_pos_embed_2d(dim, grid):
    """Fixed 2-D sin/cos table, same construction as reference dit.py:307-354 (w axis first)."""
    gh = np.arange(grid, dtype=np.float32)
    gw = np.arange(grid, dtype=np.float32)
    mesh = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, grid, grid)
    return np.concatenate([_sincos_1d(dim // 2, mesh[0]), _sincos_1d(dim // 2, mesh[1])], axis=1)


class DiT(FlatEngineModule):
    # training_losses may hand the teacher features to forward (`align_target=`): the 'mse' alignment loss is then
    # accumulated in the epilogue of the last projector GEMM and attached to the returned zs (`zs._vaw_align`)
    supports_fused_align = True

    def __init__(self, image_size=32, patch_size=2, in_channels=4, hidden_size=1152, depth=28, num_heads=16,
                 mlp_ratio=4.0, class_dropout_prob=0.1, num_classes=1000, learn_sigma=False, learn_align=False,
                 encoder_depth=8, z_dims=768, projector_dim=2048):
        super().__init__()
        self.learn_sigma = learn_sigma
        self.learn_align = learn_align
        self.in_channels = in_channels
        self.out_channels = in_channels * 2 if learn_sigma else in_channels
        self.patch_size = patch_size
        self.num_heads = num_heads
        self.encoder_depth = encoder_depth
        self.hidden_size = hidden_size
        self.depth = depth
        self.image_size = image_size
        self.num_classes = num_classes
        self.class_dropout_prob = class_dropout_prob
        self.mlp_hidden = int(hidden_size * mlp_ratio)
        self.z_dims, self.projector_dim = z_dims, projector_dim
        assert not learn_align or encoder_depth > 0, "encoder_depth must be > 0 when learn_align=True"  # dit.py:190
        D, Hd = hidden_size, self.mlp_hidden
        grid = image_size // patch_size
        self.num_patches = grid * grid
        ppc = patch_size * patch_size * self.out_channels
        table_rows = num_classes + (1 if class_dropout_prob > 0 else 0)

        # --- module tree with the reference's names (dit.py:192-202) ---
        self.x_embedder = _Named()
        self.x_embedder.proj = _ParamHolder((D, in_channels, patch_size, patch_size), (D,))
        self.x_embedder.patch_size = (patch_size, patch_size)
        self.x_embedder.num_patches = self.num_patches
        self.t_embedder = _Named()
        self.t_embedder.mlp = nn.ModuleList([_ParamHolder((D, 256), (D,)), _Slot(), _ParamHolder((D, D), (D,))])
        self.y_embedder = _Named()
        self.y_embedder.embedding_table = _ParamHolder((table_rows, D))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.num_patches, D), requires_grad=False)
        blocks = []
        for _ in range(depth):
            b = _Named()
            b.attn = _Named()
            b.attn.qkv = _ParamHolder((3 * D, D), (3 * D,))
            b.attn.proj = _ParamHolder((D, D), (D,))
            b.mlp = _Named()
            b.mlp.fc1 = _ParamHolder((Hd, D), (Hd,))
            b.mlp.fc2 = _ParamHolder((D, Hd), (D,))
            b.adaLN_modulation = nn.ModuleList([_Slot(), _ParamHolder((6 * D, D), (6 * D,))])
            blocks.append(b)
        self.blocks = nn.ModuleList(blocks)
        if learn_align:
            self.projectors = nn.ModuleList([
                _ParamHolder((projector_dim, D), (projector_dim,)), _Slot(),
                _ParamHolder((projector_dim, projector_dim), (projector_dim,)), _Slot(),
                _ParamHolder((z_dims, projector_dim), (z_dims,))])
        else:
            self.projectors = None
        self.final_layer = _Named()
        self.final_layer.linear = _ParamHolder((ppc, D), (ppc,))
        self.final_layer.adaLN_modulation = nn.ModuleList([_Slot(), _ParamHolder((2 * D, D), (2 * D,))])

        self._cfg_static = dict(T=self.num_patches, D=D, H=num_heads, depth=depth, hidden=Hd, C_in=in_channels,
                                C_out=self.out_channels, P=patch_size, img_h=image_size, img_w=image_size,
                                table_rows=table_rows, freq_dim=256, learn_align=int(learn_align),
                                encoder_depth=encoder_depth if learn_align else 0,
                                proj_dim=projector_dim if learn_align else 0, z_dim=z_dims if learn_align else 0)
        self.initialize_weights()

    # ------------------------------------------------------------------------------------------------
    def _ordered_params(self):
        """Parameters in the engine's layout order (csrc/dit_engine.cu ParamId).  The stacked adaLN slot is
        expanded into the per-block tensors, which are contiguous slices of it."""
        t = self.t_embedder.mlp
        head = [self.x_embedder.proj.weight, self.x_embedder.proj.bias, t[0].weight, t[0].bias, t[2].weight,
                t[2].bias, self.y_embedder.embedding_table.weight, self.pos_embed,
                self.final_layer.adaLN_modulation[1].weight, self.final_layer.adaLN_modulation[1].bias,
                self.final_layer.linear.weight, self.final_layer.linear.bias]
        if self.learn_align:
            pr = self.projectors
            head += [pr[0].weight, pr[0].bias, pr[2].weight, pr[2].bias, pr[4].weight, pr[4].bias]
        else:
            head += [None] * 6
        return head

    def _cfg(self, batch):
        return DiTCfg(B=batch, **self._cfg_static)

    def _layout(self):
        cfg = self._cfg(1)
        cap = 20 + 8 * self.depth
        off = (C.c_longlong * cap)()
        num = (C.c_longlong * cap)()
        n = C.c_int()
        total = C.c_longlong()
        L.call("vaw_dit_param_layout", C.byref(cfg), off, num, cap, C.byref(n), C.byref(total))
        return list(off)[: n.value], list(num)[: n.value], total.value

    def _slots(self):
        """[(parameter, flat offset)] for every parameter tensor."""
        off, num, total = self._layout()
        D = self.hidden_size
        slots = []
        for i, p in enumerate(self._ordered_params()):
            if p is not None and p.numel() > 0:
                assert p.numel() == num[i], (i, p.shape, num[i])
                slots.append((p, off[i]))
        for i, b in enumerate(self.blocks):  # stacked adaLN: rows [i*6D, (i+1)*6D)
            slots.append((b.adaLN_modulation[1].weight, off[18] + i * 6 * D * D))
            slots.append((b.adaLN_modulation[1].bias, off[19] + i * 6 * D))
            base = 20 + 8 * i
            # engine order (csrc/dit_engine.cu BlockParam): the four weight matrices, then the four biases
            for j, p in enumerate((b.attn.qkv.weight, b.attn.proj.weight, b.mlp.fc1.weight, b.mlp.fc2.weight,
                                   b.attn.qkv.bias, b.attn.proj.bias, b.mlp.fc1.bias, b.mlp.fc2.bias)):
                slots.append((p, off[base + j]))
        return slots, total

    def _workspace_bytes(self, batch):
        cfg = self._cfg(batch)
        nbytes = C.c_longlong()
        L.call("vaw_dit_workspace_bytes", C.byref(cfg), C.byref(nbytes))
        return nbytes.value

    def block_grad_ranges(self):
        """Element ranges of the flat gradient buffer that become final with each transformer block (its eight
        attention / MLP tensors plus its slice of the stacked adaLN weights and biases), the ranges that only become
        final at the end of backward (embedders, final layer, projectors), and the buffer length."""
        off, num, total = self._layout()
        D = self.hidden_size
        per_block = []
        for i in range(self.depth):
            base = 20 + 8 * i
            per_block.append([(off[base], off[base + 7] + num[base + 7]),
                              (off[18] + i * 6 * D * D, off[18] + (i + 1) * 6 * D * D),
                              (off[19] + i * 6 * D, off[19] + (i + 1) * 6 * D)])
        tail = [(0, off[18])]
        return per_block, tail, total

    def stacked_range(self):
        """Element range of the stacked adaLN weights of all blocks (read by one GEMM at the top of the forward)."""
        off, num, _ = self._layout()
        return off[18], off[18] + num[18]

    def block_shard_ranges(self):
        """For the sharded-optimizer data-parallel mode: per block, the contiguous ranges of LARGE tensors (the four
        weight matrices; the block's slice of the stacked adaLN weight) whose gradients are reduce-scattered, and the
        SMALL ones (the four biases; the adaLN bias slice) that stay replicated (all-reduced)."""
        off, num, total = self._layout()
        D = self.hidden_size
        big, small = [], []
        for i in range(self.depth):
            base = 20 + 8 * i
            big.append([(off[base], off[base + 4]), (off[18] + i * 6 * D * D, off[18] + (i + 1) * 6 * D * D)])
            small.append([(off[base + 4], off[base + 7] + num[base + 7]), (off[19] + i * 6 * D, off[19] + (i + 1) * 6 * D)])
        tail = [(0, off[18])]
        return big, small, tail, total

    # ------------------------------------------------------------------------------------------------
    def initialize_weights(self):
        """Same initialisation scheme as the reference (dit.py:206-241)."""
        D = self.hidden_size
        with torch.no_grad():
            for m in self.modules():
                if isinstance(m, _ParamHolder) and m.weight.dim() == 2 and m is not self.y_embedder.embedding_table:
                    nn.init.xavier_uniform_(m.weight)
                    if hasattr(m, "bias"):
                        nn.init.zeros_(m.bias)
            grid = int(self.num_patches ** 0.5)
            self.pos_embed.copy_(torch.from_numpy(sincos_pos_embed_2d(D, grid)).float().unsqueeze(0))
            w = self.x_embedder.proj.weight
            nn.init.xavier_uniform_(w.view(w.shape[0], -1))
            nn.init.zeros_(self.x_embedder.proj.bias)
            nn.init.normal_(self.y_embedder.embedding_table.weight, std=0.02)
            nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
            nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
            for b in self.blocks:
                nn.init.zeros_(b.adaLN_modulation[1].weight)
                nn.init.zeros_(b.adaLN_modulation[1].bias)
            nn.init.zeros_(self.final_layer.adaLN_modulation[1].weight)
            nn.init.zeros_(self.final_layer.adaLN_modulation[1].bias)
            nn.init.zeros_(self.final_layer.linear.weight)
            nn.init.zeros_(self.final_layer.linear.bias)

    def unpatchify(self, x):
        """(N, T, p*p*C) -> (N, C, H, W); host-side helper kept for API parity (dit.py:243-256)."""
        c, p = self.out_channels, self.patch_size
        h = w = int(x.shape[1] ** 0.5)
        x = x.reshape(x.shape[0], h, w, p, p, c)
        return torch.einsum("nhwpqc->nchpwq", x).reshape(x.shape[0], c, h * p, w * p)

    def token_drop(self, labels, force_drop_ids=None):
        """Label dropout for classifier-free guidance (dit.py:94-103); the draw stays a torch.rand on the device so
        the CUDA generator is consumed exactly like the reference."""
        if force_drop_ids is None:
            drop = torch.rand(labels.shape[0], device=labels.device) < self.class_dropout_prob
        else:
            drop = force_drop_ids == 1
        return torch.where(drop, self.num_classes, labels)

    def forward(self, x, t, y=None, force_drop_ids=None, align_target=None, **kwargs):
        if not x.is_cuda:
            raise L.VawError("vaw_b200.models.DiT runs on CUDA only (no CPU fallback)")
        if x.shape[1:] != (self.in_channels, self.image_size, self.image_size):
            raise AssertionError(f"expected input [N,{self.in_channels},{self.image_size},{self.image_size}], got {tuple(x.shape)}")
        if self._cfg_static["table_rows"] > 0:
            if y is None:
                raise ValueError("class-conditional DiT needs labels y")
            if (self.training and self.class_dropout_prob > 0) or force_drop_ids is not None:
                y = self.token_drop(y, force_drop_ids)
            y = y.to(torch.int64).contiguous()
        else:
            y = None
        self._ensure_flat(x.device)
        pending = self._fwd_wait     # sharded optimizer: events of the in-flight all-gather of the bf16 weights
        if pending is not None and (not torch.is_grad_enabled() or align_target is not None):
            torch.cuda.current_stream().wait_event(pending[-1])   # paths without per-block gating wait for all of it
            self._fwd_wait = pending = None
        if not torch.is_grad_enabled():
            return self._forward_infer(x.float().contiguous(), t.float().contiguous(), y)
        self._ensure_workspace(x.shape[0], x.device)
        feat = None
        if align_target is not None and self.learn_align and align_target.dtype == torch.bfloat16:
            if align_target.shape != (x.shape[0], self.num_patches, self.z_dims):
                raise ValueError(f"align_target {tuple(align_target.shape)} does not match the projector output "
                                 f"{(x.shape[0], self.num_patches, self.z_dims)}")
            feat = align_target.detach().contiguous()
        out, zs, align = _DiTFunction.apply(self, x.float().contiguous(), t.float().contiguous(), y, self._flat, feat)
        if feat is not None:
            zs._vaw_align = (align, align_target, feat)
        return out, zs


def _dit_forward_infer(self, x, t, y):
    """torch.no_grad() forward (the samplers): vaw_dit_forward_infer on a compact workspace of its own - no activation
    stash, and a training forward whose backward is still pending keeps its saved activations."""
    B = x.shape[0]
    cfg = self._cfg(B)
    if self._ws_inf is None or self._ws_inf_batch != B or self._ws_inf.device != x.device:
        nbytes = C.c_longlong()
        L.call("vaw_dit_infer_workspace_bytes", C.byref(cfg), C.byref(nbytes))
        self._ws_inf = None
        self._ws_inf = torch.empty(nbytes.value, dtype=torch.uint8, device=x.device)
        self._ws_inf_batch = B
    self._refresh_shadow()
    out = torch.empty(B, self.out_channels, self.image_size, self.image_size, dtype=torch.bfloat16, device=x.device)
    zs = None
    if self.learn_align:
        zs = torch.empty(B, self.num_patches, self.z_dims, dtype=torch.bfloat16, device=x.device)
    L.call("vaw_dit_forward_infer", C.byref(cfg), self._flat.data_ptr(), self._shadow.data_ptr(), self._ws_inf.data_ptr(),
           x.data_ptr(), t.data_ptr(), L.ptr(y), out.data_ptr(), L.ptr(zs), L.stream_ptr())
    return out, zs


DiT._forward_infer = _dit_forward_infer
DiT._fwd_wait = None
DiT._ws_inf = None
DiT._ws_inf_batch = -1


class _DiTFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, t, y, flat, feat=None):
        B = x.shape[0]
        cfg = model._cfg(B)
        model._refresh_shadow()
        out = torch.empty(B, model.out_channels, model.image_size, model.image_size, dtype=torch.bfloat16, device=x.device)
        zs = None
        if model.learn_align:
            zs = torch.empty(B, model.num_patches, model.z_dims, dtype=torch.bfloat16, device=x.device)
        align = None
        if feat is not None:
            align = torch.empty((), dtype=torch.float32, device=x.device)
            L.call("vaw_dit_forward_align", C.byref(cfg), flat.data_ptr(), model._shadow.data_ptr(),
                   model._ws.data_ptr(), x.data_ptr(), t.data_ptr(), L.ptr(y), out.data_ptr(), L.ptr(zs),
                   feat.data_ptr(), align.data_ptr(), L.stream_ptr())
            ctx.mark_non_differentiable(align)
        elif model._fwd_wait is not None:
            evs = model._fwd_wait
            model._fwd_wait = None
            arr = (C.c_void_p * len(evs))(*[e.cuda_event for e in evs])
            L.call("vaw_dit_forward_ev", C.byref(cfg), flat.data_ptr(), model._shadow.data_ptr(), model._ws.data_ptr(),
                   x.data_ptr(), t.data_ptr(), L.ptr(y), out.data_ptr(), L.ptr(zs), arr, L.stream_ptr())
        else:
            L.call("vaw_dit_forward", C.byref(cfg), flat.data_ptr(), model._shadow.data_ptr(), model._ws.data_ptr(),
                   x.data_ptr(), t.data_ptr(), L.ptr(y), out.data_ptr(), L.ptr(zs), L.stream_ptr())
        model._fwd_serial += 1
        ctx.model, ctx.serial, ctx.batch, ctx.y = model, model._fwd_serial, B, y
        return out, zs, align

    @staticmethod
    def backward(ctx, dout, dzs, _dalign=None):
        model = ctx.model
        if ctx.serial != model._fwd_serial:
            raise L.VawError("DiT backward after a newer forward: the activation workspace was overwritten "
                             "(run forward/backward pairs in order)")
        cfg = model._cfg(ctx.batch)
        if dout is None:
            dout = torch.zeros(ctx.batch, model.out_channels, model.image_size, model.image_size,
                               dtype=torch.bfloat16, device=model._flat.device)
        dout = dout.to(torch.bfloat16).contiguous()
        if dzs is not None:
            dzs = dzs.to(torch.bfloat16).contiguous()
        fresh = model._bind_grads()
        if fresh and model.learn_align and dzs is None:
            model._gflat.zero_()  # projector gradients are not produced without an alignment loss
        ev = None
        if model._events is not None:
            ev = (C.c_void_p * len(model._events))(*[e.cuda_event for e in model._events])
        L.call("vaw_dit_backward", C.byref(cfg), model._flat.data_ptr(), model._shadow.data_ptr(),
               model._gflat.data_ptr(), model._ws.data_ptr(), dout.data_ptr(), L.ptr(dzs), L.ptr(ctx.y),
               0 if fresh else 1, ev, L.stream_ptr())
        if model._post_backward is not None:
            model._post_backward()
        return None, None, None, None, None, None




def DiT_S(image_size, patch_size, in_channels, class_dropout_prob, num_classes, learn_sigma, **kwargs):
    return DiT(image_size=image_size, patch_size=patch_size, in_channels=in_channels, hidden_size=384, depth=12,
               num_heads=6, class_dropout_prob=class_dropout_prob, num_classes=num_classes, learn_sigma=learn_sigma,
               **kwargs)


def DiT_B(image_size, patch_size, in_channels, class_dropout_prob, num_classes, learn_sigma, **kwargs):
    return DiT(image_size=image_size, patch_size=patch_size, in_channels=in_channels, hidden_size=768, depth=12,
               num_heads=12, class_dropout_prob=class_dropout_prob, num_classes=num_classes, learn_sigma=learn_sigma,
               **kwargs)


def DiT_L(image_size, patch_size, in_channels, class_dropout_prob, num_classes, learn_sigma, **kwargs):
    return DiT(image_size=image_size, patch_size=patch_size, in_channels=in_channels, hidden_size=1024, depth=24,
               num_heads=16, class_dropout_prob=class_dropout_prob, num_classes=num_classes, learn_sigma=learn_sigma,
               **kwargs)


def DiT_XL(image_size, patch_size, in_channels, class_dropout_prob, num_classes, learn_sigma, **kwargs):
    return DiT(image_size=image_size, patch_size=patch_size, in_channels=in_channels, hidden_size=1152, depth=28,
               num_heads=16, class_dropout_prob=class_dropout_prob, num_classes=num_classes, learn_sigma=learn_sigma,
               **kwargs)


DiT_models = {"DiT-S": DiT_S, "DiT-B": DiT_B, "DiT-L": DiT_L, "DiT-XL": DiT_XL}
