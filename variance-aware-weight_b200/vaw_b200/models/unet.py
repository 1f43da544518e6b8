"""ADM-style UNet with the reference's constructors, forward signature and state-dict layout.

SURVEY §8 row a17 scopes this model as "API only": the UNet is BASELINE config 1 (CIFAR-10 32x32, a CPU-runnable
parity configuration), its convolutions are not the dense contractions the B200 kernels target, so it stays a PyTorch
module (cuDNN convolutions) that is fed by the B200 diffusion kernels (K1 q_sample/target, K2 weighted MSE, K3 sampler).
Interface mirrored from /root/reference/models/unet.py: UNetModel(...) (:397-687), create_unet_model (:921-981) and
the UNet_32 / ADM_* / UNet_64 / LDM factories (:983-1031); parameter names follow the reference so checkpoints load
(input_blocks.k.0.in_layers.{0,2}, .emb_layers.1, .out_layers.{0,3}, .skip_connection, input_blocks.k.1.{norm,qkv,proj_out},
middle_block.{0,1,2}, output_blocks.k.*, out.{0,2}, time_embed.{0,2}, label_emb).

Written as a table-driven builder: `_plan()` turns the hyper-parameters into a list of stage descriptions, which
`_make_stage()` instantiates.  Differences from the reference that do not change results in fp32: attention uses
F.scaled_dot_product_attention (same softmax(q k^T / sqrt(c)) v), and no activation checkpointing / forced fp16
autocast inside the attention block (unet.py:297,302).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _sinusoid(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    a = t[:, None].float() * freqs[None]
    e = torch.cat([a.cos(), a.sin()], dim=-1)
    return F.pad(e, (0, dim % 2))


class _GN32(nn.GroupNorm):
    """GroupNorm(32, C) evaluated in fp32 (tools/nn.py:17-19)."""

    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


def _zeroed(m):
    for p in m.parameters():
        nn.init.zeros_(p)
    return m


class _Resample(nn.Module):
    """Parameter-free 2x nearest upsampling / 2x average pooling (Upsample / Downsample with use_conv=False)."""

    def __init__(self, up):
        super().__init__()
        self.up = up

    def forward(self, x):
        return F.interpolate(x, scale_factor=2, mode="nearest") if self.up else F.avg_pool2d(x, 2, 2)


class Upsample(nn.Module):
    def __init__(self, channels, use_conv, out_channels=None):
        super().__init__()
        self.channels, self.out_channels, self.use_conv = channels, out_channels or channels, use_conv
        if use_conv:
            self.conv = nn.Conv2d(channels, self.out_channels, 3, padding=1)

    def forward(self, x):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        return self.conv(x) if self.use_conv else x


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, out_channels=None):
        super().__init__()
        self.channels, self.out_channels, self.use_conv = channels, out_channels or channels, use_conv
        self.op = nn.Conv2d(channels, self.out_channels, 3, stride=2, padding=1) if use_conv else nn.AvgPool2d(2, 2)

    def forward(self, x):
        return self.op(x)


class ResBlock(nn.Module):
    """GN-SiLU-conv, timestep conditioning (additive or scale-shift), GN-SiLU-dropout-zero-conv, skip (unet.py:143-256)."""

    takes_emb = True

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_scale_shift_norm=False, up=False, down=False):
        super().__init__()
        oc = out_channels or channels
        self.channels, self.out_channels, self.use_scale_shift_norm = channels, oc, use_scale_shift_norm
        self.in_layers = nn.Sequential(_GN32(32, channels), nn.SiLU(), nn.Conv2d(channels, oc, 3, padding=1))
        self.updown = up or down
        self.h_upd = _Resample(up) if self.updown else nn.Identity()
        self.x_upd = _Resample(up) if self.updown else nn.Identity()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, 2 * oc if use_scale_shift_norm else oc))
        self.out_layers = nn.Sequential(_GN32(32, oc), nn.SiLU(), nn.Dropout(p=dropout),
                                        _zeroed(nn.Conv2d(oc, oc, 3, padding=1)))
        self.skip_connection = nn.Identity() if oc == channels else nn.Conv2d(channels, oc, 1)

    def forward(self, x, emb):
        h = self.in_layers[1](self.in_layers[0](x))
        if self.updown:
            h, x = self.h_upd(h), self.x_upd(x)
        h = self.in_layers[2](h)
        e = self.emb_layers(emb).type(h.dtype)[:, :, None, None]
        if self.use_scale_shift_norm:
            scale, shift = e.chunk(2, dim=1)
            h = self.out_layers[0](h) * (1 + scale) + shift
            h = self.out_layers[3](self.out_layers[2](self.out_layers[1](h)))
        else:
            h = self.out_layers(h + e)
        return self.skip_connection(x) + h


class AttentionBlock(nn.Module):
    """Self-attention over spatial positions (unet.py:259-306, QKVAttention :362-390 / QKVAttentionLegacy :329-359)."""

    takes_emb = False

    def __init__(self, channels, num_heads=1, num_head_channels=-1, use_new_attention_order=False):
        super().__init__()
        if num_head_channels != -1:
            assert channels % num_head_channels == 0, (
                f"q,k,v channels {channels} is not divisible by num_head_channels {num_head_channels}")
            num_heads = channels // num_head_channels
        self.channels, self.num_heads, self.new_order = channels, num_heads, use_new_attention_order
        self.norm = _GN32(32, channels)
        self.qkv = nn.Conv1d(channels, 3 * channels, 1)
        self.proj_out = _zeroed(nn.Conv1d(channels, channels, 1))

    def forward(self, x):
        b, c, *spatial = x.shape
        xf = x.reshape(b, c, -1)
        n = xf.shape[-1]
        qkv = self.qkv(self.norm(xf))
        hc = c // self.num_heads
        if self.new_order:   # [q | k | v] along channels, then heads
            q, k, v = (t.reshape(b, self.num_heads, hc, n) for t in qkv.chunk(3, dim=1))
        else:                # heads first, then [q | k | v] inside each head
            q, k, v = qkv.reshape(b, self.num_heads, 3 * hc, n).split(hc, dim=2)
        a = F.scaled_dot_product_attention(q.transpose(-1, -2), k.transpose(-1, -2), v.transpose(-1, -2))
        a = a.transpose(-1, -2).reshape(b, c, n)
        return (xf + self.proj_out(a)).reshape(b, c, *spatial)


class TimestepEmbedSequential(nn.Sequential):
    def forward(self, x, emb):
        for layer in self:
            x = layer(x, emb) if getattr(layer, "takes_emb", False) else layer(x)
        return x


class UNetModel(nn.Module):
    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=0, use_checkpoint=False,
                 use_fp16=False, num_heads=1, num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False,
                 resblock_updown=False, use_new_attention_order=False, drop_label_prob=0.0):
        super().__init__()
        if dims != 2:
            raise NotImplementedError("only 2-D UNets are used by the reference's configurations")
        self.image_size, self.in_channels, self.model_channels, self.out_channels = image_size, in_channels, model_channels, out_channels
        self.num_res_blocks, self.attention_resolutions, self.dropout = num_res_blocks, tuple(attention_resolutions), dropout
        self.channel_mult, self.conv_resample, self.num_classes = tuple(channel_mult), conv_resample, num_classes
        self.dtype = torch.float16 if use_fp16 else torch.float32
        self.num_heads, self.num_head_channels = num_heads, num_head_channels
        self.num_heads_upsample = num_heads if num_heads_upsample == -1 else num_heads_upsample
        self.drop_label_prob = drop_label_prob
        self._opts = dict(scale_shift=use_scale_shift_norm, updown=resblock_updown, new_order=use_new_attention_order)
        emb_dim = 512 if in_channels == 4 else model_channels * 4      # unet.py:482-485
        self.time_embed = nn.Sequential(nn.Linear(model_channels, emb_dim), nn.SiLU(), nn.Linear(emb_dim, emb_dim))
        if num_classes > 0:
            self.label_emb = nn.Embedding(num_classes + int(drop_label_prob > 0), emb_dim)
        down, mid, up, ch_out = self._plan()
        self.input_blocks = nn.ModuleList([self._make_stage(st, emb_dim) for st in down])
        self.middle_block = self._make_stage(mid, emb_dim)
        self.output_blocks = nn.ModuleList([self._make_stage(st, emb_dim) for st in up])
        stem = int(self.channel_mult[0] * model_channels)
        self.out = nn.Sequential(_GN32(32, ch_out), nn.SiLU(), _zeroed(nn.Conv2d(stem, out_channels, 3, padding=1)))

    # ---- architecture plan: lists of stages, each stage a list of ("kind", ...) tuples ---------------------------
    def _plan(self):
        mc, mult, nres = self.model_channels, self.channel_mult, self.num_res_blocks
        ch = int(mult[0] * mc)
        down = [[("stem", self.in_channels, ch)]]
        widths = [ch]
        ds = 1
        for level, m in enumerate(mult):
            for _ in range(nres):
                stage = [("res", ch, int(m * mc), None)]
                ch = int(m * mc)
                if ds in self.attention_resolutions:
                    stage.append(("attn", ch, self.num_heads))
                down.append(stage)
                widths.append(ch)
            if level != len(mult) - 1:
                down.append([("res", ch, ch, "down")] if self._opts["updown"] else [("down", ch)])
                widths.append(ch)
                ds *= 2
        mid = [("res", ch, ch, None), ("attn", ch, self.num_heads), ("res", ch, ch, None)]
        up = []
        for level, m in reversed(list(enumerate(mult))):
            for i in range(nres + 1):
                stage = [("res", ch + widths.pop(), int(mc * m), None)]
                ch = int(mc * m)
                if ds in self.attention_resolutions:
                    stage.append(("attn", ch, self.num_heads_upsample))
                if level and i == nres:
                    stage.append(("res", ch, ch, "up") if self._opts["updown"] else ("up", ch))
                    ds //= 2
                up.append(stage)
        return down, mid, up, ch

    def _make_stage(self, stage, emb_dim):
        layers = []
        for item in stage:
            kind = item[0]
            if kind == "stem":
                layers.append(nn.Conv2d(item[1], item[2], 3, padding=1))
            elif kind == "res":
                layers.append(ResBlock(item[1], emb_dim, self.dropout, out_channels=item[2],
                                       use_scale_shift_norm=self._opts["scale_shift"], up=item[3] == "up",
                                       down=item[3] == "down"))
            elif kind == "attn":
                layers.append(AttentionBlock(item[1], num_heads=item[2], num_head_channels=self.num_head_channels,
                                             use_new_attention_order=self._opts["new_order"]))
            elif kind == "down":
                layers.append(Downsample(item[1], self.conv_resample, out_channels=item[1]))
            elif kind == "up":
                layers.append(Upsample(item[1], self.conv_resample, out_channels=item[1]))
        return TimestepEmbedSequential(*layers)

    def token_drop(self, labels, force_drop_ids=None):
        """Label dropout; the reference draws on the CPU generator here (unet.py:649)."""
        if force_drop_ids is None:
            drop = torch.rand(labels.shape[0]).to(labels.device) < self.drop_label_prob
        else:
            drop = force_drop_ids == 1
        return torch.where(drop, self.num_classes, labels)

    def forward(self, x, timesteps, y=None, force_drop_ids=None, **kwargs):
        assert (y is not None) == (self.num_classes > 0), "must specify y if and only if the model is class-conditional"
        emb = self.time_embed(_sinusoid(timesteps, self.model_channels))
        if self.num_classes > 0:
            if (self.drop_label_prob > 0 and self.training) or force_drop_ids is not None:
                y = self.token_drop(y, force_drop_ids)
            assert y.shape == (x.shape[0],)
            emb = emb + self.label_emb(y)
        h = x.type(self.dtype)
        skips = []
        for blk in self.input_blocks:
            h = blk(h, emb)
            skips.append(h)
        h = self.middle_block(h, emb)
        for blk in self.output_blocks:
            h = blk(torch.cat([h, skips.pop()], dim=1), emb)
        return self.out(h.type(x.dtype))


_DEFAULT_MULT = {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4), 32: (1, 2, 2, 2)}


def create_unet_model(image_size, num_channels, num_res_blocks, channel_mult="", in_channels=3, num_classes=10,
                      learn_sigma=False, class_cond=True, use_checkpoint=False, attention_resolutions="16", num_heads=1,
                      num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=True, dropout=0,
                      resblock_updown=True, use_fp16=False, use_new_attention_order=True, drop_label_prob=0.0):
    if channel_mult == "":
        if image_size not in _DEFAULT_MULT:
            raise ValueError(f"unsupported image size: {image_size}")
        mult = _DEFAULT_MULT[image_size]
    else:
        mult = tuple(int(m) for m in channel_mult.split(","))
    att = tuple(image_size // int(r) for r in attention_resolutions.split(","))
    return UNetModel(image_size=image_size, in_channels=in_channels, model_channels=num_channels,
                     out_channels=in_channels * (2 if learn_sigma else 1), num_res_blocks=num_res_blocks,
                     attention_resolutions=att, dropout=dropout, channel_mult=mult,
                     num_classes=num_classes if class_cond else 0, use_checkpoint=use_checkpoint, use_fp16=use_fp16,
                     num_heads=num_heads, num_head_channels=num_head_channels, num_heads_upsample=num_heads_upsample,
                     use_scale_shift_norm=use_scale_shift_norm, resblock_updown=resblock_updown,
                     use_new_attention_order=use_new_attention_order, drop_label_prob=drop_label_prob)


def _factory(**fixed):
    def make(num_classes=10, in_channels=fixed.pop("default_in", 3), dropout=0, learn_sigma=False, class_cond=True,
             drop_label_prob=0.0, **kwargs):
        return create_unet_model(num_classes=num_classes, dropout=dropout, in_channels=in_channels,
                                 drop_label_prob=drop_label_prob, learn_sigma=learn_sigma, class_cond=class_cond,
                                 **fixed, **kwargs)
    return make


UNet_32 = _factory(image_size=32, num_channels=128, num_res_blocks=2, attention_resolutions="16,8", num_heads=4, num_head_channels=-1)
ADM_32 = _factory(image_size=32, num_channels=128, num_res_blocks=3, attention_resolutions="16,8", num_heads=1, num_head_channels=32)
ADM_64 = _factory(image_size=64, num_channels=192, num_res_blocks=3, attention_resolutions="32,16,8", num_heads=1, num_head_channels=64)
ADM_128 = _factory(image_size=128, num_channels=256, num_res_blocks=2, attention_resolutions="32,16,8", num_heads=1, num_head_channels=64)
ADM_256 = _factory(image_size=256, num_channels=256, num_res_blocks=2, attention_resolutions="32,16,8", num_heads=1, num_head_channels=64)
ADM_512 = _factory(image_size=512, num_channels=256, num_res_blocks=2, attention_resolutions="32,16,8", num_heads=1, num_head_channels=64)
UNet_64 = _factory(image_size=64, num_channels=192, num_res_blocks=3, attention_resolutions="16,8", num_heads=4, channel_mult="1,2,2,2", num_head_channels=-1)
LDM = _factory(default_in=4, image_size=32, num_channels=256, num_res_blocks=2, attention_resolutions="32,16,8", num_heads=1, channel_mult="1,2,4", num_head_channels=32)

# same table as the reference, including its "ADM-64" -> UNet_64 entry (unet.py:1023-1032)
UNet_models = {"UNet-32": UNet_32, "ADM-32": ADM_32, "ADM-64": UNet_64, "ADM-128": ADM_128, "ADM-256": ADM_256,
               "ADM-512": ADM_512, "UNet-64": UNet_64, "LDM": LDM}
