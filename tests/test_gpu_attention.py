"""GPU tier: flash-style attention forward/backward through the C ABI vs torch SDPA in fp32 on the same bf16 inputs."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from gpu_util import relerr
from vaw_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda"
L.register("vaw_attn_fwd", [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p])
L.register("vaw_attn_bwd", [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_void_p])
L.register("vaw_attn_bwd_ws", [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_void_p])


# T in (256, 264]: tensor cores on the leading 256 tokens + the border strip (attention_border.cu): 258 = U-ViT's
# sequence, 257 = the ViT teacher's; 265 falls back to the mma.sync kernels
@pytest.mark.parametrize("B,T,H,hd", [(2, 256, 3, 64), (2, 256, 2, 72), (3, 258, 2, 64), (2, 64, 2, 72), (1, 100, 1, 64),
                                      (1, 1, 1, 64), (2, 17, 2, 72), (64, 256, 16, 72), (2, 257, 12, 64), (2, 258, 3, 72),
                                      (1, 264, 2, 64), (1, 265, 1, 64), (64, 258, 12, 64)])
def test_forward_backward(B, T, H, hd):
    torch.manual_seed(0)
    qkv = (torch.randn(B, T, 3, H, hd, device=DEV) * 0.7).bfloat16()
    o = torch.full((B, T, H, hd), float("nan"), device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device=DEV)
    L.call("vaw_attn_fwd", qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, H, hd, L.stream_ptr())
    q, k, v = [qkv[:, :, i].float().permute(0, 2, 1, 3).requires_grad_(True) for i in range(3)]
    ref = F.scaled_dot_product_attention(q, k, v)
    assert relerr(o.permute(0, 2, 1, 3), ref) < 6e-3
    # log-sum-exp (log2 domain) against the definition
    s = (q @ k.transpose(-1, -2)) * (hd ** -0.5)
    assert relerr(lse, torch.logsumexp(s, -1) * 1.4426950408889634) < 1e-3
    do = torch.randn_like(o)
    dqkv = torch.full_like(qkv, float("nan"))
    L.call("vaw_attn_bwd", qkv.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), B, T, H, hd,
           L.stream_ptr())
    # the variant with a caller-provided Delta scratch (used by the engines) must give the same result
    dqkv_ws = torch.full_like(qkv, float("nan"))
    ws = torch.empty(B * H * T, device=DEV)
    L.call("vaw_attn_bwd_ws", qkv.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv_ws.data_ptr(),
           ws.data_ptr(), B, T, H, hd, L.stream_ptr())
    ref.backward(do.float().permute(0, 2, 1, 3))
    scale = max(g.abs().max().item() for g in (q.grad, k.grad, v.grad))
    for res in (dqkv, dqkv_ws):
        for i, g in enumerate((q.grad, k.grad, v.grad)):
            got = res[:, :, i].permute(0, 2, 1, 3).float()
            # relative to the tensor's norm, with a floor for gradients that are analytically zero (dq, dk at T = 1)
            assert (got - g).norm().item() <= 1.2e-2 * g.norm().item() + 1e-4 * scale * g.numel() ** 0.5


@pytest.mark.parametrize("T,hd,peak_key", [(256, 72, 200), (256, 64, 77), (130, 72, 129), (258, 64, 150)])
def test_forward_rows_whose_maximum_is_far_above_the_estimate(T, hd, peak_key):
    """The single-pass softmax takes its reference point from the first 32 keys of each half; a row whose true maximum is
    more than 2^80 above that estimate makes the CTA fall back to the exact two-pass path.  Half the heads get such rows
    (a key far outside the first chunks with a huge logit), the others stay on the fast path: both must match SDPA."""
    torch.manual_seed(5)
    B, H = 2, 4
    qkv = (torch.randn(B, T, 3, H, hd, device=DEV) * 0.5).bfloat16()
    # head 1 and 3: key `peak_key` is aligned with every query of rows 3.. and 40x longer -> logit ~ +150 nats there
    for h in (1, 3):
        qkv[:, 3:, 1, h] *= 0.05                                   # all other keys: tiny logits
        qkv[:, 3:, 0, h] = (torch.ones(hd, device=DEV) * 4.0).bfloat16()
        qkv[:, peak_key, 1, h] = (torch.ones(hd, device=DEV) * 5.0).bfloat16()
    o = torch.full((B, T, H, hd), float("nan"), device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device=DEV)
    L.call("vaw_attn_fwd", qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, H, hd, L.stream_ptr())
    q, k, v = [qkv[:, :, i].float().permute(0, 2, 1, 3) for i in range(3)]
    s = (q @ k.transpose(-1, -2)) * (hd ** -0.5)
    gap = (s.max(-1).values - s[..., :32].max(-1).values) * 1.4426950408889634
    assert gap[:, 1].max().item() > 100 and gap[:, 0].max().item() < 40      # the construction does what it says
    ref = F.scaled_dot_product_attention(q, k, v)
    assert not torch.isnan(o.float()).any()
    assert relerr(o.permute(0, 2, 1, 3), ref) < 6e-3
    assert relerr(lse, torch.logsumexp(s, -1) * 1.4426950408889634) < 1e-3


def test_exact_softmax_switch_gives_the_same_result_in_a_subprocess():
    """VAW_ATTN_EXACT=1 forces the two-pass softmax for every CTA (read once per process): same outputs to bf16
    rounding (the reference point of the exponentials differs, nothing else)."""
    import os, subprocess, sys
    code = (
        "import sys, ctypes as C, torch; sys.path[:0] = %r\n"
        "from vaw_b200 import _lib as L\n"
        "L.register('vaw_attn_fwd', [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p])\n"
        "B, T, H, hd = 3, 256, 4, 72\n"
        "torch.manual_seed(1)\n"
        "qkv = (torch.randn(B, T, 3, H, hd, device='cuda') * 0.7).bfloat16()\n"
        "o = torch.empty(B, T, H, hd, device='cuda', dtype=torch.bfloat16); lse = torch.empty(B, H, T, device='cuda')\n"
        "L.call('vaw_attn_fwd', qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, H, hd, L.stream_ptr())\n"
        "torch.save((o.cpu(), lse.cpu()), sys.argv[1])\n"
        "print('ok')\n" % ([os.path.dirname(os.path.abspath(L.__file__)) + "/..", os.path.dirname(os.path.abspath(__file__))],))
    import tempfile
    outs = []
    with tempfile.TemporaryDirectory() as d:
        for exact in ("0", "1"):
            path = os.path.join(d, f"o{exact}.pt")
            r = subprocess.run([sys.executable, "-c", code, path], env=dict(os.environ, VAW_ATTN_EXACT=exact),
                               capture_output=True, text=True, timeout=300)
            assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
            outs.append(torch.load(path))
    (o0, l0), (o1, l1) = outs
    assert relerr(o0, o1) < 4e-3 and relerr(l0, l1) < 1e-5


@pytest.mark.parametrize("legacy", ["0", "1"])
def test_both_attention_paths_in_a_subprocess(legacy):
    """The mma.sync kernels stay in the library for T > 256 (U-ViT); VAW_ATTN_LEGACY=1 forces them for every shape so
    both implementations are exercised on the DiT shapes (the switch is read once per process)."""
    import os, subprocess, sys
    code = (
        "import sys, ctypes as C, torch; sys.path[:0] = %r\n"
        "from vaw_b200 import _lib as L\n"
        "import torch.nn.functional as F\n"
        "L.register('vaw_attn_fwd', [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p])\n"
        "L.register('vaw_attn_bwd', [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_void_p])\n"
        "B, T, H, hd = 3, 256, 4, 72\n"
        "torch.manual_seed(1)\n"
        "qkv = (torch.randn(B, T, 3, H, hd, device='cuda') * 0.7).bfloat16()\n"
        "o = torch.empty(B, T, H, hd, device='cuda', dtype=torch.bfloat16); lse = torch.empty(B, H, T, device='cuda')\n"
        "L.call('vaw_attn_fwd', qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, H, hd, L.stream_ptr())\n"
        "q, k, v = [qkv[:, :, i].float().permute(0, 2, 1, 3).requires_grad_(True) for i in range(3)]\n"
        "ref = F.scaled_dot_product_attention(q, k, v)\n"
        "do = torch.randn_like(o); dqkv = torch.empty_like(qkv)\n"
        "L.call('vaw_attn_bwd', qkv.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), B, T, H, hd, L.stream_ptr())\n"
        "ref.backward(do.float().permute(0, 2, 1, 3))\n"
        "e = [((dqkv[:, :, i].permute(0, 2, 1, 3).float() - g.grad).norm() / g.grad.norm()).item() for i, g in enumerate((q, k, v))]\n"
        "eo = ((o.permute(0, 2, 1, 3).float() - ref).norm() / ref.norm()).item()\n"
        "assert eo < 6e-3 and max(e) < 1.2e-2, (eo, e)\n"
        "print('ok')\n" % ([os.path.dirname(os.path.abspath(L.__file__ )) + "/..", os.path.dirname(os.path.abspath(__file__))],))
    env = dict(os.environ, VAW_ATTN_LEGACY=legacy)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_unsupported_head_dim_fails_loudly():
    x = torch.zeros(1, 16, 3, 1, 48, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(L.VawError):
        L.call("vaw_attn_fwd", x.data_ptr(), x.data_ptr(), x.data_ptr(), 1, 16, 1, 48, L.stream_ptr())


@pytest.mark.parametrize("B,T,H,hd", [(2, 100, 2, 72), (1, 256, 3, 64), (2, 37, 1, 72), (2, 258, 2, 64), (1, 257, 2, 72)])
def test_attention_outputs_stay_inside_their_buffers(B, T, H, hd):
    """Guard bands around o, lse and dqkv (TMA stores clip at the tensor bounds; ragged T exercises the clipping)."""
    torch.manual_seed(3)
    pad = 4096
    qkv = (torch.randn(B, T, 3, H, hd, device=DEV) * 0.7).bfloat16()

    def guarded(n, dtype):
        buf = torch.full((n + 2 * pad,), 7.0, device=DEV, dtype=dtype)
        return buf, buf[pad:pad + n]

    bo, o = guarded(B * T * H * hd, torch.bfloat16)
    bl, lse = guarded(B * H * T, torch.float32)
    L.call("vaw_attn_fwd", qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, H, hd, L.stream_ptr())
    bd, dqkv = guarded(qkv.numel(), torch.bfloat16)
    bw, ws = guarded(B * H * T, torch.float32)
    do = torch.randn(B, T, H, hd, device=DEV).bfloat16()
    L.call("vaw_attn_bwd_ws", qkv.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), ws.data_ptr(),
           B, T, H, hd, L.stream_ptr())
    torch.cuda.synchronize()
    for buf in (bo, bl, bd, bw):
        assert bool((buf[:pad] == 7.0).all()) and bool((buf[-pad:] == 7.0).all())
    assert not torch.isnan(o.float()).any() and not torch.isnan(dqkv.float()).any()
