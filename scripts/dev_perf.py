"""GPU perf probe: per-kernel timings at DiT-XL/2 B=64 shapes + whole step.  Prints us and TFLOP/s or GB/s."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
import torch
from vaw_b200 import _lib as L
from gpu_util import run_gemm
import vaw_b200.models.dit  # registers engine sigs
dev = "cuda"
L.register("vaw_attn_fwd", [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p])
L.register("vaw_attn_bwd", [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_void_p])
L.register("vaw_attn_bwd_ws", [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_void_p])
L.register("vaw_ln_fwd", [C.c_void_p] * 3 + [C.c_longlong, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_float, C.c_void_p])
L.register("vaw_ln_bwd", [C.c_void_p] * 5 + [C.c_longlong] + [C.c_void_p] * 2 + [C.c_int, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p])
L.register("vaw_gate_bwd", [C.c_void_p] * 3 + [C.c_longlong] + [C.c_void_p] * 2 + [C.c_int] * 5 + [C.c_void_p])

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3

B, T, D, H = 64, 256, int(os.environ.get("D", 1152)), int(os.environ.get("H", 16))
M, Hd = B * T, 4 * D
bf = lambda *s: torch.randn(*s, device=dev).bfloat16()
ws = torch.empty(148 * 128 * 256, device=dev)
def gemm_case(name, M_, N_, K_, a_mn, b_mn, epi, tile_n=0, k_splits=0, cta_group=0, **kw):
    A = bf(K_, M_) if a_mn else bf(M_, K_); Bm = bf(K_, N_) if b_mn else bf(N_, K_)
    args = dict(kw)
    if epi in (L.EPI_F32,): args["out"] = torch.empty(M_, N_, device=dev)
    else: args["out"] = torch.empty(M_, N_, device=dev, dtype=torch.bfloat16)
    if epi in (L.EPI_GELU_TANH, L.EPI_GELU_ERF, L.EPI_SILU): args["out2"] = torch.empty(M_, N_, device=dev, dtype=torch.bfloat16)
    if epi in (L.EPI_GATE_RES, L.EPI_RES):
        args["out2"] = torch.empty(M_, N_, device=dev); args["resid"] = torch.randn(M_, N_, device=dev)
    if epi == L.EPI_GATE_RES: args["gate"] = torch.randn(B, N_, device=dev); args["rows_per_sample"] = T
    if epi in (L.EPI_DGELU_TANH, L.EPI_DGELU_ERF, L.EPI_DSILU): args["aux"] = bf(M_, N_)
    if epi != L.EPI_F32 or True: args["bias"] = torch.zeros(N_, device=dev) if epi not in (L.EPI_DGELU_TANH,) else None
    us = timeit(lambda: run_gemm(A, Bm, a_mn, b_mn, M_, N_, K_, epi, tile_n=tile_n, k_splits=k_splits, split_ws=ws if k_splits else None, cta_group=cta_group, **args))
    print(f"{name:30s} cg{cta_group} M{M_:6d} N{N_:5d} K{K_:6d} bn{tile_n:3d} ks{k_splits:3d}: {us:8.1f} us {2*M_*N_*K_/us/1e6:7.1f} TFLOP/s", flush=True)
    return us

for cg in (() if os.environ.get("SKIP_GEMM") else (1, 2)):
  for bn in (192, 256):
    print(f"--- cta_group {cg} tile_n {bn}")
    gemm_case("fwd qkv (BF16)", M, 3 * D, D, 0, 0, L.EPI_BF16, bn, 0, cg)
    gemm_case("fwd proj (GATE_RES)", M, D, D, 0, 0, L.EPI_GATE_RES, bn, 0, cg)
    gemm_case("fwd fc1 (GELU)", M, Hd, D, 0, 0, L.EPI_GELU_TANH, bn, 0, cg)
    gemm_case("fwd fc2 (GATE_RES)", M, D, Hd, 0, 0, L.EPI_GATE_RES, bn, 0, cg)
    gemm_case("dgrad fc2 (DGELU)", M, Hd, D, 0, 1, L.EPI_DGELU_TANH, bn, 0, cg)
    gemm_case("dgrad fc1 (BF16)", M, D, Hd, 0, 1, L.EPI_BF16, bn, 0, cg)
    gemm_case("dgrad proj (BF16)", M, D, D, 0, 1, L.EPI_BF16, bn, 0, cg)
    gemm_case("dgrad qkv (BF16)", M, D, 3 * D, 0, 1, L.EPI_BF16, bn, 0, cg)
    gemm_case("wgrad fc2", D, Hd, M, 1, 1, L.EPI_F32, bn, -1, cg)
    gemm_case("wgrad fc1", Hd, D, M, 1, 1, L.EPI_F32, bn, -1, cg)
    gemm_case("wgrad proj", D, D, M, 1, 1, L.EPI_F32, bn, -1, cg)
    gemm_case("wgrad qkv", 3 * D, D, M, 1, 1, L.EPI_F32, bn, -1, cg)
if not os.environ.get("SKIP_GEMM"):
 gemm_case("square 8192^3", 8192, 8192, 8192, 0, 0, L.EPI_BF16, 256, 0, 1)
gemm_case("square 8192^3", 8192, 8192, 8192, 0, 0, L.EPI_BF16, 256, 0, 2)
gemm_case("square 8192^3", 8192, 8192, 8192, 0, 0, L.EPI_BF16, 192, 0, 2)
gemm_case("square 8192^3", 8192, 8192, 8192, 0, 0, L.EPI_BF16, 128, 0, 2)

# memory-bound kernels
x = torch.randn(M, D, device=dev); mod = torch.randn(B, 6 * D, device=dev) * 0.1
y = torch.empty(M, D, device=dev, dtype=torch.bfloat16); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
st = L.stream_ptr
us = timeit(lambda: L.call("vaw_ln_fwd", x.data_ptr(), mod.data_ptr(), mod[:, D:].data_ptr(), 6 * D, T, None, None, y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), M, D, 1e-6, st()))
print(f"ln_fwd: {us:.1f} us  {M*D*6/us/1e3:.0f} GB/s")
dy = bf(M, D); dx = torch.randn(M, D, device=dev)
for ch in (4, 8, 16):
    part = torch.empty(B * ch * 2 * D, device=dev)
    us = timeit(lambda: L.call("vaw_ln_bwd", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), mod[:, D:].data_ptr(), 6 * D, None, dx.data_ptr(), 1, part.data_ptr(), T, B, ch, M, D, st()))
    print(f"ln_bwd chunks={ch}: {us:.1f} us  {M*D*14/us/1e3:.0f} GB/s")
    dyo = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    us = timeit(lambda: L.call("vaw_gate_bwd", dx.data_ptr(), dy.data_ptr(), mod.data_ptr(), 6 * D, dyo.data_ptr(), part.data_ptr(), T, B, ch, M, D, st()))
    print(f"gate_bwd chunks={ch}: {us:.1f} us  {M*D*8/us/1e3:.0f} GB/s")
L.register("vaw_ln_bwd_gate", [C.c_void_p] * 5 + [C.c_longlong] + [C.c_void_p] * 2 + [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p])
for ch in (8, 9, 16):
    part = torch.empty(B * ch * 2 * D, device=dev); part2 = torch.empty_like(part)
    dyo = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    us = timeit(lambda: L.call("vaw_ln_bwd_gate", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), mod[:, D:].data_ptr(), 6 * D, None, dx.data_ptr(), 1, part.data_ptr(), dy.data_ptr(), mod.data_ptr(), 6 * D, dyo.data_ptr(), part2.data_ptr(), T, B, ch, M, D, st()))
    print(f"ln_bwd_gate fused chunks={ch}: {us:.1f} us  {M*D*18/us/1e3:.0f} GB/s")
    us = timeit(lambda: L.call("vaw_ln_bwd", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), mod[:, D:].data_ptr(), 6 * D, None, dx.data_ptr(), 1, part.data_ptr(), T, B, ch, M, D, st()))
    print(f"ln_bwd chunks={ch}: {us:.1f} us")
hd = D // H
qkv = bf(B, T, 3, H, hd); o = torch.empty(B, T, H, hd, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, H, T, device=dev)
us = timeit(lambda: L.call("vaw_attn_fwd", qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, H, hd, st()))
print(f"attn_fwd: {us:.1f} us  {4*B*H*T*T*hd/us/1e6:.0f} TFLOP/s")
do = bf(B, T, H, hd); dqkv = torch.empty_like(qkv); dws = torch.empty(B * H * T, device=dev)
us = timeit(lambda: L.call("vaw_attn_bwd_ws", qkv.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), dws.data_ptr(), B, T, H, hd, st()))
print(f"attn_bwd: {us:.1f} us  {10*B*H*T*T*hd/us/1e6:.0f} TFLOP/s (5 matmuls counted)")

# whole model
from vaw_b200.models.dit import DiT
m = DiT(image_size=32, patch_size=2, in_channels=4, hidden_size=D, depth=28 if D == 1152 else 12, num_heads=H, class_dropout_prob=0.0, num_classes=1000).to(dev)
with torch.no_grad():
    for p in m.parameters():
        if p.requires_grad and p.abs().sum() == 0: p.normal_(0, 0.02)
xx = torch.randn(B, 4, 32, 32, device=dev); tt = torch.rand(B, device=dev) * 999; yy = torch.randint(0, 1000, (B,), device=dev)
g = torch.randn(B, 4, 32, 32, device=dev).bfloat16()
def step():
    out, _ = m(xx, tt, yy); out.backward(g)
us = timeit(step, iters=5)
with torch.no_grad():
    usf = timeit(lambda: m(xx, tt, yy), iters=5)
print(f"model fwd+bwd {us/1e3:.2f} ms ({B/us*1e6:.0f} img/s); fwd {usf/1e3:.2f} ms; bwd {(us-usf)/1e3:.2f} ms")
