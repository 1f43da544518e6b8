"""Small target for `ncu --set full`: a few launches of each hot kernel at DiT-XL/2 B=64 shapes."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
import torch
from vaw_b200 import _lib as L
from gpu_util import run_gemm
dev = "cuda"
L.register("vaw_attn_fwd", [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p])
L.register("vaw_attn_bwd", [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_void_p])
L.register("vaw_attn_bwd_ws", [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_void_p])
L.register("vaw_ln_fwd", [C.c_void_p] * 3 + [C.c_longlong, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_float, C.c_void_p])
L.register("vaw_ln_bwd", [C.c_void_p] * 5 + [C.c_longlong] + [C.c_void_p] * 2 + [C.c_int, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p])
L.register("vaw_gate_bwd", [C.c_void_p] * 3 + [C.c_longlong] + [C.c_void_p] * 2 + [C.c_int] * 5 + [C.c_void_p])
B, T, D, H = 64, 256, 1152, 16
M, Hd = B * T, 4 * D
bf = lambda *s: torch.randn(*s, device=dev).bfloat16()
st = L.stream_ptr
reps = int(os.environ.get("REPS", 2))
for _ in range(reps):
    # GEMMs: fc1 fwd (GELU), proj fwd (GATE_RES), wgrad qkv (tail split), dgrad fc2 (DGELU)
    A = bf(M, D); W1 = bf(Hd, D); o1 = torch.empty(M, Hd, device=dev, dtype=torch.bfloat16); o2 = torch.empty_like(o1)
    bias = torch.zeros(Hd, device=dev)
    run_gemm(A, W1, 0, 0, M, Hd, D, L.EPI_GELU_TANH, out=o1, out2=o2, bias=bias)
    Wp = bf(D, D); y = torch.empty(M, D, device=dev, dtype=torch.bfloat16); xo = torch.empty(M, D, device=dev)
    resid = torch.randn(M, D, device=dev); gate = torch.randn(B, D, device=dev)
    run_gemm(A, Wp, 0, 0, M, D, D, L.EPI_GATE_RES, out=y, out2=xo, bias=bias[:D], resid=resid, gate=gate, rows_per_sample=T)
    dq = bf(M, 3 * D); ws = torch.empty(148 * 128 * 256, device=dev); gw = torch.empty(3 * D, D, device=dev)
    run_gemm(dq, A, 1, 1, 3 * D, D, M, L.EPI_F32, out=gw, k_splits=-1, split_ws=ws)
    dyb = bf(M, D); W2 = bf(D, Hd); dh = torch.empty(M, Hd, device=dev, dtype=torch.bfloat16)
    run_gemm(dyb, W2, 0, 1, M, Hd, D, L.EPI_DGELU_TANH, out=dh, aux=o1)
    # LN / gate
    x = torch.randn(M, D, device=dev); mod = torch.randn(B, 6 * D, device=dev) * 0.1
    yn = torch.empty(M, D, device=dev, dtype=torch.bfloat16); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
    L.call("vaw_ln_fwd", x.data_ptr(), mod.data_ptr(), mod[:, D:].data_ptr(), 6 * D, T, None, None, yn.data_ptr(), mean.data_ptr(), rstd.data_ptr(), M, D, 1e-6, st())
    dx = torch.randn(M, D, device=dev); part = torch.empty(B * 8 * 2 * D, device=dev)
    L.call("vaw_ln_bwd", dyb.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), mod[:, D:].data_ptr(), 6 * D, None, dx.data_ptr(), 1, part.data_ptr(), T, B, 8, M, D, st())
    dyo = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    L.call("vaw_gate_bwd", dx.data_ptr(), dyb.data_ptr(), mod.data_ptr(), 6 * D, dyo.data_ptr(), part.data_ptr(), T, B, 8, M, D, st())
    hd = D // H
    qkv = bf(B, T, 3, H, hd); o = torch.empty(B, T, H, hd, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, H, T, device=dev)
    L.call("vaw_attn_fwd", qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, H, hd, st())
    do = bf(B, T, H, hd); dqkv = torch.empty_like(qkv); dws = torch.empty(B * H * T, device=dev)
    L.call("vaw_attn_bwd_ws", qkv.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), dws.data_ptr(), B, T, H, hd, st())
    torch.cuda.synchronize()
print("done")
