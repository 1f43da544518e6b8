"""GPU dev check for the tcgen05 GEMM: every layout x a few epilogues vs torch, with diagnostics."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "variance-aware-weight_b200"))
import torch
from vaw_b200 import _lib as L

torch.manual_seed(0)
dev = "cuda"

def run_gemm(A, B, a_mn, b_mn, M, N, K, epi, out=None, out2=None, bias=None, resid=None, gate=None, aux=None,
             rows_per_sample=1, accumulate=0, tile_n=0):
    g = L.GemmArgs()
    g.A, g.B = A.data_ptr(), B.data_ptr()
    g.lda, g.ldb = A.stride(0), B.stride(0)
    g.a_mn, g.b_mn = a_mn, b_mn
    g.M, g.N, g.K = M, N, K
    g.epilogue = epi
    g.out = L.ptr(out); g.out2 = L.ptr(out2); g.bias = L.ptr(bias); g.resid = L.ptr(resid)
    g.gate = L.ptr(gate); g.aux = L.ptr(aux)
    g.ldo = 0; g.ldg = 0
    g.rows_per_sample = rows_per_sample; g.accumulate = accumulate; g.tile_n = tile_n
    L.call("vaw_gemm_bf16", C.byref(g), L.stream_ptr())

def report(name, got, ref):
    got = got.float(); ref = ref.float()
    err = (got - ref).abs()
    rel = err.max().item() / (ref.abs().max().item() + 1e-9)
    bad = (err > 1e-2 * ref.abs().max()).float().mean().item()
    print(f"{name:60s} max_abs={err.max().item():.4e} rel_to_max={rel:.3e} bad_frac={bad:.4f}", flush=True)
    if bad > 0:
        e = (err > 1e-2 * ref.abs().max())
        rows = e.any(1).nonzero().flatten()[:8].tolist(); cols = e.any(0).nonzero().flatten()[:16].tolist()
        print("   first bad rows", rows, "cols", cols)
    return rel

fails = 0
for (M, N, K) in [(256, 384, 128), (512, 1152, 1152), (200, 136, 72), (16384, 1152, 1152)]:
    for tile_n in (128, 192, 256):
        for a_mn in (0, 1):
            for b_mn in (0, 1):
                Af = torch.randn(M, K, device=dev); Bf = torch.randn(N, K, device=dev)
                A = Af.bfloat16(); B = Bf.bfloat16()
                ref = A.float() @ B.float().t()
                Ain = A.t().contiguous() if a_mn else A
                Bin = B.t().contiguous() if b_mn else B
                if (a_mn and M % 8) or (b_mn and N % 8) or K % 8:
                    continue
                out = torch.empty(M, N, device=dev, dtype=torch.float32)
                try:
                    run_gemm(Ain, Bin, a_mn, b_mn, M, N, K, L.EPI_F32, out=out, tile_n=tile_n)
                    torch.cuda.synchronize()
                    r = report(f"M{M} N{N} K{K} bn{tile_n} a_mn{a_mn} b_mn{b_mn} F32", out, ref)
                    fails += r > 2e-3
                except Exception as e:
                    print("EXC", M, N, K, tile_n, a_mn, b_mn, repr(e)[:300]); fails += 1
                    sys.exit(2)

# epilogues (K-major), one shape
M, N, K = 1024, 1152, 384
A = torch.randn(M, K, device=dev).bfloat16(); B = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
bias = torch.randn(N, device=dev)
acc = A.float() @ B.float().t() + bias
o = torch.empty(M, N, device=dev, dtype=torch.bfloat16); o2 = torch.empty_like(o)
run_gemm(A, B, 0, 0, M, N, K, L.EPI_BF16, out=o, bias=bias); torch.cuda.synchronize()
fails += report("EPI_BF16", o, acc) > 1e-2
run_gemm(A, B, 0, 0, M, N, K, L.EPI_GELU_TANH, out=o, out2=o2, bias=bias); torch.cuda.synchronize()
fails += report("EPI_GELU_TANH pre", o, acc) > 1e-2
fails += report("EPI_GELU_TANH act", o2, torch.nn.functional.gelu(acc.bfloat16().float(), approximate="tanh")) > 1e-2
run_gemm(A, B, 0, 0, M, N, K, L.EPI_GELU_ERF, out=o, out2=o2, bias=bias); torch.cuda.synchronize()
fails += report("EPI_GELU_ERF act", o2, torch.nn.functional.gelu(acc.bfloat16().float())) > 1e-2
run_gemm(A, B, 0, 0, M, N, K, L.EPI_SILU, out=o, out2=o2, bias=bias); torch.cuda.synchronize()
fails += report("EPI_SILU act", o2, torch.nn.functional.silu(acc.bfloat16().float())) > 1e-2
resid = torch.randn(M, N, device=dev); rps = 256
gate = torch.randn(M // rps, N, device=dev)
xo = torch.empty(M, N, device=dev)
run_gemm(A, B, 0, 0, M, N, K, L.EPI_GATE_RES, out=o, out2=xo, bias=bias, resid=resid, gate=gate, rows_per_sample=rps)
torch.cuda.synchronize()
fails += report("EPI_GATE_RES y", o, acc) > 1e-2
fails += report("EPI_GATE_RES x", xo, resid + gate.repeat_interleave(rps, 0) * acc.bfloat16().float()) > 1e-2
run_gemm(A, B, 0, 0, M, N, K, L.EPI_RES, out2=xo, bias=bias, resid=resid); torch.cuda.synchronize()
fails += report("EPI_RES x", xo, resid + acc.bfloat16().float()) > 1e-2
aux = torch.randn(M, N, device=dev).bfloat16()
ha = aux.float().requires_grad_(True)
torch.nn.functional.gelu(ha, approximate="tanh").sum().backward()
run_gemm(A, B, 0, 0, M, N, K, L.EPI_DGELU_TANH, out=o, aux=aux); torch.cuda.synchronize()
fails += report("EPI_DGELU_TANH", o, (acc - bias) * ha.grad) > 1e-2
ha.grad = None; torch.nn.functional.gelu(ha).sum().backward()
run_gemm(A, B, 0, 0, M, N, K, L.EPI_DGELU_ERF, out=o, aux=aux); torch.cuda.synchronize()
fails += report("EPI_DGELU_ERF", o, (acc - bias) * ha.grad) > 1e-2
ha.grad = None; torch.nn.functional.silu(ha).sum().backward()
run_gemm(A, B, 0, 0, M, N, K, L.EPI_DSILU, out=o, aux=aux); torch.cuda.synchronize()
fails += report("EPI_DSILU", o, (acc - bias) * ha.grad) > 1e-2
of = torch.ones(M, N, device=dev)
run_gemm(A, B, 0, 0, M, N, K, L.EPI_F32, out=of, accumulate=1); torch.cuda.synchronize()
fails += report("EPI_F32 accumulate", of, acc - bias + 1) > 2e-3

# timing
def bench(M, N, K, a_mn, b_mn, epi, tile_n=0, iters=20):
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if epi == L.EPI_F32 else torch.bfloat16)
    for _ in range(3): run_gemm(A, B, a_mn, b_mn, M, N, K, epi, out=out, tile_n=tile_n)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): run_gemm(A, B, a_mn, b_mn, M, N, K, epi, out=out, tile_n=tile_n)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    print(f"bench M{M} N{N} K{K} a_mn{a_mn} b_mn{b_mn} epi{epi} bn{tile_n}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    # torch reference
    A2 = torch.randn(M, K, device=dev).bfloat16(); B2 = torch.randn(N, K, device=dev).bfloat16()
    for _ in range(3): A2 @ B2.t()
    s.record()
    for _ in range(iters): A2 @ B2.t()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    print(f"   cuBLAS same shape: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)

for bn in (128, 192, 256):
    bench(16384, 4608, 1152, 0, 0, L.EPI_BF16, bn)
bench(16384, 1152, 4608, 0, 0, L.EPI_BF16, 192)
bench(16384, 3456, 1152, 0, 0, L.EPI_BF16, 192)
bench(16384, 1152, 4608, 0, 1, L.EPI_BF16, 192)
bench(4608, 1152, 16384, 1, 1, L.EPI_F32, 192)
bench(8192, 8192, 8192, 0, 0, L.EPI_BF16, 256)
print("FAILS", fails)
sys.exit(1 if fails else 0)
