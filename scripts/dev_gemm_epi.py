"""Experiment: cost of the GEMM epilogue pieces at the short-K DiT shapes (VAW_DBG=8 / 16, see gemm_sm100.cu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
import torch
from vaw_b200 import _lib as L
from gpu_util import run_gemm
dev = "cuda"
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3
M, D = 16384, 1152
bf = lambda *s: torch.randn(*s, device=dev).bfloat16()
for name, N, K, epi in (("qkv BF16", 3 * D, D, L.EPI_BF16), ("fc1 GELU", 4 * D, D, L.EPI_GELU_TANH), ("dgrad fc2 DGELU", 4 * D, D, L.EPI_DGELU_TANH)):
    A = bf(M, K); W = bf(N, K); o1 = torch.empty(M, N, device=dev, dtype=torch.bfloat16); o2 = torch.empty_like(o1)
    bias = torch.zeros(N, device=dev); aux = bf(M, N)
    for cg in (2,):
        us = timeit(lambda: run_gemm(A, W, 0, 0, M, N, K, epi, out=o1, out2=o2 if epi == L.EPI_GELU_TANH else None, bias=bias if epi != L.EPI_DGELU_TANH else None, aux=aux if epi == L.EPI_DGELU_TANH else None, tile_n=256, cta_group=cg))
        print(f"VAW_DBG={os.environ.get('VAW_DBG','0'):>2s} {name:16s} cg{cg}: {us:7.1f} us {2*M*N*K/us/1e6:7.1f} TF", flush=True)
