import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
import torch
from vaw_b200 import _lib as L
from gpu_util import run_gemm
dev = "cuda"
def timeit(fn, iters=10, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3
M = N = K = 8192
A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
for cg in (1, 2):
    for bn in (256, 128):
        us = timeit(lambda: run_gemm(A, B, 0, 0, M, N, K, L.EPI_BF16, out=out, tile_n=bn, cta_group=cg))
        print(f"VAW_DBG={os.environ.get('VAW_DBG','0')} cg{cg} bn{bn}: {us:.1f} us {2*M*N*K/us/1e6:.1f} TFLOP/s", flush=True)
