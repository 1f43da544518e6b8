"""LayerNorm forward (plain and with the previous branch folded in) at the DiT-XL/2 shape: us and GB/s."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200")]
import torch
from vaw_b200 import _lib as L
P = C.c_void_p
L.register("vaw_ln_fwd", [P] * 3 + [C.c_longlong, C.c_int] + [P] * 5 + [C.c_int, C.c_int, C.c_float, P])
L.register("vaw_ln_fwd_res", [P] * 3 + [C.c_longlong] + [P] * 3 + [C.c_longlong, C.c_int] + [P] * 3 + [C.c_int, C.c_int, C.c_float, P])
B, T, D = int(os.environ.get("B", 64)), 256, int(os.environ.get("D", 1152)); M = B * T
n_buf = 4   # rotate buffers: 6 x 340 MB > L2
xs = [torch.randn(M, D, device="cuda") for _ in range(n_buf)]
brs = [torch.randn(M, D, device="cuda").bfloat16() for _ in range(n_buf)]
xo = [torch.empty(M, D, device="cuda") for _ in range(n_buf)]
ys = [torch.empty(M, D, device="cuda", dtype=torch.bfloat16) for _ in range(n_buf)]
mod = torch.randn(B, 6 * D, device="cuda") * 0.3
mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
def plain(i):
    L.call("vaw_ln_fwd", xs[i].data_ptr(), mod[:, 3 * D:].data_ptr(), mod[:, 4 * D:].data_ptr(), 6 * D, T, None, None,
           ys[i].data_ptr(), mean.data_ptr(), rstd.data_ptr(), M, D, 1e-6, L.stream_ptr())
def res(i):
    L.call("vaw_ln_fwd_res", xs[i].data_ptr(), brs[i].data_ptr(), mod[:, 2 * D:].data_ptr(), 6 * D, xo[i].data_ptr(),
           mod[:, 3 * D:].data_ptr(), mod[:, 4 * D:].data_ptr(), 6 * D, T, ys[i].data_ptr(), mean.data_ptr(), rstd.data_ptr(),
           M, D, 1e-6, L.stream_ptr())
for name, f, bytes_ in (("ln_fwd", plain, M * D * 6), ("ln_fwd_res", res, M * D * 12)):
    for i in range(n_buf): f(i)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for it in range(60): f(it % n_buf)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 60 * 1e3
    print(f"{name} M={M} D={D}: {us:.1f} us  {bytes_ / us / 1e3:.0f} GB/s")
