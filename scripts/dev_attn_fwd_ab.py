"""A/B timing of the tcgen05 attention forward at the DiT-XL/2 and U-ViT-M shapes (run twice: VAW_ATTN_EXACT=0/1)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200")]
import torch
from vaw_b200 import _lib as L
L.register("vaw_attn_fwd", [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p])
for B, T, H, hd in [(64, 256, 16, 72), (64, 258, 12, 64), (128, 256, 16, 72)]:
    qkv = (torch.randn(B, T, 3, H, hd, device="cuda") * 0.7).bfloat16()
    o = torch.empty(B, T, H, hd, device="cuda", dtype=torch.bfloat16); lse = torch.empty(B, H, T, device="cuda")
    f = lambda: L.call("vaw_attn_fwd", qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, H, hd, L.stream_ptr())
    for _ in range(5): f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(50): f()
    b.record(); torch.cuda.synchronize()
    print(f"VAW_ATTN_EXACT={os.environ.get('VAW_ATTN_EXACT', '0')} B={B} T={T} H={H} hd={hd}: {a.elapsed_time(b) / 50 * 1e3:.1f} us")
