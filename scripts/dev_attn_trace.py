"""Timeline of one CTA of the tcgen05 attention backward (globaltimer stamps, see vaw_attn_set_trace)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200")]
import torch
from vaw_b200 import _lib as L
dev = "cuda"
L.register("vaw_attn_fwd", [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p])
L.register("vaw_attn_bwd", [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_void_p])
L.register("vaw_attn_bwd_ws", [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_void_p])
L.register("vaw_attn_set_trace", [C.c_void_p])
B, T, H, hd = 64, 256, 16, 72
bf = lambda *s: torch.randn(*s, device=dev).bfloat16()
qkv = bf(B, T, 3, H, hd); o = torch.empty(B, T, H, hd, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, H, T, device=dev)
do = bf(B, T, H, hd); dqkv = torch.empty_like(qkv); dws = torch.empty(B * H * T, device=dev)
st = L.stream_ptr
L.call("vaw_attn_fwd", qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, H, hd, st())
for _ in range(3):
    L.call("vaw_attn_bwd_ws", qkv.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), dws.data_ptr(), B, T, H, hd, st())
tr = torch.zeros(128, dtype=torch.int64, device=dev)
L.call("vaw_attn_set_trace", tr.data_ptr())
L.call("vaw_attn_bwd_ws", qkv.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), dws.data_ptr(), B, T, H, hd, st())
torch.cuda.synchronize()
L.call("vaw_attn_set_trace", None)
t = tr.cpu().tolist()
t0 = min(x for x in t if x > 0)
for role, name in ((0, "EW "), (1, "MMA")):
    xs = [x - t0 for x in t[role * 64:(role + 1) * 64] if x > 0]
    print(name, len(xs), " ".join(f"{x/1e3:.2f}" for x in xs))
