"""Locate run-to-run nondeterminism: run the DiT forward (and optionally backward) twice on the same inputs, diff the
activation workspace byte-wise and name the buffers (carve order of dit_engine.cu) that differ."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
import torch
from vaw_b200.models.dit import DiT_S, DiT_XL
from vaw_b200.tools import gaussian_diffusion as gd
from gpu_util import dezero
dev = torch.device("cuda", 0)
which, B, bwd = sys.argv[1], int(sys.argv[2]), len(sys.argv) > 3
mk = DiT_XL if which == "xl" else DiT_S
torch.manual_seed(0)
net = mk(image_size=32, patch_size=2, in_channels=4, class_dropout_prob=0.0, num_classes=1000, learn_sigma=False).to(dev)
dezero(net)
D, depth, H, T, Hd = net.hidden_size, net.depth, net.num_heads, 256, 4 * net.hidden_size
M, Kp, PPC = B * T, 16, 16
def carve():
    cur, out = 0, []
    def take(name, n, sz):
        nonlocal cur
        out.append((name, cur, n * sz)); cur += (n * sz + 255) // 256 * 256
    take("patches", M * Kp, 2); take("freq", B * 256, 2); take("t_h_pre", B * D, 2); take("t_h", B * D, 2)
    take("c_silu", B * D, 2); take("t_emb", B * D, 4); take("c", B * D, 4); take("mod_all", B * depth * 6 * D, 4)
    take("mod_final", B * 2 * D, 4)
    for i in range(2 * depth + 1): take(f"x[{i}]", M * D, 4)
    for i in range(depth):
        for nm in ("mean1", "rstd1", "mean2", "rstd2"): take(f"b{i}.{nm}", M, 4)
        take(f"b{i}.lse", B * H * T, 4)
        take(f"b{i}.xn1", M * D, 2); take(f"b{i}.qkv", M * 3 * D, 2); take(f"b{i}.attn_o", M * D, 2)
        take(f"b{i}.y_attn", M * D, 2); take(f"b{i}.xn2", M * D, 2); take(f"b{i}.h_pre", M * Hd, 2)
        take(f"b{i}.h_act", M * Hd, 2); take(f"b{i}.y_mlp", M * D, 2)
    take("meanf", M, 4); take("rstdf", M, 4); take("xnf", M * D, 2); take("out_tok", M * PPC, 2)
    for nm in ("xa", "z1_pre", "z1", "z2_pre", "z2"): take(nm, 0, 2)
    take("dx", M * D, 4); take("dy", M * D, 2); take("dh", M * Hd, 2); take("dqkv", M * 3 * D, 2); take("d_o", M * D, 2)
    take("attn_delta", B * H * T, 4); take("dxn", M * D, 2); take("dtok", M * PPC, 2)
    return out, cur
bufs, end = carve()
d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda")
x = torch.randn(B, 4, 32, 32, device=dev); y = torch.randint(0, 1000, (B,), device=dev)
t = torch.randint(0, 1000, (B,), device=dev); eps = torch.randn_like(x)
snaps = []
for r in range(2):
    for p in net.parameters(): p.grad = None
    if bwd:
        terms = d.training_losses(net, x, None, t=t, model_kwargs={"y": y}, noise=eps)
        terms["loss"].mean().backward()
    else:
        with torch.no_grad(): net(x, t.float(), y)
    torch.cuda.synchronize()
    snaps.append(net._ws.view(torch.uint8)[:end].clone())
diff = snaps[0] != snaps[1]
print(f"{which} B={B} bwd={bwd}: workspace {end/1e9:.2f} GB carved, differing bytes: {int(diff.sum())}")
shown = 0
for name, off, nbytes in bufs:
    if nbytes == 0: continue
    dd = diff[off:off + nbytes]
    n = int(dd.sum())
    if n:
        idx = dd.nonzero().flatten()
        print(f"  {name:14s} {n:10d} of {nbytes} bytes differ; first at byte {int(idx[0])}, last {int(idx[-1])}")
        shown += 1
        if shown >= 14: break
