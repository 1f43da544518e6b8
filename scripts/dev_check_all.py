"""GPU dev check: K1/K2/K3, attention, full DiT forward/backward vs the oracle.  Prints diagnostics."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "variance-aware-weight_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.nn.functional as F
from vaw_b200 import _lib as L
from vaw_b200.tools import gaussian_diffusion as gd, resample as rs
from vaw_b200.models.dit import DiT
from oracle.dit import dit_forward

dev = "cuda"; fails = 0
def relerr(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()
def check(name, a, b, tol):
    global fails
    e = relerr(a, b); bad = not (e <= tol)
    print(f"{'FAIL' if bad else 'ok  '} {name:55s} rel_l2={e:.3e} (tol {tol:g})", flush=True); fails += bad

# ---------------- K1 / K2 ----------------
torch.manual_seed(0)
for mean in ("epsilon", "start_x", "velocity", "previous_x"):
    d = gd.create_gaussian_diffusion(noise_schedule="linear", mean_type=mean, weight_type="lambda" if mean != "previous_x" else "constant")
    N = 16; x0 = torch.randn(N, 3, 32, 32, device=dev).clamp(-1, 1); eps = torch.randn_like(x0); t = torch.randint(0, 1000, (N,), device=dev)
    a = gd._extract_into_tensor(d.sqrt_alphas_cumprod, t, x0.shape); s = gd._extract_into_tensor(d.sqrt_one_minus_alphas_cumprod, t, x0.shape)
    xt_ref = a * x0 + s * eps
    xt = d.q_sample(x0, t, eps)
    ok = torch.equal(xt, xt_ref); print(("ok  " if ok else "FAIL"), f"K1 x_t bit-exact [{mean}]"); fails += (not ok)
    tgt = d.compute_target(x0, eps, t)
    if mean == "velocity":
        ref = a * eps - s * x0
    elif mean == "previous_x":
        ref = gd._extract_into_tensor(d.posterior_mean_coef1, t, x0.shape) * x0 + gd._extract_into_tensor(d.posterior_mean_coef2, t, x0.shape) * xt_ref
    else:
        ref = eps if mean == "epsilon" else x0
    ok = torch.equal(tgt, ref); print(("ok  " if ok else "FAIL"), f"K1 target bit-exact [{mean}]"); fails += (not ok)
    for dt in (torch.float32, torch.bfloat16):
        out = torch.randn_like(x0).to(dt).requires_grad_(True)
        model = lambda x, tt, **k: out
        terms = d.training_losses(model, x0, t=t, noise=eps)
        wsamp = torch.rand(N, device=dev)
        (terms["loss"] * wsamp).mean().backward()
        o2 = out.detach().clone().requires_grad_(True)
        w = gd.compute_mse_loss_weight(d.model_mean_type, d.mse_loss_weight_type, t, a[:, 0, 0, 0], s[:, 0, 0, 0])
        mse_ref = w * ((ref - o2) ** 2).mean(dim=(1, 2, 3))
        (mse_ref * wsamp).mean().backward()
        check(f"K2 mse [{mean},{dt}]", terms["mse"], mse_ref, 2e-6)
        check(f"K2 grad [{mean},{dt}]", out.grad, o2.grad, 1e-5 if dt == torch.float32 else 6e-3)

# ---------------- K3 sampler ----------------
d = gd.create_gaussian_diffusion()
for trial in range(3):
    T, H = 1000, 10
    hist = np.abs(np.random.RandomState(trial).randn(T, H)) * (0.01 + np.random.RandomState(trial + 7).rand(T, 1))
    counts = np.full(T, H)
    samp = rs.LossSecondMomentResampler(d); samp.load_history(hist, counts, dev)
    w_ref = np.sqrt(np.mean(hist ** 2, axis=-1)); w_ref /= np.sum(w_ref); w_ref *= 1 - 0.001; w_ref += 0.001 / T
    ok = np.array_equal(samp.weights(), w_ref); print(("ok  " if ok else "FAIL"), "K3 weights bit-exact"); fails += (not ok)
    np.random.seed(123 + trial); p = w_ref / np.sum(w_ref); idx_ref = np.random.choice(T, size=(64,), p=p); wr = (1 / (T * p[idx_ref])).astype(np.float32)
    st_ref = np.random.get_state()[1].copy()
    np.random.seed(123 + trial); idx, iw = samp.sample(64, dev)
    ok = np.array_equal(idx.cpu().numpy(), idx_ref) and np.array_equal(iw.cpu().numpy(), wr) and np.array_equal(np.random.get_state()[1], st_ref)
    print(("ok  " if ok else "FAIL"), "K3 sample idx/weights/rng-state bit-exact"); fails += (not ok)
us = rs.UniformSampler(d); np.random.seed(5); i1, w1 = us.sample(32, dev); np.random.seed(5); i2 = np.random.choice(1000, size=(32,), p=np.ones(1000) / 1000)
ok = np.array_equal(i1.cpu().numpy(), i2) and bool((w1 == 1).all()); print(("ok  " if ok else "FAIL"), "K3 uniform"); fails += (not ok)
# history update with duplicates
samp = rs.LossSecondMomentResampler(d); samp.load_history(np.zeros((1000, 10)), np.zeros(1000), dev)
h_ref = np.zeros((1000, 10)); c_ref = np.zeros(1000, dtype=int); rng = np.random.RandomState(0)
for step in range(40):
    ts = rng.randint(0, 30, size=64); ls = rng.rand(64).astype(np.float32)
    samp.update_with_local_losses(torch.from_numpy(ts).to(dev), torch.from_numpy(ls).to(dev))
    for tt, l in zip(ts.tolist(), ls.tolist()):
        if c_ref[tt] == 10: h_ref[tt, :-1] = h_ref[tt, 1:]; h_ref[tt, -1] = l
        else: h_ref[tt, c_ref[tt]] = l; c_ref[tt] += 1
ok = np.array_equal(samp._loss_history, h_ref) and np.array_equal(samp._loss_counts, c_ref); print(("ok  " if ok else "FAIL"), "K3 history update"); fails += (not ok)

# ---------------- attention ----------------
L.register("vaw_attn_fwd", [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p])
L.register("vaw_attn_bwd", [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_void_p])
for (B, T, H, hd) in [(2, 256, 3, 64), (2, 256, 2, 72), (3, 258, 2, 64), (2, 64, 2, 72), (1, 100, 1, 64)]:
    D = H * hd
    qkv = (torch.randn(B, T, 3, H, hd, device=dev) * 0.7).bfloat16()
    o = torch.empty(B, T, H, hd, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, H, T, device=dev)
    L.call("vaw_attn_fwd", qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, H, hd, L.stream_ptr())
    q, k, v = [qkv[:, :, i].float().permute(0, 2, 1, 3).requires_grad_(True) for i in range(3)]
    ref = F.scaled_dot_product_attention(q, k, v)
    check(f"attn fwd B{B} T{T} H{H} hd{hd}", o.permute(0, 2, 1, 3), ref, 6e-3)
    do = torch.randn_like(o)
    dqkv = torch.full_like(qkv, float("nan"))
    L.call("vaw_attn_bwd", qkv.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), B, T, H, hd, L.stream_ptr())
    ref.backward(do.float().permute(0, 2, 1, 3))
    for i, (nm, g) in enumerate((("dq", q.grad), ("dk", k.grad), ("dv", v.grad))):
        check(f"attn bwd {nm} B{B} T{T} H{H} hd{hd}", dqkv[:, :, i].permute(0, 2, 1, 3), g, 1.2e-2)

# ---------------- full DiT vs oracle (bf16 autocast) ----------------
def dit_parity(hidden, heads, depth, img, align, B=4, classes=10):
    torch.manual_seed(1)
    m = DiT(image_size=img, patch_size=2, in_channels=4, hidden_size=hidden, depth=depth, num_heads=heads, class_dropout_prob=0.0,
            num_classes=classes, learn_align=align, encoder_depth=max(1, depth // 2), z_dims=48, projector_dim=64).to(dev)
    with torch.no_grad():
        for p in m.parameters():
            if p.requires_grad and p.abs().sum() == 0: p.normal_(0, 0.02)
    x = torch.randn(B, 4, img, img, device=dev); t = torch.rand(B, device=dev) * 999; y = torch.randint(0, classes, (B,), device=dev)
    gout = torch.randn(B, 4, img, img, device=dev); T = (img // 2) ** 2
    gz = torch.randn(B, T, 48, device=dev) * 0.1 if align else None
    m.train(); out, zs = m(x, t, y)
    loss = (out.float() * gout).sum() + ((zs.float() * gz).sum() if align else 0.0)
    loss.backward()
    sd = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in m.state_dict(keep_vars=True).items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o_ref, z_ref = dit_forward(sd, x, t, y, patch_size=2, num_heads=heads, depth=depth, learn_align=align, encoder_depth=max(1, depth // 2))
    l_ref = (o_ref.float() * gout).sum() + ((z_ref.float() * gz).sum() if align else 0.0)
    l_ref.backward()
    tag = f"DiT D{hidden} h{heads} L{depth} img{img} align{int(align)}"
    check(tag + " out", out, o_ref, 2e-2)
    if align: check(tag + " zs", zs, z_ref, 2e-2)
    worst = 0; worst_name = ""
    for k, p in m.named_parameters():
        if not p.requires_grad: continue
        e = relerr(p.grad, sd[k].grad)
        if e > worst: worst, worst_name = e, k
        if e > 2e-2: print(f"   grad {k:50s} rel={e:.3e} |g|={sd[k].grad.norm().item():.3e}")
    global fails
    bad = worst > 2e-2; fails += bad
    print(f"{'FAIL' if bad else 'ok  '} {tag} worst grad rel_l2={worst:.3e} ({worst_name})", flush=True)
    # fp32 oracle as the yardstick: how far is the bf16 oracle itself from fp32?
    sd32 = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in m.state_dict(keep_vars=True).items()}
    o32, z32 = dit_forward(sd32, x, t, y, patch_size=2, num_heads=heads, depth=depth, learn_align=align, encoder_depth=max(1, depth // 2))
    ((o32 * gout).sum() + ((z32 * gz).sum() if align else 0.0)).backward()
    w_o = max(relerr(sd[k].grad, sd32[k].grad) for k, p in m.named_parameters() if p.requires_grad)
    w_m = max(relerr(p.grad, sd32[k].grad) for k, p in m.named_parameters() if p.requires_grad)
    print(f"     vs fp32 oracle: bf16-oracle worst {w_o:.3e}, engine worst {w_m:.3e}; out engine {relerr(out, o32):.3e} oracle-bf16 {relerr(o_ref, o32):.3e}")

dit_parity(128, 2, 2, 16, False)
dit_parity(144, 2, 3, 16, True)
dit_parity(384, 6, 4, 32, False, B=8, classes=1000)

# ---------------- timing: DiT-S and DiT-XL forward+backward ----------------
def time_model(name, hidden, heads, depth, B):
    m = DiT(image_size=32, patch_size=2, in_channels=4, hidden_size=hidden, depth=depth, num_heads=heads, class_dropout_prob=0.0, num_classes=1000).to(dev)
    with torch.no_grad():
        for p in m.parameters():
            if p.requires_grad and p.abs().sum() == 0: p.normal_(0, 0.02)
    x = torch.randn(B, 4, 32, 32, device=dev); t = torch.rand(B, device=dev) * 999; y = torch.randint(0, 1000, (B,), device=dev)
    g = torch.randn(B, 4, 32, 32, device=dev).bfloat16()
    def step():
        out, _ = m(x, t, y); out.backward(g)
    for _ in range(3): step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); n = 5
    for _ in range(n): step()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / n
    s.record()
    with torch.no_grad():
        for _ in range(n): m(x, t, y)
    e.record(); torch.cuda.synchronize()
    msf = s.elapsed_time(e) / n
    print(f"time {name} B={B}: fwd+bwd {ms:.2f} ms ({B / ms * 1e3:.0f} img/s), fwd only {msf:.2f} ms, ws {m._ws.numel() / 2**30:.1f} GiB", flush=True)
    del m; torch.cuda.empty_cache()
time_model("DiT-S/2", 384, 6, 12, 64)
time_model("DiT-S/2", 384, 6, 12, 256)
time_model("DiT-XL/2", 1152, 16, 28, 64)
print("FAILS", fails)
sys.exit(1 if fails else 0)
