"""GPU tier: the HBM-bound row kernels of a block (LayerNorm forward/backward with adaLN modulate or affine, the
gate*branch backward, the partial-sum finishers) through the C ABI, against fp32 torch on the same inputs.
Covers the shared-memory staged LayerNorm backward (D % 8 == 0) and its two-pass fallback (D % 8 == 4)."""
import ctypes as C

import pytest
import torch

from gpu_util import relerr
from vaw_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda"
P = C.c_void_p
L.register("vaw_ln_fwd", [P] * 3 + [C.c_longlong, C.c_int] + [P] * 5 + [C.c_int, C.c_int, C.c_float, P])
L.register("vaw_ln_bwd", [P] * 5 + [C.c_longlong] + [P] * 2 + [C.c_int, P] + [C.c_int] * 5 + [P])
L.register("vaw_gate_bwd", [P] * 3 + [C.c_longlong] + [P] * 2 + [C.c_int] * 5 + [P])
L.register("vaw_ln_bwd_gate", [P] * 5 + [C.c_longlong] + [P] * 2 + [C.c_int, P, P, P, C.c_longlong, P, P] + [C.c_int] * 5 + [P])
L.register("vaw_finish_group", [P, C.c_int, C.c_int, C.c_int, C.c_int, P, C.c_longlong, C.c_int, P])
L.register("vaw_finish_all", [P, C.c_int, C.c_int, C.c_int, C.c_int, P, C.c_longlong, P, C.c_int, P])


L.register("vaw_ln_fwd_res", [P] * 3 + [C.c_longlong] + [P] * 3 + [C.c_longlong, C.c_int] + [P] * 3 + [C.c_int, C.c_int, C.c_float, P])


@pytest.mark.parametrize("B,T,D", [(4, 256, 1152), (3, 258, 768), (5, 64, 384), (2, 100, 132), (2, 64, 2048), (1, 1, 128)])
def test_layernorm_fwd_with_the_previous_branch_folded_in(B, T, D):
    """vaw_ln_fwd_res: x_out = x + gate * branch (fp32, one fma per element: bit-exact against torch.addcmul on the same
    bf16 branch), then LayerNorm + modulate of x_out - the same numbers vaw_ln_fwd gives on x_out."""
    torch.manual_seed(B * 7 + D)
    M = B * T
    x = torch.randn(M, D, device=DEV) * 1.5 + 0.3
    branch = torch.randn(M, D, device=DEV).bfloat16()
    mod = torch.randn(B, 6 * D, device=DEV) * 0.3                    # [.., gate at 2D, shift at 3D, scale at 4D, ..]
    gate, shift, scale = mod[:, 2 * D:3 * D], mod[:, 3 * D:4 * D], mod[:, 4 * D:5 * D]
    pad = 1024
    xo_buf = torch.full((M * D + 2 * pad,), 7.0, device=DEV)
    x_out = xo_buf[pad:pad + M * D].view(M, D)
    y = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    L.call("vaw_ln_fwd_res", x.data_ptr(), branch.data_ptr(), gate.data_ptr(), 6 * D, x_out.data_ptr(), shift.data_ptr(),
           scale.data_ptr(), 6 * D, T, y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), M, D, 1e-6, L.stream_ptr())
    g = gate.repeat_interleave(T, 0)
    want = torch.addcmul(x.double(), g.double(), branch.double()).float()     # correctly rounded fma
    assert torch.equal(x_out, want)
    assert bool((xo_buf[:pad] == 7.0).all()) and bool((xo_buf[-pad:] == 7.0).all())
    y2 = torch.empty_like(y); mean2, rstd2 = torch.empty_like(mean), torch.empty_like(rstd)
    L.call("vaw_ln_fwd", x_out.data_ptr(), shift.data_ptr(), scale.data_ptr(), 6 * D, T, None, None, y2.data_ptr(),
           mean2.data_ptr(), rstd2.data_ptr(), M, D, 1e-6, L.stream_ptr())
    assert torch.equal(y, y2) and torch.equal(mean, mean2) and torch.equal(rstd, rstd2)
    y_ref, _ = _ln_ref(want, (1 + scale).repeat_interleave(T, 0), shift.repeat_interleave(T, 0))
    assert relerr(y, y_ref) < 4e-3


L.register("vaw_ln_fwd_ex", [P] * 3 + [C.c_longlong] + [P] * 3 + [C.c_longlong, C.c_int] + [P] * 3 +
           [C.c_longlong, C.c_int] + [P] * 2 + [C.c_int, C.c_int, C.c_float, P])


@pytest.mark.parametrize("B,T,D,fold", [(4, 256, 1152, True), (3, 64, 384, False), (2, 258, 768, True), (2, 33, 136, False)])
def test_layernorm_fwd_with_row_stride_and_ones_block(B, T, D, fold):
    """vaw_ln_fwd_ex: the same values as vaw_ln_fwd / vaw_ln_fwd_res in a [M, D + 32] buffer whose 32 trailing columns are
    [1, 0, ..., 0] in every row (the B operand of the weight + bias gradient GEMM)."""
    torch.manual_seed(B + D)
    M = B * T
    x = torch.randn(M, D, device=DEV) * 1.5 + 0.3
    branch = torch.randn(M, D, device=DEV).bfloat16()
    mod = torch.randn(B, 6 * D, device=DEV) * 0.3
    gate, shift, scale = mod[:, 2 * D:3 * D], mod[:, 3 * D:4 * D], mod[:, 4 * D:5 * D]
    y = torch.full((M + 2, D + 32), 5.0, device=DEV, dtype=torch.bfloat16)     # one guard row on each side
    yv = y[1:M + 1]
    x_out = torch.empty(M, D, device=DEV)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    L.call("vaw_ln_fwd_ex", x.data_ptr(), branch.data_ptr() if fold else None, gate.data_ptr(), 6 * D, x_out.data_ptr(),
           shift.data_ptr(), scale.data_ptr(), 6 * D, T, None, None, yv.data_ptr(), D + 32, 1, mean.data_ptr(), rstd.data_ptr(),
           M, D, 1e-6, L.stream_ptr())
    y2 = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    m2, r2 = torch.empty_like(mean), torch.empty_like(rstd)
    if fold:
        xo2 = torch.empty_like(x_out)
        L.call("vaw_ln_fwd_res", x.data_ptr(), branch.data_ptr(), gate.data_ptr(), 6 * D, xo2.data_ptr(), shift.data_ptr(),
               scale.data_ptr(), 6 * D, T, y2.data_ptr(), m2.data_ptr(), r2.data_ptr(), M, D, 1e-6, L.stream_ptr())
        assert torch.equal(x_out, xo2)
    else:
        L.call("vaw_ln_fwd", x.data_ptr(), shift.data_ptr(), scale.data_ptr(), 6 * D, T, None, None, y2.data_ptr(),
               m2.data_ptr(), r2.data_ptr(), M, D, 1e-6, L.stream_ptr())
    assert torch.equal(yv[:, :D], y2) and torch.equal(mean, m2) and torch.equal(rstd, r2)
    assert bool((yv[:, D] == 1).all()) and bool((yv[:, D + 1:] == 0).all())
    assert bool((y[0] == 5).all()) and bool((y[M + 1] == 5).all())


def test_layernorm_fwd_res_rejects_bad_arguments():
    x = torch.zeros(4, 128, device=DEV)
    b = torch.zeros(4, 128, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(L.VawError):
        L.call("vaw_ln_fwd_res", x.data_ptr(), None, x.data_ptr(), 128, x.data_ptr(), x.data_ptr(), x.data_ptr(), 128, 4,
               b.data_ptr(), x.data_ptr(), x.data_ptr(), 4, 128, 1e-6, L.stream_ptr())


def _ln_ref(x, A, Bv, eps=1e-6):
    xh = torch.nn.functional.layer_norm(x, x.shape[-1:], eps=eps)
    return xh * A + Bv, xh


@pytest.mark.parametrize("B,T,D,chunks,mode", [
    (4, 256, 1152, 8, "mod"), (3, 258, 768, 9, "mod"), (5, 64, 384, 2, "mod"), (2, 100, 132, 4, "mod"),
    (1, 37, 1152, 3, "affine"), (6, 17, 512, 1, "affine"), (2, 257, 1536, 5, "plain"), (2, 64, 2048, 2, "mod"),
    (3, 300, 128, 1, "mod"),
])
@pytest.mark.parametrize("add_into", [0, 1])
def test_layernorm_fwd_bwd(B, T, D, chunks, mode, add_into):
    torch.manual_seed(B * 1000 + D)
    M = B * T
    x = torch.randn(M, D, device=DEV) * 1.5 + 0.3
    mod = torch.randn(B, 3 * D, device=DEV) * 0.3
    shift, scale = mod[:, :D], mod[:, D:2 * D]
    w, b = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    y = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    st = L.stream_ptr()
    if mode == "mod":
        L.call("vaw_ln_fwd", x.data_ptr(), shift.data_ptr(), scale.data_ptr(), 3 * D, T, None, None, y.data_ptr(),
               mean.data_ptr(), rstd.data_ptr(), M, D, 1e-6, st)
        A = (1 + scale).repeat_interleave(T, 0); Bv = shift.repeat_interleave(T, 0)
        groups, rpg = B, T
    elif mode == "affine":
        L.call("vaw_ln_fwd", x.data_ptr(), None, None, 0, 1, w.data_ptr(), b.data_ptr(), y.data_ptr(),
               mean.data_ptr(), rstd.data_ptr(), M, D, 1e-6, st)
        A, Bv = w.expand(M, D), b.expand(M, D)
        groups, rpg = B, T
    else:
        L.call("vaw_ln_fwd", x.data_ptr(), None, None, 0, 1, None, None, y.data_ptr(), mean.data_ptr(),
               rstd.data_ptr(), M, D, 1e-6, st)
        A, Bv = torch.ones(M, D, device=DEV), torch.zeros(M, D, device=DEV)
        groups, rpg = B, T
    y_ref, xh = _ln_ref(x, A, Bv)
    assert relerr(y, y_ref) < 4e-3            # bf16 output rounding
    torch.testing.assert_close(mean, x.mean(1), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(rstd, (x.var(1, unbiased=False) + 1e-6).rsqrt(), rtol=1e-5, atol=1e-5)

    dy = torch.randn(M, D, device=DEV).bfloat16()
    dx0 = torch.randn(M, D, device=DEV)
    dx = dx0.clone()
    rpc = -(-rpg // chunks)
    if rpc > 64 and D % 8 != 0:
        pytest.skip("fallback kernel handles at most 64 rows per chunk")
    part = torch.full((groups, chunks, 2, D), float("nan"), device=DEV)
    L.call("vaw_ln_bwd", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
           scale.data_ptr() if mode == "mod" else None, 3 * D if mode == "mod" else 0,
           w.data_ptr() if mode == "affine" else None, dx.data_ptr(), add_into, part.data_ptr(), rpg, groups, chunks,
           M, D, st)
    xr = x.clone().requires_grad_(True)
    yr, _ = _ln_ref(xr, A, Bv)
    yr.backward(dy.float())
    want = xr.grad + (dx0 if add_into else 0)
    assert relerr(dx, want) < 2e-5
    dyf = dy.float()
    sum_dy = dyf.view(groups, rpg, D).sum(1)
    sum_dyxh = (dyf * xh).view(groups, rpg, D).sum(1)
    assert not torch.isnan(part).any()
    assert relerr(part[:, :, 0].sum(1), sum_dy) < 1e-5
    assert relerr(part[:, :, 1].sum(1), sum_dyxh) < 1e-5
    # finishers: per-group vector and weighted all-group vector
    out_g = torch.zeros(groups, D, device=DEV)
    L.call("vaw_finish_group", part.data_ptr(), 1, groups, chunks, D, out_g.data_ptr(), D, 0, st)
    assert relerr(out_g, sum_dyxh) < 1e-5
    out_a = torch.ones(D, device=DEV)
    L.call("vaw_finish_all", part.data_ptr(), 0, groups, chunks, D, None, 0, out_a.data_ptr(), 1, st)
    assert relerr(out_a, 1 + sum_dy.sum(0)) < 1e-5
    # determinism: a second run gives the same bits
    dx2 = dx0.clone(); part2 = torch.empty_like(part)
    L.call("vaw_ln_bwd", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
           scale.data_ptr() if mode == "mod" else None, 3 * D if mode == "mod" else 0,
           w.data_ptr() if mode == "affine" else None, dx2.data_ptr(), add_into, part2.data_ptr(), rpg, groups, chunks,
           M, D, st)
    assert torch.equal(dx, dx2) and torch.equal(part, part2)


@pytest.mark.parametrize("B,T,D,chunks,gated", [(4, 256, 1152, 8, True), (3, 258, 768, 9, False), (2, 50, 132, 2, True)])
def test_gate_bwd(B, T, D, chunks, gated):
    torch.manual_seed(D)
    M = B * T
    dx = torch.randn(M, D, device=DEV)
    y = torch.randn(M, D, device=DEV).bfloat16()
    gate = torch.randn(B, 2 * D, device=DEV)[:, D:]
    dy = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    part = torch.empty(B, chunks, 2, D, device=DEV)
    L.call("vaw_gate_bwd", dx.data_ptr(), y.data_ptr() if gated else None, gate.data_ptr() if gated else None, 2 * D,
           dy.data_ptr(), part.data_ptr(), T, B, chunks, M, D, L.stream_ptr())
    g = gate.repeat_interleave(T, 0) if gated else 1.0
    assert relerr(dy, dx * g) < 4e-3
    assert relerr(part[:, :, 0].sum(1), dx.view(B, T, D).sum(1)) < 1e-5
    if gated:
        assert relerr(part[:, :, 1].sum(1), (dx * y.float()).view(B, T, D).sum(1)) < 1e-5


@pytest.mark.parametrize("M,N,rows", [(16384, 3456, 1496), (1000, 4608, 256), (777, 130, 64), (64, 72, 8), (4096, 2050, 512)])
@pytest.mark.parametrize("acc", [0, 1])
def test_colsum_bf16(M, N, rows, acc):
    """Bias-gradient column sums (16-byte-load kernel for N % 8 == 0, 4-byte fallback otherwise), deterministic."""
    L.register("vaw_colsum_bf16", [P, C.c_longlong, C.c_int, C.c_int, P, C.c_int, P, C.c_int, P])
    torch.manual_seed(N)
    a = torch.randn(M, N, device=DEV).bfloat16()
    chunks = -(-M // rows)
    part = torch.empty(chunks * N + 1024, device=DEV)
    out = torch.ones(N, device=DEV)
    L.call("vaw_colsum_bf16", a.data_ptr(), N, M, N, part.data_ptr(), rows, out.data_ptr(), acc, L.stream_ptr())
    want = a.float().sum(0) + (1.0 if acc else 0.0)
    assert relerr(out, want) < 1e-5
    out2 = torch.ones(N, device=DEV)
    L.call("vaw_colsum_bf16", a.data_ptr(), N, M, N, part.data_ptr(), rows, out2.data_ptr(), acc, L.stream_ptr())
    assert torch.equal(out, out2)


@pytest.mark.parametrize("B,T,D,chunks", [(4, 256, 1152, 9), (3, 64, 384, 2), (2, 100, 132, 4), (2, 257, 1536, 5)])
@pytest.mark.parametrize("gated", [True, False])
def test_ln_bwd_fused_with_next_gate_bwd(B, T, D, chunks, gated):
    """vaw_ln_bwd_gate == vaw_ln_bwd followed by vaw_gate_bwd on the updated residual gradient (bit-identical dx and
    LayerNorm partials; dy / gate partials to rounding), for the fused kernel (D <= 1280, D % 8 == 0) and its fallbacks."""
    torch.manual_seed(D + chunks)
    M = B * T
    x = torch.randn(M, D, device=DEV); mod = torch.randn(B, 3 * D, device=DEV) * 0.3
    scale, gate = mod[:, D:2 * D], mod[:, 2 * D:]
    y = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    st = L.stream_ptr()
    L.call("vaw_ln_fwd", x.data_ptr(), mod.data_ptr(), scale.data_ptr(), 3 * D, T, None, None, y.data_ptr(),
           mean.data_ptr(), rstd.data_ptr(), M, D, 1e-6, st)
    dy = torch.randn(M, D, device=DEV).bfloat16()
    y_next = torch.randn(M, D, device=DEV).bfloat16()
    dx0 = torch.randn(M, D, device=DEV)
    # reference: two launches
    dx_a = dx0.clone(); part_a = torch.empty(B, chunks, 2, D, device=DEV); pg_a = torch.empty_like(part_a)
    dyn_a = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    L.call("vaw_ln_bwd", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), scale.data_ptr(), 3 * D, None,
           dx_a.data_ptr(), 1, part_a.data_ptr(), T, B, chunks, M, D, st)
    L.call("vaw_gate_bwd", dx_a.data_ptr(), y_next.data_ptr() if gated else None, gate.data_ptr() if gated else None,
           3 * D, dyn_a.data_ptr(), pg_a.data_ptr(), T, B, chunks, M, D, st)
    # fused
    dx_b = dx0.clone(); part_b = torch.empty_like(part_a); pg_b = torch.full_like(part_a, float("nan"))
    dyn_b = torch.empty_like(dyn_a)
    L.call("vaw_ln_bwd_gate", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), scale.data_ptr(), 3 * D, None,
           dx_b.data_ptr(), 1, part_b.data_ptr(), y_next.data_ptr() if gated else None, gate.data_ptr() if gated else None,
           3 * D, dyn_b.data_ptr(), pg_b.data_ptr(), T, B, chunks, M, D, st)
    assert relerr(dx_b, dx_a) < 1e-6
    assert relerr(part_b.sum(1), part_a.sum(1)) < 1e-5
    assert relerr(dyn_b, dyn_a) < 4e-3
    assert relerr(pg_b[:, :, 0].sum(1), pg_a[:, :, 0].sum(1)) < 1e-5
    if gated:
        assert relerr(pg_b[:, :, 1].sum(1), pg_a[:, :, 1].sum(1)) < 1e-5
