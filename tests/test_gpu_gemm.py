"""GPU tier: the tcgen05 GEMM through the C ABI vs an fp32 torch reference on the same bf16-rounded operands:
every operand layout (forward / dgrad / wgrad), every tile width, ragged shapes, every fused epilogue, split-K."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import relerr, run_gemm
from gpu_util import run_gemm as _rg
from vaw_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("shape", [(256, 384, 128), (512, 1152, 1152), (200, 136, 72), (64, 2304, 384), (16, 1152, 4096),
                                   (1152, 16, 2048)])
@pytest.mark.parametrize("tile_n", [128, 192, 256])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("cta_group", [1, 2])
def test_layouts_and_tiles(shape, tile_n, a_mn, b_mn, cta_group):
    M, N, K = shape
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("MN-major operands need the MN extent to be a multiple of 8")
    torch.manual_seed(0)
    A = torch.randn(M, K, device=DEV).bfloat16(); B = torch.randn(N, K, device=DEV).bfloat16()
    ref = A.float() @ B.float().t()
    out = torch.full((M, N), float("nan"), device=DEV)
    run_gemm(A.t().contiguous() if a_mn else A, B.t().contiguous() if b_mn else B, a_mn, b_mn, M, N, K, L.EPI_F32,
             out=out, tile_n=tile_n, cta_group=cta_group)
    assert relerr(out, ref) < 1e-5


def test_full_size_block_gemms_linearity():
    """At the benchmark's full size (M = 64*256 tokens, DiT-XL widths): GEMM(A1 + A2) == GEMM(A1) + GEMM(A2) with the
    fp32 accumulate epilogue, and the result equals torch within bf16-input rounding."""
    torch.manual_seed(1)
    M, N, K = 16384, 4608, 1152
    A1 = torch.randn(M, K, device=DEV).bfloat16(); W = (torch.randn(N, K, device=DEV) * 0.03).bfloat16()
    out = torch.zeros(M, N, device=DEV)
    run_gemm(A1, W, 0, 0, M, N, K, L.EPI_F32, out=out)
    run_gemm(A1, W, 0, 0, M, N, K, L.EPI_F32, out=out, accumulate=1)
    ref = A1.float() @ W.float().t()
    assert relerr(out, 2 * ref) < 1e-5


@pytest.mark.parametrize("cta_group", [1, 2])
def test_epilogues(cta_group):
    import functools
    global run_gemm
    from gpu_util import run_gemm as _rg
    run_gemm = functools.partial(_rg, cta_group=cta_group)
    try:
        _epilogues_body()
    finally:
        run_gemm = _rg


def _epilogues_body():
    torch.manual_seed(2)
    M, N, K, rps = 1024, 1152, 384, 256
    A = torch.randn(M, K, device=DEV).bfloat16(); B = (torch.randn(N, K, device=DEV) * 0.05).bfloat16()
    bias = torch.randn(N, device=DEV)
    acc = A.float() @ B.float().t() + bias
    pre = acc.bfloat16().float()
    o = torch.empty(M, N, device=DEV, dtype=torch.bfloat16); o2 = torch.empty_like(o)
    run_gemm(A, B, 0, 0, M, N, K, L.EPI_BF16, out=o, bias=bias)
    assert torch.equal(o, acc.bfloat16()) or relerr(o, acc) < 3e-3
    for epi, fn in ((L.EPI_GELU_TANH, lambda x: F.gelu(x, approximate="tanh")), (L.EPI_GELU_ERF, F.gelu), (L.EPI_SILU, F.silu)):
        run_gemm(A, B, 0, 0, M, N, K, epi, out=o, out2=o2, bias=bias)
        assert relerr(o, acc) < 3e-3 and relerr(o2, fn(pre)) < 4e-3
    resid = torch.randn(M, N, device=DEV); gate = torch.randn(M // rps, N, device=DEV); xo = torch.empty(M, N, device=DEV)
    run_gemm(A, B, 0, 0, M, N, K, L.EPI_GATE_RES, out=o, out2=xo, bias=bias, resid=resid, gate=gate, rows_per_sample=rps)
    assert relerr(xo, resid + gate.repeat_interleave(rps, 0) * pre) < 2e-3
    run_gemm(A, B, 0, 0, M, N, K, L.EPI_RES, out2=xo, bias=bias, resid=resid)
    assert relerr(xo, resid + pre) < 2e-3
    pos = torch.randn(rps, N, device=DEV)
    run_gemm(A, B, 0, 0, M, N, K, L.EPI_RES, out2=xo, bias=bias, resid=pos, resid_mod=rps)
    assert relerr(xo, pos.repeat(M // rps, 1) + pre) < 2e-3
    aux = torch.randn(M, N, device=DEV).bfloat16()
    for epi, fn in ((L.EPI_DGELU_TANH, lambda x: F.gelu(x, approximate="tanh")), (L.EPI_DGELU_ERF, F.gelu), (L.EPI_DSILU, F.silu)):
        h = aux.float().requires_grad_(True)
        fn(h).sum().backward()
        run_gemm(A, B, 0, 0, M, N, K, epi, out=o, aux=aux)
        assert relerr(o, (acc - bias) * h.grad) < 4e-3


def test_tail_split_wgrad_shapes():
    """k_splits = -1: whole tiles for the full waves, the partial last wave split along K (wgrad of a DiT-XL block)."""
    torch.manual_seed(4)
    for (Mo, No, Kt) in [(3456, 1152, 4096), (1152, 1152, 4096), (4608, 1152, 2048), (1152, 4608, 2048)]:
        A = torch.randn(Kt, Mo, device=DEV).bfloat16(); B = (torch.randn(Kt, No, device=DEV) * 0.05).bfloat16()
        ref = A.float().t() @ B.float()
        ws = torch.empty(148 * 128 * 256, device=DEV)
        out = torch.full((Mo, No), 0.5, device=DEV)
        run_gemm(A, B, 1, 1, Mo, No, Kt, L.EPI_F32, out=out, accumulate=1, k_splits=-1, split_ws=ws)
        assert relerr(out, ref + 0.5) < 1e-5
        for cg, bn in ((1, 192), (2, 256), (2, 192), (2, 128)):
            o = torch.empty(Mo, No, device=DEV)
            run_gemm(A, B, 1, 1, Mo, No, Kt, L.EPI_F32, out=o, k_splits=-1, split_ws=ws, cta_group=cg, tile_n=bn)
            assert relerr(o, ref) < 1e-5, (cg, bn)
        out2 = torch.empty(Mo, No, device=DEV)
        run_gemm(A, B, 1, 1, Mo, No, Kt, L.EPI_F32, out=out2, k_splits=-1, split_ws=ws)
        out3 = torch.empty(Mo, No, device=DEV)
        run_gemm(A, B, 1, 1, Mo, No, Kt, L.EPI_F32, out=out3, k_splits=-1, split_ws=ws)
        assert torch.equal(out2, out3) and relerr(out2, ref) < 1e-5


@pytest.mark.parametrize("Mo,D,Kt", [(4608, 1152, 4096), (3456, 1152, 2048), (1536, 384, 1024), (3072, 768, 2304), (136, 72, 256)])
def test_weight_and_bias_gradient_in_one_gemm(Mo, D, Kt):
    """Row-sum form of VAW_EPI_F32 (out2 != NULL): B = [X | 1 0 ... 0] is read as [K, D + 32]; out[Mo, D] = A^T X lands in
    a dense [Mo, D] buffer (ldo = D), column D of the product - the column sums of A, i.e. a Linear's bias gradient - goes
    to out2[Mo].  Every tile shape, with and without tail split-K, with accumulate, and nothing written out of bounds."""
    torch.manual_seed(Mo + D)
    A = torch.randn(Kt, Mo, device=DEV).bfloat16()
    X = torch.empty(Kt, D + 32, device=DEV, dtype=torch.bfloat16)
    X[:, :D] = (torch.randn(Kt, D, device=DEV) * 0.05).bfloat16()
    X[:, D:] = 0
    X[:, D] = 1
    ref_w = A.float().t() @ X[:, :D].float()
    ref_b = A.float().sum(0)
    ws = torch.empty(148 * 128 * 256, device=DEV)
    pad = 4096

    def guarded(n):
        buf = torch.full((n + 2 * pad,), 7.0, device=DEV)
        return buf, buf[pad:pad + n]

    for kw in (dict(), dict(k_splits=-1, split_ws=ws), dict(cta_group=1, tile_n=192, k_splits=-1, split_ws=ws),
               dict(cta_group=2, tile_n=256, k_splits=-1, split_ws=ws), dict(cta_group=1, tile_n=128),
               dict(cta_group=2, tile_n=128, k_splits=-1, split_ws=ws), dict(cta_group=1, tile_n=256, k_splits=3, split_ws=ws)):
        if kw.get("cta_group") == 2 and Mo < 256:
            continue
        bw, w = guarded(Mo * D)
        bb, b = guarded(Mo)
        run_gemm(A, X, 1, 1, Mo, D + 32, Kt, L.EPI_F32, out=w, out2=b, **kw)
        assert relerr(w.view(Mo, D), ref_w) < 1e-5, kw
        assert relerr(b, ref_b) < 1e-5, kw
        run_gemm(A, X, 1, 1, Mo, D + 32, Kt, L.EPI_F32, out=w, out2=b, accumulate=1, **kw)
        assert relerr(w.view(Mo, D), 2 * ref_w) < 1e-5 and relerr(b, 2 * ref_b) < 1e-5, kw
        for buf in (bw, bb):
            assert bool((buf[:pad] == 7.0).all()) and bool((buf[-pad:] == 7.0).all()), kw
    with pytest.raises(L.VawError):   # the row-sum form takes no bias
        run_gemm(A, X, 1, 1, Mo, D + 32, Kt, L.EPI_F32, out=w, out2=b, bias=torch.zeros(D + 32, device=DEV))


@pytest.mark.parametrize("splits", [2, 7, 24])
def test_split_k_deterministic_and_correct(splits):
    torch.manual_seed(3)
    M, N, K = 64, 1152, 193536 // 8
    A = torch.randn(M, K, device=DEV).bfloat16(); B = (torch.randn(K, N, device=DEV) * 0.02).bfloat16()
    ref = A.float() @ B.float()
    ws = torch.empty(splits * 128 * 6 * 192, device=DEV)
    bias = torch.randn(N, device=DEV)
    out = torch.ones(M, N, device=DEV)
    run_gemm(A, B, 0, 1, M, N, K, L.EPI_F32, out=out, bias=bias, accumulate=1, k_splits=splits, split_ws=ws)
    assert relerr(out, ref + bias + 1) < 3e-5  # K = 24192 fp32 accumulation order
    out2 = torch.ones(M, N, device=DEV)
    run_gemm(A, B, 0, 1, M, N, K, L.EPI_F32, out=out2, bias=bias, accumulate=1, k_splits=splits, split_ws=ws)
    assert torch.equal(out, out2)


def test_rejects_bad_arguments():
    A = torch.zeros(128, 64, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(L.VawError):
        run_gemm(A, A, 0, 0, 128, 100, 64, L.EPI_F32, out=torch.zeros(128, 100, device=DEV))  # N % 8
    with pytest.raises(L.VawError):
        run_gemm(A, A, 0, 0, 128, 128, 64, L.EPI_GATE_RES, out=A)  # missing out2/resid/gate


@pytest.mark.parametrize("tile_n,cta_group", [(128, 1), (192, 1), (256, 1), (128, 2), (192, 2), (256, 2)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 200, 136), (512, 1152, 384)])
def test_tma_store_epilogues_every_tile_config(tile_n, cta_group, M, N, K):
    """The bf16 epilogues leave through swizzled shared-memory boxes + TMA stores; the staging area's alignment (and with
    it the swizzle phase) depends on the tile configuration, so every configuration is checked, ragged shapes included."""
    torch.manual_seed(M + N)
    A = torch.randn(M, K, device=DEV).bfloat16(); B = (torch.randn(N, K, device=DEV) * 0.1).bfloat16()
    bias = torch.randn(N, device=DEV)
    acc = A.float() @ B.float().t() + bias
    pre = acc.bfloat16().float()
    o = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16); o2 = torch.full_like(o, float("nan"))
    _rg(A, B, 0, 0, M, N, K, L.EPI_BF16, out=o, bias=bias, tile_n=tile_n, cta_group=cta_group)
    assert relerr(o, acc) < 3e-3
    for epi, fn in ((L.EPI_GELU_TANH, lambda x: F.gelu(x, approximate="tanh")), (L.EPI_GELU_ERF, F.gelu), (L.EPI_SILU, F.silu)):
        o.fill_(float("nan")); o2.fill_(float("nan"))
        _rg(A, B, 0, 0, M, N, K, epi, out=o, out2=o2, bias=bias, tile_n=tile_n, cta_group=cta_group)
        assert relerr(o, acc) < 3e-3 and relerr(o2, fn(pre)) < 4e-3, epi
    # forward-only callers drop the saved pre-activation: out = NULL, out2 unchanged
    o2b = torch.full_like(o2, float("nan"))
    _rg(A, B, 0, 0, M, N, K, L.EPI_SILU, out=None, out2=o2b, bias=bias, tile_n=tile_n, cta_group=cta_group)
    assert torch.equal(o2b, o2)
    # residual epilogues: resid in / out2 + y out through TMA in 16-column halves (ragged sample boundaries included)
    rps = {128: 64, 300: 100, 512: 256}[M]
    resid = torch.randn(M, N, device=DEV); gate = torch.randn(M // rps, N, device=DEV)
    xo = torch.full((M, N), float("nan"), device=DEV)
    o.fill_(float("nan"))
    _rg(A, B, 0, 0, M, N, K, L.EPI_GATE_RES, out=o, out2=xo, bias=bias, resid=resid, gate=gate, rows_per_sample=rps,
        tile_n=tile_n, cta_group=cta_group)
    assert relerr(o, acc) < 3e-3 and relerr(xo, resid + gate.repeat_interleave(rps, 0) * pre) < 2e-3
    xo_b = torch.full_like(xo, float("nan"))       # without the branch output y (forward-only)
    _rg(A, B, 0, 0, M, N, K, L.EPI_GATE_RES, out=None, out2=xo_b, bias=bias, resid=resid, gate=gate, rows_per_sample=rps,
        tile_n=tile_n, cta_group=cta_group)
    assert torch.equal(xo_b, xo)
    xo.fill_(float("nan"))
    _rg(A, B, 0, 0, M, N, K, L.EPI_RES, out2=xo, bias=bias, resid=resid, tile_n=tile_n, cta_group=cta_group)
    assert relerr(xo, resid + pre) < 2e-3
    xo2 = resid.clone()    # in place: out2 aliases resid (the U-ViT residual stream is updated in place)
    _rg(A, B, 0, 0, M, N, K, L.EPI_RES, out2=xo2, bias=bias, resid=xo2, tile_n=tile_n, cta_group=cta_group)
    assert relerr(xo2, resid + pre) < 2e-3
    # d-activation epilogues: the saved pre-activation tile arrives by TMA as well
    aux = torch.randn(M, N, device=DEV).bfloat16()
    for epi, fn in ((L.EPI_DGELU_TANH, lambda x: F.gelu(x, approximate="tanh")), (L.EPI_DGELU_ERF, F.gelu), (L.EPI_DSILU, F.silu)):
        h = aux.float().requires_grad_(True)
        fn(h).sum().backward()
        o.fill_(float("nan"))
        _rg(A, B, 0, 0, M, N, K, epi, out=o, aux=aux, tile_n=tile_n, cta_group=cta_group)
        assert relerr(o, (acc - bias) * h.grad) < 4e-3, epi
    # REPA alignment loss accumulated in the epilogue: zs out, one partial per 32 x 32 block of (zs - feat)^2
    if N % 4 == 0:
        feat = torch.randn(M, N, device=DEV).bfloat16()
        nrb, ncb = (M + 31) // 32, (N + 31) // 32
        part = torch.full((nrb * ncb,), float("nan"), device=DEV)
        o.fill_(float("nan"))
        _rg(A, B, 0, 0, M, N, K, L.EPI_ALIGN_MSE, out=o, out2=part, bias=bias, aux=feat, tile_n=tile_n, cta_group=cta_group)
        assert relerr(o, acc) < 3e-3
        d2 = (o.float() - feat.float()) ** 2                      # the loss sees the bf16 zs that was written
        pad = torch.zeros(nrb * 32, ncb * 32, device=DEV)
        pad[:M, :N] = d2
        want = pad.view(nrb, 32, ncb, 32).sum(dim=(1, 3)).reshape(-1)
        torch.testing.assert_close(part, want, rtol=1e-5, atol=1e-6)
        part2 = torch.empty_like(part)
        _rg(A, B, 0, 0, M, N, K, L.EPI_ALIGN_MSE, out=o, out2=part2, bias=bias, aux=feat, tile_n=tile_n, cta_group=cta_group)
        assert torch.equal(part, part2)                           # fixed reduction tree: bit-reproducible


@pytest.mark.parametrize("epi", ["bf16", "gelu", "dgelu", "gate_res", "f32"])
def test_outputs_stay_inside_their_buffers(epi):
    """Guard bands around every output: ragged shapes (M, N not multiples of the tile), TMA stores and direct stores
    must not touch a byte outside [M, N]."""
    torch.manual_seed(7)
    M, N, K, rps = 300, 200, 136, 100
    A = torch.randn(M, K, device=DEV).bfloat16(); B = (torch.randn(N, K, device=DEV) * 0.1).bfloat16()
    bias = torch.randn(N, device=DEV)
    pad = 4096

    def guarded(dtype):
        buf = torch.full((M * N + 2 * pad,), 7.0, device=DEV, dtype=dtype)
        return buf, buf[pad:pad + M * N].view(M, N)

    def check(buf):
        assert bool((buf[:pad] == 7.0).all()) and bool((buf[-pad:] == 7.0).all())

    if epi == "bf16":
        b1, o = guarded(torch.bfloat16)
        _rg(A, B, 0, 0, M, N, K, L.EPI_BF16, out=o, bias=bias)
        check(b1)
    elif epi == "gelu":
        b1, o = guarded(torch.bfloat16); b2, o2 = guarded(torch.bfloat16)
        _rg(A, B, 0, 0, M, N, K, L.EPI_GELU_TANH, out=o, out2=o2, bias=bias)
        check(b1); check(b2)
    elif epi == "dgelu":
        b1, o = guarded(torch.bfloat16)
        aux = torch.randn(M, N, device=DEV).bfloat16()
        _rg(A, B, 0, 0, M, N, K, L.EPI_DGELU_TANH, out=o, aux=aux)
        check(b1)
    elif epi == "gate_res":
        b1, o = guarded(torch.bfloat16); b2, xo = guarded(torch.float32)
        resid = torch.randn(M, N, device=DEV); gate = torch.randn(M // rps, N, device=DEV)
        _rg(A, B, 0, 0, M, N, K, L.EPI_GATE_RES, out=o, out2=xo, bias=bias, resid=resid, gate=gate, rows_per_sample=rps)
        check(b1); check(b2)
        assert relerr(xo, resid + gate.repeat_interleave(rps, 0) * (A.float() @ B.float().t() + bias).bfloat16().float()) < 2e-3
    else:
        b1, o = guarded(torch.float32)
        _rg(A, B, 0, 0, M, N, K, L.EPI_F32, out=o, bias=bias)
        check(b1)
        assert relerr(o, A.float() @ B.float().t() + bias) < 1e-5


@pytest.mark.parametrize("M,N,K", [(6912, 1152, 64), (768, 384, 8), (300, 130, 5), (256, 256, 200), (128, 128, 16)])
@pytest.mark.parametrize("accumulate", [0, 1])
def test_wgrad_smallk_vs_torch(M, N, K, accumulate):
    """vaw_wgrad_smallk: dW = A^T B over K = batch rows (the adaLN / timestep-embedder weight gradients), ragged M / N,
    K not a multiple of 16, overwrite and accumulate; bit-reproducible."""
    import ctypes as C
    L.register("vaw_wgrad_smallk", [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p])
    torch.manual_seed(M + K)
    lda, ldb = (M + 7) // 8 * 8 + 8, (N + 7) // 8 * 8
    A = torch.randn(K, lda, device=DEV).bfloat16(); B = torch.randn(K, ldb, device=DEV).bfloat16()
    ldo = N + (N & 1)
    base = torch.randn(M, ldo, device=DEV)
    out = base.clone()
    L.call("vaw_wgrad_smallk", A.data_ptr(), lda, B.data_ptr(), ldb, out.data_ptr(), ldo, M, N, K, accumulate, L.stream_ptr())
    want = A[:, :M].float().t() @ B[:, :N].float() + (base[:, :N] if accumulate else 0)
    assert relerr(out[:, :N], want) < 1e-5
    if ldo > N:
        assert torch.equal(out[:, N:], base[:, N:])      # the padding column is not touched
    out2 = base.clone()
    L.call("vaw_wgrad_smallk", A.data_ptr(), lda, B.data_ptr(), ldb, out2.data_ptr(), ldo, M, N, K, accumulate, L.stream_ptr())
    assert torch.equal(out, out2)
