"""GPU tier: K1 (q_sample + target) and K2 (weighted MSE fwd+bwd) through the C ABI vs the oracle and the golden
fixtures of the executed reference.  fp32 integer-indexed arithmetic: bit-exact; reductions: 1e-5 relative."""
import os

import numpy as np
import pytest
import torch

from oracle import diffusion as odiff
from vaw_b200.tools import gaussian_diffusion as gd

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"


@pytest.mark.parametrize("mean", ["EPSILON", "START_X", "VELOCITY", "PREVIOUS_X"])
def test_k1_matches_reference_golden_bit_exact(mean):
    g = np.load(os.path.join(G, "diffusion_golden.npz"))
    d = gd.create_gaussian_diffusion(noise_schedule="linear", mean_type=mean.lower(),
                                     weight_type="constant" if mean == "PREVIOUS_X" else "lambda")
    x0, eps, t = (torch.from_numpy(g[k]).to(DEV) for k in ("x0", "eps", "t"))
    assert np.array_equal(d.q_sample(x0, t, eps).cpu().numpy(), g[f"xt_{mean}"])
    assert np.array_equal(d.compute_target(x0, eps, t).cpu().numpy(), g[f"target_{mean}"])


@pytest.mark.parametrize("shape", [(16, 3, 32, 32), (64, 4, 32, 32), (5, 3, 64, 64), (3, 1, 7, 9), (1, 4, 32, 32)])
@pytest.mark.parametrize("sched", ["linear", "cosine"])
def test_k1_vs_oracle_bit_exact(shape, sched):
    torch.manual_seed(0)
    d = gd.create_gaussian_diffusion(noise_schedule=sched, mean_type="velocity")
    tb = odiff.tables(odiff.named_beta_schedule(sched, 1000))
    x0 = torch.randn(*shape, device=DEV).clamp(-1, 1)
    eps = torch.randn_like(x0)
    t = torch.randint(0, 1000, (shape[0],), device=DEV)
    t[0] = 999
    xt = d.q_sample(x0, t, eps)
    tg = d.compute_target(x0, eps, t)
    assert np.array_equal(xt.cpu().numpy(), odiff.q_sample(tb, x0.cpu().numpy(), t.cpu().numpy(), eps.cpu().numpy()))
    assert np.array_equal(tg.cpu().numpy(), odiff.target(tb, "VELOCITY", x0.cpu().numpy(), t.cpu().numpy(), eps.cpu().numpy()))


def test_k1_empty_batch():
    d = gd.create_gaussian_diffusion()
    x0 = torch.zeros(0, 3, 8, 8, device=DEV)
    assert d.q_sample(x0, torch.zeros(0, dtype=torch.long, device=DEV), x0).shape == (0, 3, 8, 8)


@pytest.mark.parametrize("mean,wt", [("EPSILON", "lambda"), ("START_X", "lambda"), ("VELOCITY", "min_snr_5"),
                                     ("PREVIOUS_X", "constant"), ("EPSILON", "min_snr_5")])
def test_k2_matches_reference_golden(mean, wt):
    g = np.load(os.path.join(G, "diffusion_golden.npz"))
    d = gd.create_gaussian_diffusion(noise_schedule="linear", mean_type=mean.lower(), weight_type=wt)
    x0, eps, t = (torch.from_numpy(g[k]).to(DEV) for k in ("x0", "eps", "t"))
    out = torch.from_numpy(g["model_out"]).to(DEV).requires_grad_(True)
    terms = d.training_losses(lambda x, ts, **k: out, x0, None, t=t, noise=eps)
    terms["loss"].mean().backward()   # what trainer.py:107-108 does
    np.testing.assert_allclose(terms["mse"].detach().cpu().numpy(), g[f"mse_{mean}_{wt}"], rtol=1e-5)
    np.testing.assert_allclose(out.grad.cpu().numpy(), g[f"grad_{mean}_{wt}"], rtol=1e-5, atol=1e-9)
    assert terms["loss"].dtype == torch.float32 and terms["loss"].shape == (6,)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 8e-3), (torch.float16, 2e-3)])
@pytest.mark.parametrize("chw", [(4, 32, 32), (3, 32, 32), (3, 5, 7)])
def test_k2_vs_oracle_with_sample_weights(dtype, tol, chw):
    torch.manual_seed(1)
    N = 32
    d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda")
    tb = odiff.tables(odiff.named_beta_schedule("cosine", 1000))
    x0 = torch.randn(N, *chw, device=DEV)
    eps = torch.randn_like(x0)
    t = torch.randint(0, 1000, (N,), device=DEV)
    out = torch.randn_like(x0).to(dtype).requires_grad_(True)
    w = torch.rand(N, device=DEV) + 0.5
    terms = d.training_losses(lambda x, ts, **k: out, x0, None, t=t, noise=eps)
    (terms["loss"] * w).mean().backward()
    mse, grad = odiff.mse_terms(tb, "EPSILON", "lambda", x0.cpu().numpy(), t.cpu().numpy(), eps.cpu().numpy(),
                                out.detach().float().cpu().numpy())
    np.testing.assert_allclose(terms["mse"].cpu().detach().numpy(), mse, rtol=1e-5)
    ref_grad = grad * (w.cpu().numpy().astype(np.float64) / N).reshape(-1, 1, 1, 1)
    got = out.grad.float().cpu().numpy()
    assert out.grad.dtype == dtype
    assert np.linalg.norm(got - ref_grad) / np.linalg.norm(ref_grad) < tol


def test_k2_linearity_at_scale():
    """Size-independent property at the benchmark's saturating size: the gradient is linear in the upstream weights,
    and mse is invariant under a permutation of the batch."""
    torch.manual_seed(2)
    N = 4096
    d = gd.create_gaussian_diffusion(noise_schedule="cosine")
    x0 = torch.randn(N, 4, 32, 32, device=DEV); eps = torch.randn_like(x0)
    t = torch.randint(0, 1000, (N,), device=DEV)
    out = torch.randn_like(x0).requires_grad_(True)
    mse = d.training_losses(lambda x, ts, **k: out, x0, None, t=t, noise=eps)["mse"]
    g1, = torch.autograd.grad(mse.sum(), out, retain_graph=True)
    g2, = torch.autograd.grad((2.0 * mse).sum(), out)
    assert torch.equal(g2, 2.0 * g1)
    perm = torch.randperm(N, device=DEV)
    mse_p = d.training_losses(lambda x, ts, **k: out[perm], x0[perm], None, t=t[perm], noise=eps[perm])["mse"]
    assert torch.equal(mse_p, mse[perm])


FLOW_CASES = [("START_X", "lambda"), ("EPSILON", "lambda"), ("EPSILON", "min_snr_5"), ("VELOCITY", "lambda"),
              ("VELOCITY", "min_snr_5"), ("VECTOR", "lambda"), ("VECTOR", "constant"), ("SCORE", "constant")]


@pytest.mark.parametrize("path", ("linear", "cosine", "linear_logsnr"))
def test_flow_matching_vs_reference_golden(path):
    """FlowMatching.training_losses / q_sample / compute_target (reference :1273-1340) against the fixture produced by
    executing the reference (tests/golden/make_golden.py::flow_golden).  The linear path has no transcendental in its
    coefficients -> x_t and targets bit-exact; cos / sin / sigmoid differ between the device's and the host's libm in
    the last bit, so those paths carry a 2-ulp tolerance.  The reduction (fp32, different summation order) 1e-5."""
    g = np.load(os.path.join(G, "flow_golden.npz"))
    x0, eps, t, mo = (torch.from_numpy(g[k]).to(DEV) for k in ("x0", "eps", "t", "model_out"))
    exact = path == "linear"
    tol = dict(rtol=0, atol=0) if exact else dict(rtol=1e-6, atol=1e-6)
    for mean, wt in FLOW_CASES:
        fm = gd.FlowMatching(args=gd.default_args(path_type=path, weight_type=wt), model_mean_type=gd.ModelMeanType[mean])
        np.testing.assert_allclose(fm.q_sample(x0, eps, t).cpu().numpy(), g[f"xt::{path}"], **tol)
        np.testing.assert_allclose(fm.compute_target(x0, eps, t).cpu().numpy(), g[f"target::{path}::{mean}"], **tol)
        out = mo.clone().requires_grad_(True)
        seen = {}
        def model(x, ts, **k):
            seen["x"], seen["t"] = x, ts
            return out
        terms = fm.training_losses(model, x0, None, t=t, noise=eps)
        terms["loss"].mean().backward()
        assert torch.equal(seen["t"], t)                      # the model is handed the continuous time itself (:1319)
        np.testing.assert_allclose(seen["x"].cpu().numpy(), g[f"xt::{path}"], **tol)
        np.testing.assert_allclose(terms["mse"].detach().cpu().numpy(), g[f"mse::{path}::{mean}::{wt}"], rtol=1e-5)
        np.testing.assert_allclose(out.grad.cpu().numpy(), g[f"grad::{path}::{mean}::{wt}"], rtol=2e-5, atol=1e-9)
    with pytest.raises(ValueError):
        bad = gd.FlowMatching(args=gd.default_args(path_type=path, weight_type="snr"), model_mean_type=gd.ModelMeanType.VECTOR)
        bad.training_losses(lambda x, ts, **k: mo, x0, None, t=t, noise=eps)


def test_seeded_draw_order_noise_then_t():
    """training_losses(model, x) with t=None, noise=None draws `randn_like(x_start)` FIRST and the timesteps SECOND from
    the device generator (reference :849-852 diffusion, :1300-1303 flow): the seeded implicit call must equal the
    explicit call fed with draws made in that order, and must differ from the opposite order."""
    x0 = torch.randn(8, 4, 16, 16, device=DEV)
    rec = []
    def model(x, ts, **k):
        rec.append((x.clone(), ts.clone()))
        return torch.zeros_like(x)
    d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda")
    torch.manual_seed(77)
    a = d.training_losses(model, x0)
    torch.manual_seed(77)
    noise = torch.randn_like(x0)
    t = torch.randint(0, d.num_timesteps, (8,), device=DEV)
    b = d.training_losses(model, x0, t=t, noise=noise)
    assert torch.equal(rec[0][0], rec[1][0]) and torch.equal(rec[0][1], rec[1][1]) and torch.equal(a["mse"], b["mse"])
    assert torch.equal(rec[0][1], d._scale_timesteps(t))
    torch.manual_seed(77)
    t_first = torch.randint(0, d.num_timesteps, (8,), device=DEV)   # the opposite order consumes the stream differently
    assert not torch.equal(t_first, t)
    # flow matching: noise, then torch.rand (uniform) or torch.randn -> sigmoid (lognorm), reference :1259-1270
    for dist, draw in ((["uniform"], lambda: torch.rand(8, device=DEV)),
                       (["lognorm", 0.0, 1.0], lambda: torch.sigmoid(torch.randn(8, device=DEV) * 1.0 + 0.0))):
        fm = gd.FlowMatching(args=gd.default_args(path_type="linear", time_dist=dist), model_mean_type=gd.ModelMeanType.VECTOR)
        rec.clear()
        torch.manual_seed(5)
        a = fm.training_losses(model, x0)
        torch.manual_seed(5)
        noise = torch.randn_like(x0)
        tt = draw()
        b = fm.training_losses(model, x0, t=tt, noise=noise)
        assert torch.equal(rec[0][1], tt) and torch.equal(rec[0][0], rec[1][0]) and torch.equal(a["mse"], b["mse"])


@pytest.mark.parametrize("N,mean", [(64, "epsilon"), (5, "velocity"), (600, "epsilon"), (3, "previous_x")])
def test_in_kernel_philox_matches_torch_generator_stream(N, mean):
    """with noise=None the Gaussian noise (and t=None: the timesteps) are drawn INSIDE K1 / K2 from the
    device generator.  They must be the very tensors torch.randn_like / torch.randint would have produced, and the
    generator must end where it would have ended: the implicit call equals the explicit one bit for bit (x_t, t, mse, the
    gradient), and the next torch draw after either call is identical.  N = 600 latents is 2.4 M elements: more than one
    Philox call per thread plus a ragged tail of ATen's grid."""
    import ctypes as C
    from vaw_b200 import _lib as L
    d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type=mean, weight_type="lambda" if mean != "previous_x" else "constant")
    x0 = torch.randn(N, 4, 32, 32, device=DEV)
    out = torch.randn(N, 4, 32, 32, device=DEV)
    rec = []
    def run(**kw):
        o = out.clone().requires_grad_(True)
        def model(x, ts, **k):
            rec.append((x.clone(), ts.clone()))
            return o
        terms = d.training_losses(model, x0, **kw)
        terms["loss"].sum().backward()
        return terms["mse"].detach().clone(), o.grad.clone(), torch.rand(7, device=DEV)
    torch.manual_seed(1234)
    a = run()
    torch.manual_seed(1234)
    noise = torch.randn_like(x0)
    t = torch.randint(0, d.num_timesteps, (N,), device=DEV)
    b = run(t=t, noise=noise)
    assert torch.equal(rec[0][0], rec[1][0]), "x_t differs: in-kernel noise != torch.randn_like"
    assert torch.equal(rec[0][1], rec[1][1]), "timesteps differ: in-kernel randint != torch.randint"
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert torch.equal(a[2], b[2]), "the generator did not end at the same offset"
    # the raw draws through the C ABI, against the tensors themselves
    torch.manual_seed(99)
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    seed, off = gen.initial_seed(), gen.get_offset()
    want = torch.randn_like(x0)
    got, xt = torch.empty_like(x0), torch.empty_like(x0)
    ta, ts_, c0, c1, _ = d._tables(x0.device)
    L.call("vaw_qsample_philox", x0.data_ptr(), None, 1.0, seed, 0, off, t.data_ptr(), ta.data_ptr(), ts_.data_ptr(),
           L.ptr(c0), L.ptr(c1), None, got.data_ptr(), xt.data_ptr(), None, d.model_mean_type.value, N, 4096, L.stream_ptr())
    assert torch.equal(got, want)
    inc = C.c_ulonglong()
    L.call("vaw_philox_offset_increment", x0.numel(), C.byref(inc))
    assert gen.get_offset() == off + inc.value


def test_deferred_latent_fuses_sample_from_latent_into_k1():
    """trainer.py:21-25 + gaussian_diffusion.py:849-854 as one kernel: sample_from_latent(..., defer=True) hands the
    8-channel latent to training_losses, which draws eps1, noise and t in the reference's order.  Equal, bit for bit, to
    the two-step path under the same seed - for the eps objective (x_start never materialised) and for one that needs it."""
    from vaw_b200.tools import trainer as vtr
    lat = torch.randn(48, 8, 32, 32, device=DEV)
    lat[:, 4:] = lat[:, 4:].abs() * 0.3
    for mean in ("epsilon", "start_x"):
        d = gd.create_gaussian_diffusion(noise_schedule="linear", mean_type=mean, weight_type="lambda")
        out = torch.randn(48, 4, 32, 32, device=DEV)
        seen = []
        def run(defer):
            o = out.clone().requires_grad_(True)
            def model(x, ts, **k):
                seen.append((x.clone(), ts.clone()))
                return o
            torch.manual_seed(7)
            x_start = vtr.sample_from_latent(lat, 0.18215, defer=defer)
            terms = d.training_losses(model, x_start)
            terms["loss"].mean().backward()
            return terms["mse"].detach().clone(), o.grad.clone(), torch.rand(3, device=DEV)
        a, b = run(True), run(False)
        assert torch.equal(seen[-2][0], seen[-1][0]) and torch.equal(seen[-2][1], seen[-1][1])
        assert all(torch.equal(u, v) for u, v in zip(a, b))
    x = vtr.sample_from_latent(lat, 0.5, defer=True)
    assert x.shape == (48, 4, 32, 32) and x.tensor().shape == (48, 4, 32, 32)


def test_align_loss_kernel_vs_torch():
    torch.manual_seed(4)
    zs = (torch.randn(4, 64, 48, device=DEV)).bfloat16().requires_grad_(True)
    feat = torch.randn(4, 64, 48, device=DEV)
    loss = gd.compute_align_loss(feat, zs, "mse")
    (loss * 0.5).backward()
    z2 = zs.detach().float().requires_grad_(True)
    ref = torch.nn.functional.mse_loss(z2, feat)
    (ref * 0.5).backward()
    assert abs(loss.item() - ref.item()) / ref.item() < 1e-5
    assert ((zs.grad.float() - z2.grad).norm() / z2.grad.norm()).item() < 5e-3
    with pytest.raises(ValueError):
        gd.compute_align_loss(feat, zs, "nope")


@pytest.mark.parametrize("kind", ("cosine", "mse_l2"))
@pytest.mark.parametrize("dtype,tol_v,tol_g", [(torch.float32, 1e-5, 2e-5), (torch.bfloat16, 1e-5, 6e-3)])
def test_rowwise_align_losses_vs_torch(kind, dtype, tol_v, tol_g):
    """compute_align_loss 'cosine' / 'mse_l2' (reference :1008-1017): vaw_align_rowwise against the reference's own
    expressions (F.cosine_similarity / F.normalize + F.mse_loss) evaluated in fp32 on the same values, value + gradient;
    REPA shapes [N, 256, 768] and a ragged width."""
    import torch.nn.functional as F
    torch.manual_seed(5)
    for shape in ((4, 256, 768), (3, 17, 50)):
        zs = torch.randn(*shape, device=DEV).to(dtype).requires_grad_(True)
        feat = (torch.randn(*shape, device=DEV) * 3).to(dtype)
        loss = gd.compute_align_loss(feat, zs, kind)
        (loss * 0.7).backward()
        z2 = zs.detach().float().requires_grad_(True)
        f2 = feat.float()
        ref = (-F.cosine_similarity(f2, z2, dim=-1).mean() if kind == "cosine"
               else F.mse_loss(F.normalize(z2, dim=-1), F.normalize(f2, dim=-1)))
        (ref * 0.7).backward()
        assert abs(loss.item() - ref.item()) <= tol_v * abs(ref.item()) + 1e-7
        assert zs.grad.dtype == dtype
        assert ((zs.grad.float() - z2.grad).norm() / z2.grad.norm()).item() < tol_g
    with pytest.raises(ValueError):
        gd.compute_align_loss(feat[:, :5], zs, kind)


def test_sample_from_latent_bit_exact_and_seeded():
    """tools/trainer.py:21-25: the kernel against the reference golden (bit-exact, explicit noise through the C ABI) and
    the Python mirror against the same formula in torch with the same generator state."""
    import ctypes as C
    from vaw_b200 import _lib as L
    from vaw_b200.tools import trainer as vtr
    g = np.load(os.path.join(G, "diffusion_golden.npz"))
    lat, eps = torch.from_numpy(g["latent"]).to(DEV), torch.from_numpy(g["latent_eps"]).to(DEV)
    out = torch.empty_like(eps)
    L.call("vaw_sample_from_latent", lat.data_ptr(), eps.data_ptr(), out.data_ptr(), lat.shape[0], eps[0].numel(),
           0.18215, L.stream_ptr())
    assert np.array_equal(out.cpu().numpy(), g["latent_out"])
    big = torch.randn(64, 8, 32, 32, device=DEV)
    torch.manual_seed(123)
    got = vtr.sample_from_latent(big, 0.18215)
    torch.manual_seed(123)
    mean, std = torch.chunk(big, 2, dim=1)
    want = (mean + std * torch.randn_like(mean)) * 0.18215
    assert torch.equal(got, want)
    assert vtr.sample_from_latent(big[:0], 1.0).shape == (0, 4, 32, 32)
    with pytest.raises(L.VawError):
        vtr.sample_from_latent(big.cpu())
