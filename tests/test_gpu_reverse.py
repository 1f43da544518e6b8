"""GPU parity of the fused reverse-process step (vaw_reverse_step / vaw_cfg_combine through the host mirror's
p_mean_variance / p_sample / ddim_sample / ddim_reverse_sample / IntervalCFG) against
  * tests/golden/reverse_golden.npz (the executed reference, tools/gaussian_diffusion.py:278-689), and
  * the oracle (oracle/diffusion.py) on seeded inputs, including ragged sizes and a full-size latent batch.
Tolerances: bit-exact for everything made of +,-,*,/,sqrt; where exp() enters (p_sample's noise scale, the learned
variances) device and host libm may differ in the last bit: rtol 1e-6 / atol 2e-7 x the largest magnitude (the
last bit of a large term survives a cancelling sum)."""
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import diffusion as odiff

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [("EPSILON", "FIXED_LARGE", "f32", True), ("EPSILON", "FIXED_SMALL", "f32", False),
         ("EPSILON", "FIXED_LARGE", "bf16", True), ("START_X", "FIXED_LARGE", "f32", True),
         ("PREVIOUS_X", "FIXED_SMALL", "f32", True), ("EPSILON", "LEARNED_RANGE", "f32", True),
         ("EPSILON", "LEARNED_RANGE", "bf16", True), ("EPSILON", "LEARNED", "f32", True),
         ("START_X", "LEARNED", "bf16", False)]


def make(mean, var, schedule="linear"):
    from vaw_b200.tools import gaussian_diffusion as gd
    return gd.GaussianDiffusion(args=gd.default_args(), betas=gd.get_named_beta_schedule(schedule, 1000),
                                model_mean_type=gd.ModelMeanType[mean], model_var_type=gd.ModelVarType[var],
                                loss_type=gd.LossType.MSE, rescale_timesteps=True)


def close(a, b, loose=False):
    a = a.float().cpu().numpy() if torch.is_tensor(a) else a
    if loose:
        np.testing.assert_allclose(a, b, rtol=1e-6, atol=2e-7 * max(1.0, float(np.abs(b).max(initial=0.0))))
    else:
        np.testing.assert_array_equal(a, b)


def run_all(d, mo, x, t, z, clip):
    model = lambda xx, ts, **k: mo
    pmv = d.p_mean_variance(model, x, t, clip_denoised=clip)
    return dict(pmv=pmv,
                p_sample=d._reverse(0, model, x, t, clip_denoised=clip, denoised_fn=None, cond_fn=None,
                                    model_kwargs=None, noise=z),
                ddim0=d._reverse(1, model, x, t, clip_denoised=clip, denoised_fn=None, cond_fn=None, model_kwargs=None,
                                 noise=z, eta=0.0),
                ddim7=d._reverse(1, model, x, t, clip_denoised=clip, denoised_fn=None, cond_fn=None, model_kwargs=None,
                                 noise=z, eta=0.7),
                ddimrev=d.ddim_reverse_sample(model, x, t, clip_denoised=clip))


@pytest.mark.parametrize("mean,var,dt,clip", CASES)
def test_reverse_step_matches_reference_fixture(mean, var, dt, clip):
    rg = np.load(os.path.join(G, "reverse_golden.npz"))
    dev = torch.device("cuda", 0)
    mo = rg["model_out2"] if var.startswith("LEARNED") else np.ascontiguousarray(rg["model_out2"][:, :3])
    mo = torch.from_numpy(mo).to(dev)
    if dt == "bf16":
        mo = mo.bfloat16()
    x, z, t = (torch.from_numpy(rg[k]).to(dev) for k in ("x", "noise", "t"))
    key = f"{mean}_{var}_{dt}_{int(clip)}"
    r = run_all(make(mean, var), mo, x, t, z, clip)
    learned = var.startswith("LEARNED")
    close(r["pmv"]["pred_xstart"], rg[f"pmv_pred_xstart::{key}"])
    close(r["pmv"]["mean"], rg[f"pmv_mean::{key}"])
    close(r["pmv"]["log_variance"], rg[f"pmv_log_variance::{key}"])
    close(r["pmv"]["variance"], rg[f"pmv_variance::{key}"], loose=learned)
    close(r["p_sample"]["sample"], rg[f"p_sample::{key}"], loose=True)
    close(r["p_sample"]["pred_xstart"], rg[f"pmv_pred_xstart::{key}"])
    # the fixture's sqrt() ran on torch's CPU kernel, which is not correctly rounded (tests/test_oracle_reverse.py)
    close(r["ddim0"]["sample"], rg[f"ddim0::{key}"], loose=True)
    close(r["ddim7"]["sample"], rg[f"ddim7::{key}"], loose=True)
    close(r["ddimrev"]["sample"], rg[f"ddimrev::{key}"], loose=True)


@pytest.mark.parametrize("mean,var,dt,clip", CASES + [("VELOCITY", "FIXED_LARGE", "f32", True)])
@pytest.mark.parametrize("shape", [(5, 3, 7, 9), (16, 4, 32, 32)])
def test_reverse_step_bit_exact_against_oracle(mean, var, dt, clip, shape):
    """IEEE sqrt/div on both sides: everything without exp() must agree to the bit, ragged (scalar path) and
    vectorised shapes alike."""
    dev = torch.device("cuda", 0)
    rng = np.random.RandomState(zlib.crc32(repr((mean, var, dt, shape)).encode()))
    N, C = shape[:2]
    learned = var.startswith("LEARNED")
    x = (rng.randn(*shape) * 1.3).astype(np.float32)
    z = rng.randn(*shape).astype(np.float32)
    mo = rng.randn(N, 2 * C if learned else C, *shape[2:]).astype(np.float32)
    t = rng.randint(0, 1000, size=N)
    t[:3] = [0, 999, 1]
    bf = dt == "bf16"
    if bf:
        mo = odiff._bf16_round(mo)
    sched = "cosine" if mean == "VELOCITY" else "linear"
    tb = odiff.tables(odiff.named_beta_schedule(sched, 1000))
    mo_d = torch.from_numpy(mo).to(dev)
    if bf:
        mo_d = mo_d.bfloat16()
    r = run_all(make(mean, var, sched), mo_d, torch.from_numpy(x).to(dev), torch.from_numpy(t).to(dev),
                torch.from_numpy(z).to(dev), clip)
    o = odiff.p_mean_variance(tb, mean, var, mo, x, t, clip, bf)
    close(r["pmv"]["pred_xstart"], o["pred_xstart"])
    close(r["pmv"]["mean"], o["mean"])
    close(r["pmv"]["log_variance"], o["log_variance"])
    close(r["pmv"]["variance"], o["variance"], loose=learned)
    close(r["p_sample"]["sample"], odiff.p_sample(tb, mean, var, mo, x, t, z, clip, bf)["sample"], loose=True)
    close(r["ddim0"]["sample"], odiff.ddim_sample(tb, mean, var, mo, x, t, z, 0.0, clip, bf)["sample"])
    close(r["ddim7"]["sample"], odiff.ddim_sample(tb, mean, var, mo, x, t, z, 0.7, clip, bf)["sample"])
    close(r["ddimrev"]["sample"], odiff.ddim_reverse_sample(tb, mean, var, mo, x, t, clip, bf)["sample"])


def test_reverse_step_rejects_what_the_reference_rejects():
    from vaw_b200 import _lib as L
    dev = torch.device("cuda", 0)
    x = torch.randn(2, 3, 4, 4, device=dev)
    t = torch.tensor([5, 6], device=dev)
    d = make("EPSILON", "FIXED_LARGE")
    with pytest.raises(AssertionError):          # wrong channel count (:313)
        d.p_sample(lambda a, b, **k: torch.zeros(2, 6, 4, 4, device=dev), x, t)
    with pytest.raises(AssertionError):          # Reverse ODE only for deterministic path (:667)
        d.ddim_reverse_sample(lambda a, b, **k: x, x, t, eta=0.5)
    with pytest.raises(NotImplementedError):
        d.p_sample(lambda a, b, **k: x, x, t, cond_fn=lambda *a, **k: x)
    with pytest.raises(L.VawError):              # CPU tensors: no fallback
        d.p_sample(lambda a, b, **k: x.cpu(), x.cpu(), t.cpu())
    empty = d.p_sample(lambda a, b, **k: a, x[:0], t[:0])
    assert empty["sample"].shape == (0, 3, 4, 4)


def test_ddim_loop_round_trip_and_determinism():
    """Size-independent properties: with an exact eps-predictor for a fixed x0 the deterministic DDIM chain recovers
    x0, ddim_reverse_sample inverts a ddim step up to fp32 rounding, and the loop is reproducible under a seed."""
    dev = torch.device("cuda", 0)
    d = make("EPSILON", "FIXED_LARGE")
    x0 = torch.randn(4, 4, 32, 32, device=dev).clamp(-1, 1)
    tab = d._reverse_table(dev)

    class Exact(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1, device=dev))

        def forward(self, xt, ts, **kw):
            i = ts.round().long()
            r, rm1 = tab[0][i].view(-1, 1, 1, 1), tab[1][i].view(-1, 1, 1, 1)
            return (r * xt - x0) / rm1

    m = Exact()
    torch.manual_seed(3)
    a = d.ddim_sample_loop(m, (4, 4, 32, 32), clip_denoised=False)
    torch.manual_seed(3)
    b = d.ddim_sample_loop(m, (4, 4, 32, 32), clip_denoised=False)
    assert torch.equal(a, b)
    assert (a - x0).abs().max().item() < 2e-3
    torch.manual_seed(4)
    c = d.p_sample_loop(m, (4, 4, 32, 32))
    assert torch.isfinite(c).all() and (c - x0).abs().max().item() < 0.1
    t = torch.full((4,), 400, device=dev)
    xt = d.q_sample(x0, t, torch.randn_like(x0))
    nxt = d.ddim_reverse_sample(m, xt, t, clip_denoised=False)["sample"]
    back = d.ddim_sample(m, nxt, t + 1, clip_denoised=False)["sample"]
    assert (back - xt).abs().max().item() < 1e-3


@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_interval_cfg_matches_reference_fixture(dt):
    from vaw_b200.tools.sampler import IntervalCFG
    rg = np.load(os.path.join(G, "reverse_golden.npz"))
    dev = torch.device("cuda", 0)
    both = torch.from_numpy(rg["cfg_both"]).to(dev)
    seen = []

    class Rec(torch.nn.Module):
        def forward(self, xx, tt, **kw):
            seen.append(kw["y"].tolist())
            o = both if xx.shape[0] == 6 else both[:3]
            return o.bfloat16() if dt == "bf16" else o

    cfg = IntervalCFG(Rec(), num_classes=10, guidance_scale=2.5, interval=(100.0, 600.0))
    x = torch.from_numpy(rg["x"][:3]).to(dev)
    y = torch.tensor([1, 2, 3], device=dev)
    close(cfg(x, torch.full((3,), 300.0, device=dev), y=y), rg[f"cfg_in::{dt}"])
    close(cfg(x, torch.full((3,), 800.0, device=dev), y=y), rg[f"cfg_out::{dt}"])
    assert seen[0] == [1, 2, 3, 10, 10, 10] and seen[1] == [1, 2, 3]
    # against the oracle at a size that uses many CTAs
    big = torch.randn(2 * 64, 4, 32, 32, device=dev)
    big = big.bfloat16() if dt == "bf16" else big
    g = IntervalCFG(lambda a, b, **k: big, 10, 1.7)(torch.zeros(64, 4, 32, 32, device=dev),
                                                     torch.zeros(64, device=dev), y=torch.zeros(64, dtype=torch.long, device=dev))
    h = big.float().cpu().numpy()
    close(g, odiff.cfg_combine(h[:64], h[64:], 1.7, dt == "bf16"))


def test_dit_guided_ddim_chain_vs_oracle():
    """The reverse path end to end, as tools/sampler.py:112-141 drives it: eval-mode DiT -> IntervalCFG (doubled batch)
    -> SpacedDiffusion("ddim10").ddim_sample_loop.  Every step of the chain is checked (teacher-forced on the chain's
    own x_t) against the oracle: oracle DiT forward under bf16 autocast, oracle guidance combine, oracle DDIM step.
    bf16 denoiser -> 2e-2 rel-L2 per step."""
    from gpu_util import dezero, relerr
    from oracle.dit import dit_forward
    from vaw_b200.models.dit import DiT
    from vaw_b200.tools import gaussian_diffusion as gd
    from vaw_b200.tools.respace import SpacedDiffusion, space_timesteps
    from vaw_b200.tools.sampler import IntervalCFG
    dev = torch.device("cuda", 0)
    torch.manual_seed(5)
    m = DiT(image_size=16, patch_size=2, in_channels=4, hidden_size=128, depth=2, num_heads=2,
            class_dropout_prob=0.1, num_classes=10).to(dev)
    dezero(m)
    m.eval()
    d = SpacedDiffusion(use_timesteps=space_timesteps(1000, "ddim10"), args=gd.default_args(),
                        betas=gd.get_named_beta_schedule("cosine", 1000), model_mean_type=gd.ModelMeanType.EPSILON,
                        model_var_type=gd.ModelVarType.FIXED_LARGE, loss_type=gd.LossType.MSE, rescale_timesteps=True)
    scale = 1.5
    cfg = IntervalCFG(m, 10, scale).eval()
    B = 4
    y = torch.tensor([1, 7, 3, 9], device=dev)
    x = torch.randn(B, 4, 16, 16, device=dev)
    tb = odiff.tables(odiff.spaced_betas(odiff.named_beta_schedule("cosine", 1000), d.timestep_map))
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    steps = list(d.ddim_sample_loop_progressive(cfg, (B, 4, 16, 16), noise=x, model_kwargs={"y": y}))
    assert len(steps) == 10
    cur = x
    for i, out in zip(reversed(range(10)), steps):
        t_model = torch.full((2 * B,), float(d.timestep_map[i]), device=dev)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            both, _ = dit_forward(sd, torch.cat([cur, cur]), t_model, torch.cat([y, torch.full_like(y, 10)]),
                                  patch_size=2, num_heads=2, depth=2)
        both = both.bfloat16().float().cpu().numpy()
        guided = odiff.cfg_combine(both[:B], both[B:], scale, True)
        want = odiff.ddim_sample(tb, "EPSILON", "FIXED_LARGE", guided, cur.cpu().numpy(), np.full(B, i),
                                 np.zeros((B, 4, 16, 16), np.float32), 0.0, True, True)
        assert relerr(out["sample"], torch.from_numpy(want["sample"]).to(dev)) < 2e-2, i
        # x0 = r x_t - rm1 eps: the denoiser's bf16 error reaches the x0 prediction multiplied by rm1
        rm1 = float(tb["sqrt_recipm1_alphas_cumprod"][i])
        assert relerr(out["pred_xstart"], torch.from_numpy(want["pred_xstart"]).to(dev)) < 2e-2 * (1 + rm1), i
        cur = out["sample"]
    assert torch.isfinite(cur).all() and cur.abs().max().item() <= 1.0 + 1e-6   # t = 0: the clipped x0 prediction


@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_cfg_combine_ragged_sizes(dt):
    """Odd element counts take the scalar kernel; results must not depend on the path."""
    from vaw_b200.tools.sampler import IntervalCFG
    dev = torch.device("cuda", 0)
    for shape in ((3, 3, 5, 7), (1, 1, 1, 1), (2, 4, 6, 6)):
        torch.manual_seed(sum(shape))
        both = torch.randn(2 * shape[0], *shape[1:], device=dev)
        both = both.bfloat16() if dt == "bf16" else both
        n = shape[0]
        g = IntervalCFG(lambda a, b, **k: both, 10, 3.25)(torch.zeros(shape, device=dev), torch.zeros(n, device=dev),
                                                          y=torch.zeros(n, dtype=torch.long, device=dev))
        h = both.float().cpu().numpy()
        close(g, odiff.cfg_combine(h[:n], h[n:], 3.25, dt == "bf16"))


def test_out_of_range_timesteps_are_memory_safe():
    """t outside [0, T) must not read outside the table (clamped; the reference's gather would device-assert)."""
    dev = torch.device("cuda", 0)
    d = make("EPSILON", "FIXED_LARGE")
    x = torch.randn(4, 3, 8, 8, device=dev)
    t = torch.tensor([-5, 1000, 10**9, 999], device=dev)
    out = d.ddim_sample(lambda a, b, **k: a, x, t)["sample"]
    ref = d.ddim_sample(lambda a, b, **k: a, x, torch.tensor([0, 999, 999, 999], device=dev))["sample"]
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
