"""GPU tier: the EDM sampler (vaw_b200.tools.cfg_edm: Net + ablation_sampler over vaw_edm_pre / vaw_edm_post) and the
flow-matching SDE sampler (FlowMatching.sde_sample over vaw_flow_sde_step) against
  * the fixtures produced by executing the reference on the CPU (edm_golden.npz, flow_golden.npz) with the fixtures' toy
    denoisers re-expressed without device transcendentals - BIT-EXACT: every float64 / float32 operation of the samplers is
    IEEE arithmetic in the reference's order, the per-step scalars are evaluated on the host with the reference's ops;
  * the oracle driven by the SAME engine-backed DiT (guided by IntervalCFG): bit-exact trajectories for identical
    denoiser outputs."""
import os

import numpy as np
import pytest
import torch

from gpu_util import dezero
from oracle import edm as oedm
from vaw_b200.models.dit import DiT
from vaw_b200.tools import cfg_edm as edm
from vaw_b200.tools import gaussian_diffusion as gd
from vaw_b200.tools.sampler import IntervalCFG

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "edm_golden.npz"))
GF = np.load(os.path.join(os.path.dirname(__file__), "golden", "flow_golden.npz"))
from test_oracle_edm import CASES   # noqa: E402


class Toy(torch.nn.Module):
    """tests/golden/make_golden.py::_ToyDenoiser with sin(t / 100) and y / 10 tabulated on the host (the device's sin
    differs in the last bit from the host's, and torch's CUDA division by a scalar multiplies by the reciprocal); the
    remaining ops are single IEEE multiplies / adds, identical on both sides."""

    def __init__(self):
        super().__init__()
        self.register_buffer("tab", torch.sin(torch.arange(1000).float() / 100.0))
        self.register_buffer("ytab", torch.arange(1000).float() / 10.0)

    def forward(self, x, t, y=None, **kw):
        tt = self.tab[t.long()].view(-1, 1, 1, 1)
        yy = self.ytab[y.long()].view(-1, 1, 1, 1) if y is not None else 0.0
        return (0.3 * x + 0.05 * tt + 0.01 * yy).to(x.dtype)


@pytest.mark.parametrize("tag", sorted(CASES))
def test_ablation_sampler_matches_reference_bit_exact(tag):
    nkw, skw = CASES[tag]
    net = edm.Net(Toy().to(DEV), img_resolution=8, img_channels=3, **nkw).to(DEV)
    it = iter(torch.from_numpy(G["noises"]).to(DEV))
    got = edm.ablation_sampler(net, torch.from_numpy(G["latents"]).to(DEV), class_labels=torch.from_numpy(G["labels"]).to(DEV),
                               randn_like=lambda a: next(it).to(a.dtype), num_steps=7, **skw)
    assert got.dtype == torch.float64
    assert np.array_equal(got.cpu().numpy(), G[f"sample::{tag}"])


@pytest.mark.parametrize("sched", ("linear", "cosine", "linear_logsnr"))
def test_net_forward_matches_reference(sched):
    x = torch.from_numpy(G[f"net_x::{sched}"]).to(DEV)
    y = torch.from_numpy(G["labels"]).to(DEV)
    for pred in ("EPSILON", "START_X", "VELOCITY"):
        net = edm.Net(Toy().to(DEV), img_resolution=8, img_channels=3, noise_schedule=sched, pred_type=pred).to(DEV)
        got = net(x, torch.tensor(2.5, dtype=torch.float64), y)
        assert np.array_equal(got.cpu().numpy(), G[f"net_out::{sched}::{pred}"])


def test_edm_heun_with_guided_dit_vs_oracle():
    """The recipe's sampling stack: DiT -> IntervalCFG (guidance inside [200, 800)) -> Net -> Heun.  The oracle integrates
    on the CPU calling the same guided model; identical denoiser outputs -> identical float64 trajectories."""
    torch.manual_seed(0)
    m = DiT(image_size=16, patch_size=2, in_channels=4, hidden_size=128, depth=2, num_heads=2, class_dropout_prob=0.1,
            num_classes=10).to(DEV).eval()
    dezero(m)
    cfg = IntervalCFG(m, num_classes=10, guidance_scale=2.0, interval=(200.0, 800.0)).eval()
    net = edm.Net(cfg, img_resolution=16, img_channels=4, pred_type="EPSILON", noise_schedule="cosine").to(DEV)
    N = 4
    lat = torch.randn(N, 4, 16, 16, device=DEV)
    y = torch.randint(0, 10, (N,), device=DEV)
    noises = [torch.randn(N, 4, 16, 16, dtype=torch.float64) for _ in range(6)]
    it = iter(noises)
    got = edm.ablation_sampler(net, lat, class_labels=y, randn_like=lambda a: next(it).to(a), num_steps=6, solver="heun")
    calls = []

    def model_fn(x_in, t):
        calls.append(int(t[0]))
        with torch.no_grad():
            raw = cfg(x_in.to(DEV), t.to(DEV), y=y)        # outside the guidance interval: the DiT's (out, zs) tuple
            return (raw[0] if isinstance(raw, tuple) else raw).float().cpu()
    want = oedm.sample(oedm.SigmaTable("cosine"), "EPSILON", model_fn, lat.cpu(), noises, num_steps=6, solver="heun")
    assert torch.isfinite(got).all() and np.array_equal(got.cpu().numpy(), want.numpy())
    assert len(calls) == 11 and any(200 <= c < 800 for c in calls) and any(not (200 <= c < 800) for c in calls)


SDE_CASES = [("linear", "VECTOR"), ("linear", "VELOCITY"), ("linear", "START_X"), ("linear_logsnr", "VECTOR"),
             ("linear_logsnr", "VELOCITY"), ("linear_logsnr", "START_X"), ("linear_logsnr", "EPSILON"), ("cosine", "VECTOR")]


@pytest.mark.parametrize("path,mean", SDE_CASES)
@pytest.mark.parametrize("solver", ("euler", "heun"))
def test_flow_sde_sample_matches_reference(path, mean, solver):
    fm = gd.FlowMatching(args=gd.default_args(path_type=path, sampler_type="sde"), model_mean_type=gd.ModelMeanType[mean])
    toy = lambda x, tm, **k: (0.25 * x - 0.1 * tm.view(-1, 1, 1, 1).to(x.dtype)).float()
    it = iter(torch.from_numpy(GF["sde_noises"]).to(DEV))
    orig = torch.randn_like
    torch.randn_like = lambda a, **k: next(it).to(a.dtype)
    try:
        got = fm.sample(toy, torch.from_numpy(GF["sde_start"]).to(DEV), DEV, num_steps=6, solver=solver)
    finally:
        torch.randn_like = orig
    want = GF[f"sde::{path}::{mean}::{solver}"]
    if path == "cosine":     # reference quirk: sqrt of a negative diffusion coefficient at t = 1 -> NaN everywhere
        assert np.isnan(want).all() and torch.isnan(got).all()
    else:
        assert np.array_equal(got.cpu().numpy(), want)


def test_flow_sampler_api_errors():
    fm = gd.FlowMatching(args=gd.default_args(path_type="linear", sampler_type="ode"), model_mean_type=gd.ModelMeanType.VECTOR)
    with pytest.raises(NotImplementedError):
        fm.sample(lambda x, t, **k: x, torch.zeros(1, 3, 4, 4, device=DEV), DEV)
    fm.sampler_type = "sde"
    with pytest.raises(ValueError):
        fm.sample(lambda x, t, **k: x, torch.zeros(1, 3, 4, 4, device=DEV), DEV, solver="rk4")
