"""CPU tier: the EDM sampler restatement (oracle/edm.py) against the fixture produced by EXECUTING the reference's
tools/cfg_edm.py (tests/golden/make_golden.py::edm_golden -> edm_golden.npz), and the host-side scalar logic of the
product's Net (sigma table, rounding, preconditioning coefficients) - no kernels involved."""
import os

import numpy as np
import pytest
import torch

from oracle import edm as oedm

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "edm_golden.npz"))
CASES = {
    "heun_edm_eps": (dict(pred_type="EPSILON", noise_schedule="linear"), dict(solver="heun")),
    "euler_edm_eps": (dict(pred_type="EPSILON", noise_schedule="linear"), dict(solver="euler")),
    "heun_edm_x0_cos": (dict(pred_type="START_X", noise_schedule="cosine"), dict(solver="heun")),
    "heun_edm_v_logsnr": (dict(pred_type="VELOCITY", noise_schedule="linear_logsnr"), dict(solver="heun")),
    "heun_vp": (dict(pred_type="EPSILON", noise_schedule="cosine"), dict(solver="heun", discretization="vp", schedule="vp", scaling="vp")),
    "heun_ve": (dict(pred_type="EPSILON", noise_schedule="linear"), dict(solver="heun", discretization="ve", schedule="ve", scaling="none")),
    "euler_iddpm": (dict(pred_type="EPSILON", noise_schedule="cosine"), dict(solver="euler", discretization="iddpm", schedule="linear", scaling="none")),
    "heun_churn": (dict(pred_type="EPSILON", noise_schedule="linear"), dict(solver="heun", S_churn=20, S_min=0.05, S_max=50, S_noise=1.003)),
    "heun_alpha": (dict(pred_type="VELOCITY", noise_schedule="cosine"), dict(solver="heun", alpha=0.5)),
}


def toy(labels):
    """The fixture's stand-in denoiser (tests/golden/make_golden.py::_ToyDenoiser)."""
    y = torch.as_tensor(labels)

    def fn(x, t):
        tt = torch.sin(t.float() / 100.0).view(-1, 1, 1, 1)
        return (0.3 * x + 0.05 * tt + 0.01 * (y.float().view(-1, 1, 1, 1) / 10.0)).to(x.dtype)
    return fn


@pytest.mark.parametrize("sched", ("linear", "cosine", "linear_logsnr"))
def test_sigma_table_rounding_and_preconditioning(sched):
    tab = oedm.SigmaTable(sched)
    assert np.array_equal(tab.u.numpy(), G[f"u::{sched}"])
    assert np.array_equal(np.array([tab.sigma_min, tab.sigma_max]), G[f"sigma_minmax::{sched}"])
    probe = torch.tensor([0.002, 0.01, 0.5, 1.0, 7.3, 80.0, 155.0], dtype=torch.float64)
    assert np.array_equal(tab.round(probe, return_index=True).numpy(), G[f"round_idx::{sched}"])
    assert np.array_equal(tab.round(probe).numpy(), G[f"round_val::{sched}"])
    x = torch.from_numpy(G[f"net_x::{sched}"])
    for pred in ("EPSILON", "START_X", "VELOCITY"):
        got = oedm.denoise(tab, pred, toy(G["labels"]), x, torch.tensor(2.5, dtype=torch.float64), 3)
        assert np.array_equal(got.numpy(), G[f"net_out::{sched}::{pred}"])


@pytest.mark.parametrize("tag", sorted(CASES))
def test_ablation_sampler_bit_exact(tag):
    nkw, skw = CASES[tag]
    tab = oedm.SigmaTable(nkw["noise_schedule"])
    got = oedm.sample(tab, nkw["pred_type"], toy(G["labels"]), torch.from_numpy(G["latents"]),
                      [torch.from_numpy(n) for n in G["noises"]], num_steps=7, **skw)
    assert got.dtype == torch.float64 and np.array_equal(got.numpy(), G[f"sample::{tag}"])


@pytest.mark.parametrize("sched", ("linear", "cosine", "linear_logsnr"))
def test_product_net_host_logic(sched):
    """vaw_b200.tools.cfg_edm.Net without touching a kernel: sigma table, round_sigma, scalar coefficients."""
    from vaw_b200.tools.cfg_edm import Net
    net = Net(torch.nn.Identity(), img_resolution=8, img_channels=3, noise_schedule=sched, pred_type="VELOCITY")
    assert np.array_equal(net.u.numpy(), G[f"u::{sched}"])
    assert (net.sigma_min, net.sigma_max) == tuple(G[f"sigma_minmax::{sched}"])
    probe = torch.tensor([0.002, 0.01, 0.5, 1.0, 7.3, 80.0, 155.0], dtype=torch.float64)
    assert np.array_equal(net.round_sigma(probe, return_index=True).numpy(), G[f"round_idx::{sched}"])
    assert np.array_equal(net.round_sigma(probe).numpy(), G[f"round_val::{sched}"])
    tab = oedm.SigmaTable(sched)
    for sigma in (0.01, 0.7, 2.5, 80.0):
        k = net.coefficients(torch.tensor(sigma, dtype=torch.float64))
        s32 = torch.tensor(sigma, dtype=torch.float64).to(torch.float32)
        c_in = 1 / (s32 ** 2 + 1).sqrt()
        assert k["c_in"] == float(c_in) and k["c_skip"] == float(c_in ** 2) and k["c_out"] == float(-s32 * c_in)
        assert k["c_noise"] == int(999 - tab.round(s32.reshape(1), return_index=True))
    with pytest.raises(ValueError):
        Net(torch.nn.Identity(), 8, 3, pred_type="SCORE")
