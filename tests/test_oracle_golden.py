"""CPU tier: the oracle (oracle/) is pinned against fixtures produced by EXECUTING the reference
(tests/golden/make_golden.py) and against numpy itself."""
import os

import numpy as np
import pytest
import torch

from oracle import diffusion as odiff
from oracle import resample as ors
from oracle.dit import dit_forward

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def dg():
    return np.load(os.path.join(G, "diffusion_golden.npz"))


@pytest.fixture(scope="module")
def sg():
    return np.load(os.path.join(G, "sampler_golden.npz"))


@pytest.mark.parametrize("sched", ["linear", "cosine", "linear_logsnr"])
def test_schedules_and_tables_bit_exact(dg, sched):
    tb = odiff.tables(odiff.named_beta_schedule(sched, 1000))
    assert np.array_equal(tb["betas"], dg[f"betas_{sched}"])
    assert np.array_equal(tb["sqrt_alphas_cumprod"], dg[f"sqrt_ac_{sched}"])
    assert np.array_equal(tb["sqrt_one_minus_alphas_cumprod"], dg[f"sqrt_1mac_{sched}"])
    assert np.array_equal(tb["posterior_mean_coef1"], dg[f"pmc1_{sched}"])
    assert np.array_equal(tb["posterior_mean_coef2"], dg[f"pmc2_{sched}"])


def _weight_cases(dg):
    for k in dg.files:
        if k.startswith("w_"):
            _, sched, rest = k.split("_", 2)
            for mean in odiff.MEAN_TYPES:
                if rest.startswith(mean + "_"):
                    yield k, sched, mean, rest[len(mean) + 1:]


def test_loss_weight_all_timesteps(dg):
    n = 0
    for key, sched, mean, wt in _weight_cases(dg):
        tb = odiff.tables(odiff.named_beta_schedule(sched, 1000))
        w = odiff.weight_lut(tb, mean, wt)
        ref = dg[key]
        if wt == "p2":  # powf: allow 1 ulp between numpy and torch
            np.testing.assert_allclose(w, ref, rtol=2e-7)
        else:
            assert np.array_equal(w, ref), key
        n += 1
    assert n >= 30


def test_loss_weight_rejects_invalid():
    a = np.array([0.5], np.float32)
    with pytest.raises(ValueError):
        odiff.loss_weight("VELOCITY", "debias", a, a)
    with pytest.raises(ValueError):
        odiff.loss_weight("EPSILON", "no_such", a, a)


@pytest.mark.parametrize("mean", ["EPSILON", "START_X", "VELOCITY", "PREVIOUS_X"])
def test_qsample_target_bit_exact(dg, mean):
    tb = odiff.tables(odiff.named_beta_schedule("linear", 1000))
    assert np.array_equal(odiff.q_sample(tb, dg["x0"], dg["t"], dg["eps"]), dg[f"xt_{mean}"])
    assert np.array_equal(odiff.target(tb, mean, dg["x0"], dg["t"], dg["eps"]), dg[f"target_{mean}"])


def test_sample_from_latent_bit_exact(dg):
    assert np.array_equal(odiff.sample_from_latent(dg["latent"], dg["latent_eps"], 0.18215), dg["latent_out"])


@pytest.mark.parametrize("mean,wt", [("EPSILON", "lambda"), ("START_X", "lambda"), ("VELOCITY", "min_snr_5"),
                                     ("PREVIOUS_X", "constant"), ("EPSILON", "min_snr_5")])
def test_training_losses_mse_and_grad(dg, mean, wt):
    tb = odiff.tables(odiff.named_beta_schedule("linear", 1000))
    mse, grad = odiff.mse_terms(tb, mean, wt, dg["x0"], dg["t"], dg["eps"], dg["model_out"])
    np.testing.assert_allclose(mse, dg[f"mse_{mean}_{wt}"], rtol=1e-5)
    # reference gradient is d mean_n(loss_n) / d out  ->  divide by N
    np.testing.assert_allclose(grad / len(dg["t"]), dg[f"grad_{mean}_{wt}"], rtol=1e-5, atol=1e-9)


def test_pairwise_sum_matches_numpy():
    rng = np.random.RandomState(0)
    for n in (1, 7, 8, 10, 100, 128, 129, 496, 1000, 1001, 4099):
        a = rng.rand(n) * np.exp(rng.randn(n))
        assert ors.np_pairwise_sum(a) == np.sum(a), n


def test_sampler_weights_sample_update(sg):
    hist = sg["hist"]
    counts = np.full(1000, 10)
    w = ors.second_moment_weights(hist, counts)
    assert np.array_equal(w, sg["weights"])
    np.random.seed(2024)
    idx, iw = ors.sample(w, 48)
    assert np.array_equal(idx, sg["sample_idx"]) and np.array_equal(iw, sg["sample_w"])
    assert np.random.random_sample() == sg["next_uniform_after_sample"][0]  # RNG stream consumed identically
    # cold (not warmed-up) sampler is uniform
    wc = ors.second_moment_weights(np.zeros((1000, 10)), np.zeros(1000, dtype=int))
    assert np.array_equal(wc, sg["weights_cold"])
    np.random.seed(77)
    idx, iw = ors.sample(wc, 32)
    assert np.array_equal(idx, sg["cold_idx"]) and np.array_equal(iw, sg["cold_w"])
    np.random.seed(5)
    idx, iw = ors.sample(np.ones(1000), 40)
    assert np.array_equal(idx, sg["uni_idx"]) and np.array_equal(iw, sg["uni_w"])
    h, c = ors.update_history(np.zeros((1000, 10)), np.zeros(1000, dtype=int), sg["upd_ts"],
                              [float(x) for x in sg["upd_losses"]])
    assert np.array_equal(h, sg["upd_hist"]) and np.array_equal(c, sg["upd_counts"])


def test_sample_matches_np_random_choice():
    rng = np.random.RandomState(3)
    for trial in range(5):
        w = rng.rand(1000) + 1e-3
        p = w / np.sum(w)
        np.random.seed(100 + trial)
        ref = np.random.choice(1000, size=(64,), p=p)
        nxt = np.random.random_sample()
        np.random.seed(100 + trial)
        idx, _ = ors.sample(w, 64)
        assert np.array_equal(idx, ref) and np.random.random_sample() == nxt


def test_dit_oracle_matches_reference_forward_and_grads():
    g = np.load(os.path.join(G, "dit_golden.npz"))
    sd = {k[len("param::"):]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("param::")}
    for k, v in sd.items():
        v.requires_grad_(k != "pos_embed")
    tb = odiff.tables(odiff.named_beta_schedule("cosine", 1000))
    x0, eps, t, y = (torch.from_numpy(g[k]) for k in ("x0", "eps", "t", "y"))
    feats = torch.from_numpy(g["feats"])
    kw = dict(patch_size=2, num_heads=1, depth=2, learn_align=True, encoder_depth=1)
    x_t = torch.from_numpy(odiff.q_sample(tb, x0.numpy(), t.numpy(), eps.numpy()))
    out, zs = dit_forward(sd, x_t, t.float() * 1.0, torch.from_numpy(g["y"]), **kw)
    np.testing.assert_allclose(out.detach().numpy(), g["fwd_out"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(zs.detach().numpy(), g["fwd_zs"], rtol=1e-5, atol=1e-6)
    model_fn = lambda xt, ts: dit_forward(sd, xt, ts, y, **kw)
    terms = odiff.training_losses_torch(tb, "EPSILON", "lambda", model_fn, x0, t, eps, features=feats, gamma=0.5,
                                        learn_align=True)
    terms["loss"].mean().backward()
    np.testing.assert_allclose(terms["mse"].detach().numpy(), g["mse"], rtol=1e-5)
    np.testing.assert_allclose(terms["align"].detach().numpy(), g["align"], rtol=1e-5)
    checked = 0
    for k in g.files:
        if k.startswith("grad::"):
            ref = g[k]
            got = sd[k[len("grad::"):]].grad.numpy()
            denom = np.linalg.norm(ref) + 1e-12
            assert np.linalg.norm(got - ref) / denom < 1e-5, k
            checked += 1
    assert checked > 30


def test_uvit_oracle_matches_reference_forward_and_grads():
    from oracle.uvit import uvit_forward
    g = np.load(os.path.join(G, "uvit_golden.npz"))
    sd = {k[len("param::"):]: torch.from_numpy(g[k]).clone().requires_grad_(True) for k in g.files if k.startswith("param::")}
    tb = odiff.tables(odiff.named_beta_schedule("linear", 1000))
    x0, eps, t, y = (torch.from_numpy(g[k]) for k in ("x0", "eps", "t", "y"))
    kw = dict(patch_size=2, num_heads=1, depth=3)
    x_t = torch.from_numpy(odiff.q_sample(tb, x0.numpy(), t.numpy(), eps.numpy()))
    out = uvit_forward(sd, x_t, t.float(), y, **kw)
    np.testing.assert_allclose(out.detach().numpy(), g["fwd_out"], rtol=1e-5, atol=1e-6)
    terms = odiff.training_losses_torch(tb, "EPSILON", "lambda", lambda xt, ts: uvit_forward(sd, xt, ts, y, **kw), x0, t, eps)
    terms["loss"].mean().backward()
    np.testing.assert_allclose(terms["mse"].detach().numpy(), g["mse"], rtol=1e-5)
    n = 0
    for k in g.files:
        if k.startswith("grad::"):
            ref, got = g[k], sd[k[len("grad::"):]].grad.numpy()
            assert np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-12) < 1e-5, k
            n += 1
    assert n > 40
