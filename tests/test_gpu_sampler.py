"""GPU tier: K3 — timestep importance sampling and history update: bit-exact vs numpy / the oracle / the golden
fixtures of the executed reference (indices, importance weights, history, and the host RNG stream)."""
import os

import numpy as np
import pytest
import torch

from oracle import resample as ors
from oracle.train_step import synthetic_history
from vaw_b200.tools import gaussian_diffusion as gd
from vaw_b200.tools import resample as rs

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"


def test_matches_reference_golden():
    sg = np.load(os.path.join(G, "sampler_golden.npz"))
    d = gd.create_gaussian_diffusion()
    s = rs.LossSecondMomentResampler(d)
    s.load_history(sg["hist"], np.full(1000, 10), DEV)
    assert s._warmed_up()
    assert np.array_equal(s.weights(), sg["weights"])
    np.random.seed(2024)
    idx, w = s.sample(48, DEV)
    assert idx.dtype == torch.int64 and w.dtype == torch.float32
    assert np.array_equal(idx.cpu().numpy(), sg["sample_idx"]) and np.array_equal(w.cpu().numpy(), sg["sample_w"])
    assert np.random.random_sample() == sg["next_uniform_after_sample"][0]
    cold = rs.LossSecondMomentResampler(d)
    np.random.seed(77)
    idx, w = cold.sample(32, DEV)
    assert np.array_equal(idx.cpu().numpy(), sg["cold_idx"]) and np.array_equal(w.cpu().numpy(), sg["cold_w"])
    np.random.seed(5)
    idx, w = rs.UniformSampler(d).sample(40, DEV)
    assert np.array_equal(idx.cpu().numpy(), sg["uni_idx"]) and np.array_equal(w.cpu().numpy(), sg["uni_w"])
    upd = rs.LossSecondMomentResampler(d)
    upd.load_history(np.zeros((1000, 10)), np.zeros(1000), DEV)
    upd.update_with_all_losses(sg["upd_ts"].tolist(), [float(x) for x in sg["upd_losses"]])
    assert np.array_equal(upd._loss_history, sg["upd_hist"]) and np.array_equal(upd._loss_counts, sg["upd_counts"])


@pytest.mark.parametrize("trial", range(4))
def test_vs_numpy_choice_random_histories(trial):
    d = gd.create_gaussian_diffusion()
    rng = np.random.RandomState(trial)
    hist = np.abs(rng.randn(1000, 10)) * np.exp(rng.randn(1000, 1))
    s = rs.LossSecondMomentResampler(d)
    s.load_history(hist, np.full(1000, 10), DEV)
    w_ref = np.sqrt(np.mean(hist ** 2, axis=-1)); w_ref /= np.sum(w_ref); w_ref *= 1 - 0.001; w_ref += 0.001 / 1000
    assert np.array_equal(s.weights(), w_ref)
    for B in (1, 64, 256, 1000):
        p = w_ref / np.sum(w_ref)
        np.random.seed(trial * 10 + B)
        ref = np.random.choice(1000, size=(B,), p=p)
        nxt = np.random.random_sample()
        np.random.seed(trial * 10 + B)
        idx, w = s.sample(B, DEV)
        assert np.array_equal(idx.cpu().numpy(), ref)
        assert np.array_equal(w.cpu().numpy(), (1 / (1000 * p[ref])).astype(np.float32))
        assert np.random.random_sample() == nxt


def test_history_update_duplicates_and_wraparound_vs_oracle():
    d = gd.create_gaussian_diffusion()
    s = rs.LossSecondMomentResampler(d)
    s.load_history(np.zeros((1000, 10)), np.zeros(1000), DEV)
    h, c = np.zeros((1000, 10)), np.zeros(1000, dtype=int)
    rng = np.random.RandomState(0)
    for _ in range(50):
        ts = rng.randint(0, 40, size=64)
        ls = rng.rand(64).astype(np.float32)
        s.update_with_local_losses(torch.from_numpy(ts).to(DEV), torch.from_numpy(ls).to(DEV))
        ors.update_history(h, c, ts.tolist(), ls.tolist())
    assert np.array_equal(s._loss_history, h) and np.array_equal(s._loss_counts, c)
    assert not s._warmed_up()
    # empty update is a no-op
    s.update_with_local_losses(torch.zeros(0, dtype=torch.long, device=DEV), torch.zeros(0, device=DEV))
    assert np.array_equal(s._loss_history, h)


def test_warmup_transition_and_idempotent_weights():
    """Property: weights() is a pure function of the history; the sampler turns non-uniform exactly when every
    timestep holds history_per_term entries (resample.py:161-162)."""
    d = gd.create_gaussian_diffusion()
    hist, counts = synthetic_history(0)
    counts = counts.copy(); counts[123] = 9
    s = rs.LossSecondMomentResampler(d)
    s.load_history(hist, counts, DEV)
    assert np.array_equal(s.weights(), np.ones(1000))
    s.update_with_local_losses(torch.tensor([123], device=DEV), torch.tensor([0.25], device=DEV))
    w1, w2 = s.weights(), s.weights()
    assert np.array_equal(w1, w2) and not np.array_equal(w1, np.ones(1000))
    hist2 = hist.copy(); hist2[123, 9] = np.float64(np.float32(0.25))
    assert np.array_equal(w1, ors.second_moment_weights(hist2, np.full(1000, 10)))
    assert abs(w1.sum() - 1.0) < 1e-12


def test_history_survives_device_spelling():
    """'cuda' and 'cuda:0' name the same device: alternating them between sample() and update must not reset the
    device-resident history (regression)."""
    d = gd.create_gaussian_diffusion()
    s = rs.LossSecondMomentResampler(d)
    np.random.seed(1)
    for i in range(4):
        t, _ = s.sample(16, "cuda" if i % 2 == 0 else torch.device("cuda", 0))
        s.update_with_local_losses(t, torch.rand(16, device="cuda"))
    assert int(s._loss_counts.sum()) == 64
