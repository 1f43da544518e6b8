"""CPU tier: the variational-bound restatement (oracle/vb.py) against the fixture produced by EXECUTING the reference's
training_losses with a learned variance (tests/golden/make_golden.py::vb_golden -> vb_golden.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import diffusion as odiff
from oracle import vb as ovb

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "vb_golden.npz"))
CASES = [("EPSILON", "LEARNED_RANGE", "MSE", "lambda"), ("EPSILON", "LEARNED", "RESCALED_MSE", "min_snr_5"),
         ("START_X", "LEARNED_RANGE", "KL", "constant"), ("EPSILON", "LEARNED_RANGE", "RESCALED_KL", "constant"),
         ("PREVIOUS_X", "LEARNED", "KL", "constant"), ("START_X", "LEARNED_RANGE", "RESCALED_MSE", "lambda")]


@pytest.mark.parametrize("sched", ("linear", "cosine"))
@pytest.mark.parametrize("mean,var,loss,wt", CASES)
def test_vb_training_losses_vs_reference(sched, mean, var, loss, wt):
    tb = odiff.tables(odiff.named_beta_schedule(sched, 1000))
    mo = torch.from_numpy(G["model_out"]).requires_grad_(True)
    terms = ovb.training_losses(tb, mean, var, loss, wt, mo, G["x0"], G["t"], G["eps"])
    terms["loss"].mean().backward()
    key = f"{sched}::{mean}::{var}::{loss}::{wt}"
    for k in ("mse", "vb", "loss"):
        if f"{k}::{key}" in G.files:
            np.testing.assert_allclose(terms[k].detach().numpy(), G[f"{k}::{key}"], rtol=3e-5, atol=1e-6, err_msg=k)
    want = G[f"grad::{key}"]
    got = mo.grad.numpy()
    assert np.abs(got - want).max() <= 3e-5 * np.abs(want).max() + 1e-9
    assert np.isfinite(want).all() and np.abs(want[:, 3:]).max() > 0          # the variance channels do get a gradient
    if loss in ("MSE", "RESCALED_MSE"):
        # the bound must not move the mean prediction (:902 `model_output.detach()`): the mean channels' gradient is the
        # weighted-MSE gradient alone
        tb_terms = odiff.mse_terms(tb, mean, wt, G["x0"], G["t"], G["eps"], G["model_out"][:, :3])[1] / len(G["t"])
        assert np.abs(got[:, :3] - tb_terms).max() <= 1e-9 + 1e-5 * np.abs(tb_terms).max()
