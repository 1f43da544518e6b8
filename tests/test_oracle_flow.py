"""CPU tier: the flow-matching restatement (oracle/flow.py) against the fixture produced by EXECUTING the reference's
FlowMatching class (tests/golden/make_golden.py::flow_golden -> flow_golden.npz)."""
import os

import numpy as np
import pytest

from oracle import flow as oflow

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "flow_golden.npz"))
PATHS = ("linear", "cosine", "linear_logsnr")
CASES = [("START_X", "lambda"), ("EPSILON", "lambda"), ("EPSILON", "min_snr_5"), ("VELOCITY", "lambda"),
         ("VELOCITY", "min_snr_5"), ("VECTOR", "lambda"), ("VECTOR", "constant"), ("SCORE", "constant")]
# cos / sin / sigmoid come from different math libraries on the two sides (torch's vectorised CPU kernels vs libm)
ULP = dict(rtol=3e-7, atol=1e-7)


def _close(a, b, exact):
    if exact:
        assert np.array_equal(a, b)
    else:
        np.testing.assert_allclose(a, b, **ULP)


@pytest.mark.parametrize("path", PATHS)
def test_interpolant_qsample_target(path):
    t, x0, eps = G["t"], G["x0"], G["eps"]
    exact = path == "linear"
    _close(np.stack(oflow.interpolant(path, t)), G[f"interp::{path}"], exact)
    _close(oflow.q_sample(path, x0, eps, t), G[f"xt::{path}"], exact)
    for mean in ("START_X", "EPSILON", "VELOCITY", "VECTOR", "SCORE"):
        np.testing.assert_allclose(oflow.target(path, mean, x0, eps, t), G[f"target::{path}::{mean}"],
                                   rtol=0 if exact else 1e-6, atol=0 if exact else 1e-6)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("mean,wt", CASES)
def test_training_losses_and_gradient(path, mean, wt):
    mse, grad = oflow.mse_terms(path, mean, wt, G["x0"], G["eps"], G["t"], G["model_out"])
    np.testing.assert_allclose(mse, G[f"mse::{path}::{mean}::{wt}"], rtol=2e-6)
    # the fixture's gradient is d mean_n(loss_n) / d out = grad_n / N
    np.testing.assert_allclose(grad / len(G["t"]), G[f"grad::{path}::{mean}::{wt}"], rtol=2e-5, atol=1e-9)


@pytest.mark.parametrize("path", PATHS)
def test_model_output_conversions(path):
    mo, x, t = G["model_out"], G["x0"], G["t"]
    for mean in ("START_X", "EPSILON", "VELOCITY", "VECTOR", "SCORE"):
        if mean != "SCORE":
            np.testing.assert_allclose(oflow.to_vector(path, mean, mo, x, t), G[f"to_vector::{path}::{mean}"],
                                       rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(oflow.to_score(path, mean, mo, x, t), G[f"to_score::{path}::{mean}"],
                                   rtol=2e-5, atol=2e-4)


SDE_CASES = [("linear", "VECTOR"), ("linear", "VELOCITY"), ("linear", "START_X"), ("linear_logsnr", "VECTOR"),
             ("linear_logsnr", "VELOCITY"), ("linear_logsnr", "START_X"), ("linear_logsnr", "EPSILON"), ("cosine", "VECTOR")]


@pytest.mark.parametrize("path,mean", SDE_CASES)
@pytest.mark.parametrize("solver", ("euler", "heun"))
def test_sde_sample(path, mean, solver):
    toy = lambda x, t: (np.float32(0.25) * x - (np.float32(0.1) * t).astype(np.float32).reshape(-1, 1, 1, 1)).astype(np.float32)
    got = oflow.sde_sample(path, mean, toy, G["sde_start"], G["sde_noises"], 6, solver)
    want = G[f"sde::{path}::{mean}::{solver}"]
    if path == "cosine":
        # reference quirk: cos(fp32(pi/2)) < 0 -> sqrt of a negative diffusion coefficient at t = 1 -> NaN everywhere
        assert np.isnan(want).all() and np.isnan(got).all()
        return
    assert np.isfinite(want).all()
    if path == "linear" and mean == "VECTOR":
        assert np.array_equal(got, want)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)
