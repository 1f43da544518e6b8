"""Host-side respacing logic (vaw_b200.tools.respace, oracle.diffusion.space_timesteps / spaced_betas) against the
fixture tests/golden/reverse_golden.npz written from the executed reference (tools/respace.py:9-128).  Integer and
float64 host work: bit-exact."""
import os

import numpy as np
import pytest

from oracle import diffusion as odiff

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [(1000, "ddim10"), (1000, "ddim50"), (1000, "ddim250"), (300, "10,15,20"), (1000, [1000]), (1000, "250"),
         (100, "3,1,7"), (1000, [1, 1, 1]), (999, "37,2")]


@pytest.fixture(scope="module")
def rg():
    return np.load(os.path.join(G, "reverse_golden.npz"))


@pytest.mark.parametrize("n,spec", CASES)
def test_space_timesteps(rg, n, spec):
    from vaw_b200.tools.respace import space_timesteps
    want = rg[f"space::{n}::{spec}"].tolist()
    assert sorted(space_timesteps(n, spec)) == want
    assert odiff.space_timesteps(n, spec) == want


def test_space_timesteps_errors():
    from vaw_b200.tools.respace import space_timesteps
    with pytest.raises(ValueError):
        space_timesteps(1000, "ddim999")     # no integer stride
    with pytest.raises(ValueError):
        space_timesteps(10, "6,6")           # a section of 5 cannot hold 6 steps


@pytest.mark.parametrize("spec", ["ddim10", "ddim50", "10,15,20"])
def test_spaced_diffusion_tables(rg, spec):
    from vaw_b200.tools import gaussian_diffusion as gd
    from vaw_b200.tools.respace import SpacedDiffusion, space_timesteps, _WrappedModel
    d = SpacedDiffusion(use_timesteps=space_timesteps(1000, spec), args=gd.default_args(),
                        betas=gd.get_named_beta_schedule("cosine", 1000), model_mean_type=gd.ModelMeanType.EPSILON,
                        model_var_type=gd.ModelVarType.FIXED_LARGE, loss_type=gd.LossType.MSE, rescale_timesteps=True)
    np.testing.assert_array_equal(d.betas, rg[f"spaced_betas::{spec}"])
    assert d.timestep_map == rg[f"spaced_map::{spec}"].tolist()
    assert d.num_timesteps == len(d.timestep_map) and d.original_num_steps == 1000
    np.testing.assert_array_equal(
        odiff.spaced_betas(odiff.named_beta_schedule("cosine", 1000), d.timestep_map), rg[f"spaced_betas::{spec}"])
    # the wrapped model sees the base process's timestep, rescaled (reference :124-128)
    import torch
    seen = []
    w = d._wrap_model(lambda x, ts, **k: seen.append(ts))
    assert isinstance(w, _WrappedModel) and d._wrap_model(w) is w
    w(None, torch.tensor([0, 1, 2, 3, 5, 9]))
    np.testing.assert_array_equal(seen[0].numpy(), rg[f"spaced_model_t::{spec}"])
    assert d._scale_timesteps(7) == 7
