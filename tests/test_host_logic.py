"""CPU tier: host-side mirror of the reference interface (no kernels involved)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")

from vaw_b200.models import dit as vdit  # noqa: E402
from vaw_b200.tools import gaussian_diffusion as gd  # noqa: E402
from vaw_b200.tools import resample as rs  # noqa: E402


@pytest.mark.parametrize("sched", ["linear", "cosine", "linear_logsnr"])
def test_tables_bit_exact_with_reference(sched):
    g = np.load(os.path.join(G, "diffusion_golden.npz"))
    d = gd.create_gaussian_diffusion(noise_schedule=sched)
    assert np.array_equal(d.betas, g[f"betas_{sched}"])
    assert np.array_equal(d.sqrt_alphas_cumprod, g[f"sqrt_ac_{sched}"])
    assert np.array_equal(d.sqrt_one_minus_alphas_cumprod, g[f"sqrt_1mac_{sched}"])
    assert np.array_equal(d.posterior_mean_coef1, g[f"pmc1_{sched}"])
    assert np.array_equal(d.posterior_mean_coef2, g[f"pmc2_{sched}"])
    assert d.num_timesteps == 1000


def test_enum_values_match_reference_order():
    assert [m.name for m in gd.ModelMeanType] == ["PREVIOUS_X", "START_X", "EPSILON", "VELOCITY", "VECTOR", "SCORE"]
    assert gd.ModelMeanType.EPSILON.value == 3
    assert gd.LossType.KL.is_vb() and not gd.LossType.MSE.is_vb()


def test_compute_mse_loss_weight_helper_matches_golden():
    g = np.load(os.path.join(G, "diffusion_golden.npz"))
    d = gd.create_gaussian_diffusion(noise_schedule="linear")
    t = torch.arange(1000)
    a = gd._extract_into_tensor(d.sqrt_alphas_cumprod, t, t.shape)
    s = gd._extract_into_tensor(d.sqrt_one_minus_alphas_cumprod, t, t.shape)
    for mean, wt in (("EPSILON", "lambda"), ("EPSILON", "min_snr_5"), ("EPSILON", "debias"), ("START_X", "trunc_snr"),
                     ("VELOCITY", "min_snr_5"), ("VELOCITY", "lambda"), ("EPSILON", "constant")):
        w = gd.compute_mse_loss_weight(gd.ModelMeanType[mean], wt, t, a.clone(), s.clone())
        assert np.array_equal(w.float().numpy(), g[f"w_linear_{mean}_{wt}"]), (mean, wt)
    with pytest.raises(ValueError):
        gd.compute_mse_loss_weight(gd.ModelMeanType.VELOCITY, "debias", t, a, s)


def test_unknown_names_raise_like_the_reference():
    with pytest.raises(NotImplementedError):
        gd.get_named_beta_schedule("nope", 10)
    with pytest.raises(NotImplementedError):
        rs.create_named_schedule_sampler("nope", gd.create_gaussian_diffusion())
    d = gd.create_gaussian_diffusion(time_dist=["lognorm", 0, 1])
    with pytest.raises(NotImplementedError):
        d.sample_t(torch.zeros(2, 1))


def test_scale_timesteps_and_sampler_construction():
    d = gd.create_gaussian_diffusion()
    t = torch.tensor([0, 999])
    assert torch.equal(d._scale_timesteps(t), t.float())
    assert isinstance(rs.create_named_schedule_sampler("uniform", d), rs.UniformSampler)
    s = rs.create_named_schedule_sampler("loss-second-moment", d)
    assert s._loss_history.shape == (1000, 10) and s._loss_counts.shape == (1000,) and not s._warmed_up()
    assert np.array_equal(s.weights(), np.ones(1000))


def test_dit_state_dict_names_and_shapes_match_reference():
    """The golden fixture carries the reference DiT's state_dict; ours must have the same keys and shapes."""
    g = np.load(os.path.join(G, "dit_golden.npz"))
    ref = {k[len("param::"):]: g[k].shape for k in g.files if k.startswith("param::")}
    m = vdit.DiT(image_size=8, patch_size=2, in_channels=4, hidden_size=64, depth=2, num_heads=1,
                 class_dropout_prob=0.0, num_classes=10, learn_sigma=False, learn_align=True, encoder_depth=1,
                 z_dims=16, projector_dim=32)
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert set(mine) == set(ref), set(mine) ^ set(ref)
    for k in ref:
        assert mine[k] == tuple(ref[k]), k
    assert not m.pos_embed.requires_grad
    # reference init: adaLN-Zero and the output layer start at zero, pos_embed is the fixed sin-cos table
    assert float(m.blocks[0].adaLN_modulation[1].weight.abs().sum()) == 0.0
    assert float(m.final_layer.linear.weight.abs().sum()) == 0.0
    np.testing.assert_allclose(m.pos_embed.numpy(), g["param::pos_embed"], rtol=0, atol=1e-7)
    sd = {k[len("param::"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("param::")}
    m.load_state_dict(sd)  # strict


def test_dit_constructors_and_geometry():
    for name, (D, L, H) in {"DiT-S": (384, 12, 6), "DiT-B": (768, 12, 12), "DiT-L": (1024, 24, 16),
                            "DiT-XL": (1152, 28, 16)}.items():
        if name in ("DiT-L", "DiT-XL"):
            continue  # allocate only the small ones on the CPU tier
        m = vdit.DiT_models[name](image_size=32, patch_size=2, in_channels=4, class_dropout_prob=0.1, num_classes=1000,
                                  learn_sigma=False)
        assert (m.hidden_size, m.depth, m.num_heads) == (D, L, H)
        assert m.y_embedder.embedding_table.weight.shape == (1001, D)  # +1 row for the dropped label (dit.py:89-90)
    with pytest.raises(AssertionError):
        vdit.DiT(learn_align=True, encoder_depth=0, depth=1, hidden_size=64, num_heads=1)


def test_uvit_state_dict_names_and_shapes_match_reference():
    from vaw_b200.models import uvit as vuvit
    g = np.load(os.path.join(G, "uvit_golden.npz"))
    ref = {k[len("param::"):]: g[k].shape for k in g.files if k.startswith("param::")}
    m = vuvit.UViT(image_size=8, patch_size=2, in_channels=4, embed_dim=64, depth=3, num_heads=1, mlp_ratio=4,
                   num_classes=10, class_dropout_prob=0.0)
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert set(mine) == set(ref), set(mine) ^ set(ref)
    for k in ref:
        assert mine[k] == tuple(ref[k]), k
    m.load_state_dict({k: torch.from_numpy(g["param::" + k]) for k in ref})
    mm = vuvit.UViT_M(image_size=64, patch_size=4, in_channels=3, num_classes=1000, class_dropout_prob=0.0)
    assert (mm.embed_dim, mm.num_blocks, mm.num_heads, mm.pos_embed.shape[1]) == (768, 17, 12, 258)
    assert sum(p.numel() for p in mm.parameters()) == 130_940_979 or abs(sum(p.numel() for p in mm.parameters()) - 130.94e6) < 0.05e6
    with pytest.raises(NotImplementedError):
        vuvit.UViT(mlp_time_embed=True)
