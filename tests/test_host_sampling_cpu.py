"""Host-side logic of the sampling / teacher / graph mirrors that needs no GPU: guidance interval rules
(reference tools/sampler.py:19-30), pass-through without labels, and the no-CPU-fallback contract of the new entries."""
import pytest
import torch

from vaw_b200 import _lib as L


def test_interval_cfg_rules():
    from vaw_b200.tools.sampler import IntervalCFG
    ident = lambda x, t, **k: x
    assert not IntervalCFG(ident, 10, guidance_scale=1.0)._use_cfg(300.0)            # scale 1: never
    assert IntervalCFG(ident, 10, guidance_scale=2.0)._use_cfg(300.0)                # default interval (-1, -1): always
    c = IntervalCFG(ident, 10, guidance_scale=2.0, interval=(100.0, 600.0))
    assert [c._use_cfg(v) for v in (99.9, 100.0, 300.0, 599.9, 600.0, 800.0)] == [False, True, True, True, False, False]
    assert IntervalCFG(ident, 10, 2.0, interval=(600.0, 100.0))._use_cfg(5.0)        # malformed interval: always
    assert c.null_label == 10
    t = torch.tensor(7.0)
    assert c._format_time(t, 3).shape == (3,) and c._format_time(torch.tensor([7.0]), 3).shape == (3,)
    assert c._format_time(torch.arange(3.0), 3).tolist() == [0.0, 1.0, 2.0]


def test_interval_cfg_passes_through_without_guidance():
    from vaw_b200.tools.sampler import IntervalCFG
    seen = []

    def model(x, t, **kw):
        seen.append((x.shape[0], kw.get("y")))
        return x * 2

    x = torch.ones(3, 2)
    c = IntervalCFG(model, 10, guidance_scale=2.0, interval=(100.0, 600.0))
    assert torch.equal(c(x, torch.full((3,), 800.0), y=torch.tensor([1, 2, 3])), x * 2)   # outside the interval
    assert torch.equal(c(x, torch.full((3,), 300.0)), x * 2)                               # no labels
    assert torch.equal(IntervalCFG(model, 10, 2.0, class_cond=False)(x, torch.zeros(3), y=torch.tensor([1, 2, 3])), x * 2)
    assert [n for n, _ in seen] == [3, 3, 3]
    with pytest.raises(AssertionError):
        c(x, torch.full((3,), 300.0), y=torch.tensor([1, 2]))                              # label batch mismatch
    with pytest.raises(L.VawError):                                                        # guided combine: CUDA only
        c(x, torch.full((3,), 300.0), y=torch.tensor([1, 2, 3]))


def test_new_entries_have_no_cpu_fallback():
    from vaw_b200.encoders.mocov3_vit import VisionTransformerMoCo, get_feature
    from vaw_b200.graph import GraphedTrainingLosses
    from vaw_b200.tools import gaussian_diffusion as gd
    from vaw_b200.tools.respace import SpacedDiffusion, space_timesteps
    from types import SimpleNamespace
    d = SpacedDiffusion(use_timesteps=space_timesteps(1000, "ddim10"), args=gd.default_args(),
                        betas=gd.get_named_beta_schedule("linear", 1000), model_mean_type=gd.ModelMeanType.EPSILON,
                        model_var_type=gd.ModelVarType.FIXED_LARGE, loss_type=gd.LossType.MSE, rescale_timesteps=True)
    x, t = torch.zeros(2, 3, 4, 4), torch.tensor([1, 2])
    for call in (d.p_sample, d.ddim_sample, d.ddim_reverse_sample, d.p_mean_variance):
        with pytest.raises(L.VawError):
            call(lambda a, b, **k: a, x, t)
    with pytest.raises(NotImplementedError):
        d.p_sample(lambda a, b, **k: a, x, t, denoised_fn=lambda v: v)
    m = VisionTransformerMoCo(img_size=32, patch_size=8, embed_dim=128, depth=1, num_heads=2)
    with pytest.raises(L.VawError):
        m.forward_features(torch.zeros(1, 3, 32, 32))
    with pytest.raises(NotImplementedError):
        get_feature(SimpleNamespace(enc_type="clip-vit-L"), torch.zeros(1, 3, 32, 32), m)
    with pytest.raises(L.VawError):
        GraphedTrainingLosses(d, torch.nn.Linear(2, 2), (2, 3, 4, 4))
