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


def test_reverse_table_rows_follow_the_header_enum():
    """The [VAW_RT_ROWS, T] table the host uploads must be laid out in the order include/vaw_b200.h declares."""
    import os
    import re
    import numpy as np
    from vaw_b200.tools import gaussian_diffusion as gd
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "vaw_b200.h")).read()
    body = re.search(r"enum \{ (VAW_RT_SQRT_RECIP_AC = 0.*?VAW_RT_ROWS) \};", text, re.S).group(1)
    names = [n.split("=")[0].strip() for n in re.sub(r"/\*.*?\*/", "", body, flags=re.S).split(",")]
    assert names[-1] == "VAW_RT_ROWS" and len(names) == 16
    d = gd.create_gaussian_diffusion(noise_schedule="linear", var_type="fixed_small")
    want = {
        "VAW_RT_SQRT_RECIP_AC": d.sqrt_recip_alphas_cumprod, "VAW_RT_SQRT_RECIPM1_AC": d.sqrt_recipm1_alphas_cumprod,
        "VAW_RT_SQRT_AC": d.sqrt_alphas_cumprod, "VAW_RT_SQRT_1MAC": d.sqrt_one_minus_alphas_cumprod,
        "VAW_RT_INV_COEF1": 1.0 / d.posterior_mean_coef1,
        "VAW_RT_COEF2_OVER_COEF1": d.posterior_mean_coef2 / d.posterior_mean_coef1,
        "VAW_RT_COEF1": d.posterior_mean_coef1, "VAW_RT_COEF2": d.posterior_mean_coef2,
        "VAW_RT_LOGVAR": d.posterior_log_variance_clipped, "VAW_RT_MAX_LOG": np.log(d.betas),
        "VAW_RT_VARIANCE": d.posterior_variance, "VAW_RT_AC": d.alphas_cumprod, "VAW_RT_AC_PREV": d.alphas_cumprod_prev,
        "VAW_RT_AC_NEXT": d.alphas_cumprod_next, "VAW_RT_TRUE_LOGVAR": d.posterior_log_variance_clipped,
    }
    tab = d._reverse_table("cpu").numpy()
    assert tab.shape == (15, 1000) and tab.dtype == np.float32
    for i, n in enumerate(names[:-1]):
        np.testing.assert_array_equal(tab[i], np.asarray(want[n], dtype=np.float64).astype(np.float32), err_msg=n)
    # FIXED_LARGE swaps in the beta-based variance rows (reference :326-331)
    dl = gd.create_gaussian_diffusion(noise_schedule="linear", var_type="fixed_large")
    var = np.append(dl.posterior_variance[1], dl.betas[1:])
    tl = dl._reverse_table("cpu").numpy()
    np.testing.assert_array_equal(tl[names.index("VAW_RT_VARIANCE")], var.astype(np.float32))
    np.testing.assert_array_equal(tl[names.index("VAW_RT_LOGVAR")], np.log(var).astype(np.float32))
    np.testing.assert_array_equal(tl[names.index("VAW_RT_TRUE_LOGVAR")], dl.posterior_log_variance_clipped.astype(np.float32))
