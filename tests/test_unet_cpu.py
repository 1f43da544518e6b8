"""CPU tier: the UNet ("API only") against fixtures produced by executing the reference's UNet.

 * state-dict names and shapes of every factory in UNet_models equal the reference's (checkpoint compatibility);
 * forward, per-sample loss and parameter gradients of two tiny UNets (covering scale-shift/additive conditioning,
   ResBlock vs convolutional resampling, both attention orders, conditional/unconditional) equal the reference's to
   fp32 round-off.  The diffusion maths on CPU comes from oracle/ (the GPU tier runs the same case through K1/K2).
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, G)
from fill import fill_by_name, grad_digest  # noqa: E402

from oracle import diffusion as od  # noqa: E402
from vaw_b200.models import unet as vunet  # noqa: E402

CASES = {
    "a": dict(image_size=8, num_channels=32, num_res_blocks=1, channel_mult="1,2", in_channels=3, num_classes=10,
              class_cond=True, attention_resolutions="4", num_heads=2, use_scale_shift_norm=True, resblock_updown=True,
              use_new_attention_order=True),
    "b": dict(image_size=8, num_channels=32, num_res_blocks=1, channel_mult="1,2", in_channels=3, num_classes=10,
              class_cond=False, attention_resolutions="4,8", num_heads=1, num_head_channels=16,
              use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False),
}


@pytest.mark.parametrize("name", ["UNet-32", "ADM-32", "ADM-64", "ADM-128", "ADM-256", "ADM-512", "UNet-64", "LDM"])
def test_state_dict_layout_matches_reference(name):
    with open(os.path.join(G, "unet_shapes.json")) as f:
        want = json.load(f)[name]
    with torch.device("meta"):
        m = vunet.UNet_models[name]()
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert list(got) == list(want)
    assert got == want


def test_zero_init_and_label_dropout_table():
    m = vunet.create_unet_model(**CASES["a"], drop_label_prob=0.1)
    assert m.label_emb.num_embeddings == 11
    assert all(float(p.abs().sum()) == 0 for p in m.out[2].parameters())
    assert all(float(p.abs().sum()) == 0 for p in m.input_blocks[1][0].out_layers[3].parameters())
    assert all(float(p.abs().sum()) == 0 for p in m.middle_block[1].proj_out.parameters())
    y = torch.tensor([1, 2, 3])
    assert m.token_drop(y, torch.tensor([1, 0, 1])).tolist() == [10, 2, 10]
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 3, 8, 8), torch.zeros(1))           # class-conditional model needs y
    with pytest.raises(ValueError):
        vunet.create_unet_model(image_size=48, num_channels=32, num_res_blocks=1)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_forward_loss_grads_match_reference(tag):
    g = np.load(os.path.join(G, "unet_golden.npz"))
    cfg = CASES[tag]
    m = fill_by_name(vunet.create_unet_model(**cfg)).train()
    x0, eps, t, y = (torch.from_numpy(g[f"{tag}::{k}"]) for k in ("x0", "eps", "t", "y"))
    kw = {"y": y} if cfg["class_cond"] else {}
    tb = od.tables(od.named_beta_schedule("linear", 1000))
    x_t = torch.from_numpy(od.q_sample(tb, x0.numpy(), t.numpy(), eps.numpy()))
    out = m(x_t, t.float(), **kw)
    np.testing.assert_allclose(out.detach().numpy(), g[f"{tag}::fwd_out"], rtol=1e-4, atol=2e-5)
    terms = od.training_losses_torch(tb, "EPSILON", "min_snr_5", lambda x, ts: m(x, ts, **kw), x0, t, eps)
    terms["loss"].mean().backward()
    np.testing.assert_allclose(terms["mse"].detach().numpy(), g[f"{tag}::mse"], rtol=1e-5)
    n = 0
    for k, p in m.named_parameters():
        want = g[f"{tag}::grad::{k}"]
        got = grad_digest(p.grad)
        scale = max(float(np.abs(want).max()), 1e-6)
        assert np.abs(got - want).max() <= 2e-4 * scale + 1e-6, (k, np.abs(got - want).max(), scale)
        n += 1
    assert n == sum(1 for k in g.files if k.startswith(f"{tag}::grad::"))
