"""Pins oracle/vit.py (MoCo-v3 ViT teacher restatement) and the host side of vaw_b200.encoders.mocov3_vit against
tests/golden/vit_golden.npz, written by executing the reference's encoders/mocov3_vit.py + tools/align_utils.py over the
timm restatement in oracle/ref_stubs (timm itself is absent: see the header of oracle/vit.py)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import vit as ovit

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, G)
from fill import fill_by_name   # noqa: E402


@pytest.fixture(scope="module")
def vg():
    return np.load(os.path.join(G, "vit_golden.npz"))


def tiny_model():
    from vaw_b200.encoders.mocov3_vit import VisionTransformerMoCo
    m = VisionTransformerMoCo(img_size=32, patch_size=8, embed_dim=128, depth=2, num_heads=2, mlp_ratio=4)
    pos = m.pos_embed.detach().clone()
    fill_by_name(m)
    with torch.no_grad():
        m.pos_embed.copy_(pos)
    return m


def test_preprocess_bit_exact(vg):
    np.testing.assert_array_equal(ovit.preprocess_raw_image(torch.from_numpy(vg["raw"])).numpy(), vg["pre"])


def test_pos_embed_and_names(vg):
    from vaw_b200.encoders.mocov3_vit import sincos_2d, vit_base
    np.testing.assert_array_equal(ovit.sincos_pos_embed(4, 4, 128).numpy(), vg["pos_embed"])
    np.testing.assert_array_equal(sincos_2d(4, 4, 128).numpy(), vg["pos_embed"])
    np.testing.assert_array_equal(sincos_2d(16, 16, 768).numpy()[:, ::8, ::16], vg["pos_embed_base_sub"])
    m = tiny_model()
    assert sorted(m.state_dict()) == vg["names"].tolist()
    with torch.device("meta"):
        big = vit_base()
    sd = big.state_dict()
    assert sorted(sd) == vg["base_names"].tolist()
    assert [str(tuple(sd[k].shape)) for k in sorted(sd)] == vg["base_shapes"].tolist()
    assert not any(p.requires_grad for p in m.parameters())


def test_oracle_forward_matches_reference(vg):
    m = tiny_model()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    f = ovit.get_feature(sd, torch.from_numpy(vg["raw"]), patch_size=8, num_heads=2, depth=2)
    np.testing.assert_allclose(f.numpy(), vg["features"], rtol=1e-5, atol=1e-5)   # fp32 reductions
