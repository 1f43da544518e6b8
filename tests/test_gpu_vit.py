"""GPU tier: the frozen MoCo-v3 ViT teacher of the REPA loss (vaw_b200.encoders.mocov3_vit) against the
reference fixture (tests/golden/vit_golden.npz) and the oracle (oracle/vit.py).  bf16 tensor-core path: 2e-2 rel-L2
against the oracle under bf16 autocast; the fused preprocess + patchify is checked bit-exactly."""
import os
import sys

import numpy as np
import pytest
import torch

from gpu_util import relerr
from oracle import vit as ovit

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, G)
from fill import fill_by_name   # noqa: E402

TOL = 2e-2


def tiny_model():
    from vaw_b200.encoders.mocov3_vit import VisionTransformerMoCo
    m = VisionTransformerMoCo(img_size=32, patch_size=8, embed_dim=128, depth=2, num_heads=2, mlp_ratio=4)
    pos = m.pos_embed.detach().clone()
    fill_by_name(m)
    with torch.no_grad():
        m.pos_embed.copy_(pos)
    return m.to(DEV).eval()


def test_patchify_norm_bit_exact():
    from vaw_b200 import _lib as L
    from vaw_b200.encoders import mocov3_vit as mv   # registers the signatures
    torch.manual_seed(0)
    B, P, H = 3, 8, 32
    raw = torch.randint(0, 256, (B, 3, H, H), device=DEV).float()
    patches = torch.empty(B * (H // P) ** 2, 3 * P * P, dtype=torch.bfloat16, device=DEV)
    mean = torch.tensor(mv.IMAGENET_DEFAULT_MEAN, device=DEV)
    std = torch.tensor(mv.IMAGENET_DEFAULT_STD, device=DEV)
    L.call("vaw_patchify_norm", raw.data_ptr(), mean.data_ptr(), std.data_ptr(), patches.data_ptr(), B, 3, H, H, P,
           L.stream_ptr())
    pre = ovit.preprocess_raw_image(raw.cpu())                                           # [B, 3, H, W] fp32
    want = pre.unfold(2, P, P).unfold(3, P, P).permute(0, 2, 3, 1, 4, 5).reshape(patches.shape).bfloat16()
    assert torch.equal(patches.cpu(), want)


def test_tiny_teacher_matches_reference_fixture():
    from types import SimpleNamespace
    from vaw_b200.encoders.mocov3_vit import get_feature
    vg = np.load(os.path.join(G, "vit_golden.npz"))
    m = tiny_model()
    raw = torch.from_numpy(vg["raw"]).to(DEV)
    f = get_feature(SimpleNamespace(enc_type="mocov3-vit-b"), raw, m)
    assert f.shape == (3, 16, 128) and f.dtype == torch.bfloat16
    assert relerr(f, torch.from_numpy(vg["features"]).to(DEV)) < TOL
    # already-normalised input through forward_features, cls token kept
    full = m.forward_features(torch.from_numpy(vg["pre"]).to(DEV))
    assert full.shape == (3, 17, 128)
    assert relerr(full[:, 1:], torch.from_numpy(vg["features"]).to(DEV)) < TOL
    with pytest.raises(NotImplementedError):
        get_feature(SimpleNamespace(enc_type="dinov2-vit-b"), raw, m)


@pytest.mark.parametrize("B", [1, 5])
def test_vit_base_vs_oracle(B):
    """The shape the REPA recipe uses: ViT-B/16 at 256 px (257 tokens, 12 x 64 heads)."""
    from vaw_b200.encoders.mocov3_vit import vit_base
    from vaw_b200 import _lib as L
    torch.manual_seed(2)
    m = vit_base().to(DEV).eval()
    with torch.no_grad():
        for k, p in m.named_parameters():
            if k.endswith("bias") or k == "cls_token":
                p.normal_(0, 0.05)
    raw = torch.randint(0, 256, (B, 3, 256, 256), device=DEV).float()
    n0 = L.launch_count()
    f = m.forward_features(raw, raw_pixels=True)[:, 1:]
    assert L.launch_count() - n0 >= 3 + 7 * 12 + 1
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    with torch.no_grad():
        f32 = ovit.get_feature(sd, raw, patch_size=16, num_heads=12, depth=12)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            fbf = ovit.get_feature(sd, raw, patch_size=16, num_heads=12, depth=12)
    assert f.shape == (B, 256, 768)
    assert relerr(f, fbf) < TOL
    assert relerr(f, f32) < 1.5 * relerr(fbf, f32) + 1e-3    # no further from fp32 than torch's own bf16 path
    # the bf16 weight shadows follow parameter updates (load_state_dict of a checkpoint after construction)
    with torch.no_grad():
        m.norm.weight.mul_(2.0)
        m.blocks[0].mlp.fc2.weight.mul_(0.5)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        fbf2 = ovit.get_feature(sd, raw, patch_size=16, num_heads=12, depth=12)
    assert relerr(m.forward_features(raw, raw_pixels=True)[:, 1:], fbf2) < TOL


def test_teacher_feeds_the_alignment_loss():
    """Config 5 wiring (trainer.py:55-58): teacher features -> training_losses(model, x, features)."""
    from types import SimpleNamespace
    from vaw_b200.encoders.mocov3_vit import VisionTransformerMoCo, get_feature
    from vaw_b200.models.dit import DiT
    from vaw_b200.tools import gaussian_diffusion as gd
    from vaw_b200 import _lib as L
    torch.manual_seed(4)
    teacher = VisionTransformerMoCo(img_size=64, patch_size=16, embed_dim=128, depth=1, num_heads=2).to(DEV).eval()
    student = DiT(image_size=8, patch_size=2, in_channels=4, hidden_size=128, depth=2, num_heads=2,
                  class_dropout_prob=0.0, num_classes=10, learn_align=True, encoder_depth=1, z_dims=128,
                  projector_dim=64).to(DEV).train()
    d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda",
                                     learn_align=True, gamma=0.5)
    pixels = torch.randint(0, 256, (4, 3, 64, 64), device=DEV).float()
    feats = get_feature(SimpleNamespace(enc_type="mocov3-vit-s"), pixels, teacher)      # [4, 16, 128]
    x0 = torch.randn(4, 4, 8, 8, device=DEV)
    terms = d.training_losses(student, x0, feats, model_kwargs={"y": torch.randint(0, 10, (4,), device=DEV)})
    terms["loss"].mean().backward()
    assert torch.isfinite(terms["align"]) and terms["align"].item() > 0
    with pytest.raises(L.VawError):
        teacher.forward_features(pixels.cpu())
