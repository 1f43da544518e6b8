"""CUDA-graph replay of K1 -> DiT forward -> K2 -> backward (vaw_b200.graph) must reproduce the eager step bit for bit,
for inputs that change between replays, and leave the gradients where the optimizer expects them."""
import pytest
import torch

from gpu_util import dezero

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("align", [False, True])
def test_graph_replay_equals_eager(align):
    from vaw_b200.graph import GraphedTrainingLosses
    from vaw_b200.models.dit import DiT
    from vaw_b200.optim import FusedAdamW
    from vaw_b200.tools import gaussian_diffusion as gd
    torch.manual_seed(0)
    B = 8
    net = DiT(image_size=16, patch_size=2, in_channels=4, hidden_size=128, depth=2, num_heads=2,
              class_dropout_prob=0.0, num_classes=10, learn_align=align, encoder_depth=1, z_dims=48,
              projector_dim=64).to(DEV).train()
    dezero(net)
    d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda",
                                     learn_align=align, gamma=0.5)
    g = GraphedTrainingLosses(d, net, (B, 4, 16, 16), feature_shape=(B, 64, 48) if align else None)
    opt = FusedAdamW(net, lr=1e-3, betas=(0.9, 0.95))
    for step in range(3):
        x0 = torch.randn(B, 4, 16, 16, device=DEV)
        y = torch.randint(0, 10, (B,), device=DEV)
        t = torch.randint(0, 1000, (B,), device=DEV)
        w = torch.rand(B, device=DEV) + 0.5
        eps = torch.randn_like(x0)
        feats = torch.randn(B, 64, 48, device=DEV) if align else None
        terms = g(x0, t, w, y=y, features=feats, noise=eps)
        loss_g, grad_g = terms["loss"].clone(), net._gflat.clone()
        assert all(p.grad is not None for p in net.parameters() if p.requires_grad)
        for p in net.parameters():
            p.grad = None
        ref = d.training_losses(net, x0, feats, t=t, model_kwargs={"y": y}, noise=eps)
        (ref["loss"] * w).mean().backward()
        assert torch.equal(loss_g, ref["loss"].detach())
        assert torch.equal(grad_g, net._gflat)
        opt.step()            # the weights move between replays: the graph reads the live parameter buffers
        opt.zero_grad()
    # noise drawn inside the call when none is given: different draws, finite results
    a = g(x0, t, w, y=y, features=feats)["loss"].clone()
    b = g(x0, t, w, y=y, features=feats)["loss"].clone()
    assert torch.isfinite(a).all() and not torch.equal(a, b)
    with pytest.raises(ValueError):
        g(x0, t, w)


def test_graph_with_torch_adamw_and_accumulation():
    """Weights updated OUTSIDE FusedAdamW (torch.optim.AdamW over the same Parameters) must reach the replayed GEMMs (the
    bf16 shadow is refreshed before every replay), and `accumulate=True` adds to the gradients instead of replacing."""
    from vaw_b200.graph import GraphedTrainingLosses
    from vaw_b200.models.dit import DiT
    from vaw_b200.tools import gaussian_diffusion as gd
    torch.manual_seed(1)
    B = 8
    net = DiT(image_size=16, patch_size=2, in_channels=4, hidden_size=128, depth=2, num_heads=2,
              class_dropout_prob=0.0, num_classes=10).to(DEV).train()
    dezero(net)
    d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda")
    g = GraphedTrainingLosses(d, net, (B, 4, 16, 16))
    opt = torch.optim.AdamW([p for p in net.parameters() if p.requires_grad], lr=1e-2)
    x0 = torch.randn(B, 4, 16, 16, device=DEV); y = torch.randint(0, 10, (B,), device=DEV)
    t = torch.randint(0, 1000, (B,), device=DEV); eps = torch.randn_like(x0)
    losses = []
    for step in range(3):
        terms = g(x0, t, None, y=y, noise=eps)
        losses.append(terms["loss"].clone())
        for p in net.parameters():
            p.grad = None
        ref = d.training_losses(net, x0, None, t=t, model_kwargs={"y": y}, noise=eps)
        ref["loss"].mean().backward()
        assert torch.equal(losses[-1], ref["loss"].detach()), step     # replay saw the weights torch's AdamW wrote
        opt.step()
    assert not torch.equal(losses[0], losses[2])
    for p in net.parameters():
        p.grad = None
    g(x0, t, None, y=y, noise=eps)
    g1 = net._gflat.clone()
    g(x0, t, None, y=y, noise=eps, accumulate=True)
    assert torch.allclose(net._gflat, 2 * g1, rtol=1e-6, atol=1e-9)
    g(x0, t, None, y=y, noise=eps)                                     # overwrite mode again
    assert torch.equal(net._gflat, g1)
