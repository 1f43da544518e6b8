"""GPU tier: the whole training step through the public API (sampler -> training_losses -> backward -> fused AdamW ->
sampler update) vs the oracle's CPU step on identical inputs and seeds."""
import numpy as np
import pytest
import torch

import __graft_entry__ as entry
from gpu_util import dezero, relerr
from oracle.train_step import OracleTrainer, synthetic_history
from vaw_b200.models.dit import DiT
from vaw_b200.optim import FusedAdamW
from vaw_b200.tools import gaussian_diffusion as gd
from vaw_b200.tools import resample as rs

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_smoke_entry():
    entry.smoke()


def test_three_steps_vs_oracle_trainer():
    torch.manual_seed(0)
    B, img = 8, 16
    m = DiT(image_size=img, patch_size=2, in_channels=4, hidden_size=128, depth=2, num_heads=2, class_dropout_prob=0.0,
            num_classes=1000).to(DEV).train()
    dezero(m)
    state = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    d = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda")
    s = rs.LossSecondMomentResampler(d)
    hist, counts = synthetic_history(0)
    s.load_history(hist, counts, DEV)
    opt = FusedAdamW(m, lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.01)

    import oracle.train_step as ots
    ots.DIT_CONFIGS["tiny"] = dict(hidden=128, depth=2, heads=2)
    ref = OracleTrainer("tiny", state=state, lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.01)
    ref.hist, ref.counts = hist.copy(), counts.copy()

    g = torch.Generator().manual_seed(5)
    for step in range(3):
        x = torch.randn(B, 4, img, img, generator=g); y = torch.randint(0, 1000, (B,), generator=g)
        eps = torch.randn(B, 4, img, img, generator=g)
        np.random.seed(100 + step)
        t, w = s.sample(B, DEV)
        terms = d.training_losses(m, x.to(DEV), None, t=t, model_kwargs={"y": y.to(DEV)}, noise=eps.to(DEV))
        s.update_with_local_losses(t, terms["loss"].detach())
        loss = (terms["loss"] * w).mean()
        loss.backward()
        opt.step(); opt.zero_grad()
        np.random.seed(100 + step)
        # the oracle records the GPU step's fp32 losses, so both sampler histories stay bit-identical and every
        # later draw of timestep indices can be compared exactly (the bf16 model's losses differ at the 1e-3 level)
        ref_loss, ref_terms, ref_t, ref_w = ref.step(x, y, noise=eps, history_losses=terms["loss"].detach().cpu().tolist())
        np.testing.assert_allclose(terms["loss"].detach().cpu().numpy(), ref_terms["loss"].detach().numpy(), rtol=2e-2)
        assert np.array_equal(t.cpu().numpy(), ref_t.numpy()), f"step {step}"   # timestep indices: bit-exact
        assert np.array_equal(w.cpu().numpy(), ref_w), f"step {step}"           # importance weights: bit-exact
        assert np.array_equal(s._loss_history, ref.hist), f"history diverged at step {step}"
        assert abs(loss.item() - ref_loss.item()) / abs(ref_loss.item()) < 2e-2
    # sampler history after three updates: bit-exact
    assert np.array_equal(s._loss_counts, ref.counts)
    assert np.array_equal(s._loss_history, ref.hist)
    # Parameters after three AdamW steps.  Adam turns every gradient into a +-lr step, so tensors whose true gradient
    # is ~0 (e.g. the key bias: softmax is shift invariant) move by rounding noise only; compare (a) weight matrices
    # tensor by tensor and (b) the update of the whole model as one vector.
    num = den = 0.0
    for k, p in m.named_parameters():
        if not p.requires_grad:
            continue
        got, want, init = p.detach().cpu(), ref.sd[k].detach(), state[k]
        if p.dim() >= 2:
            assert relerr(got, want) < 2e-2, k
        num += float(((got - want) ** 2).sum())
        den += float(((want - init) ** 2).sum())
    assert (num / den) ** 0.5 < 0.15, (num / den) ** 0.5


def test_fused_adamw_clip_and_ema_vs_torch():
    """FusedAdamW(step(max_grad_norm), ema_decay) against torch.optim.AdamW + clip_grad_norm_ + the reference's ema()
    (tools/trainer.py:12-18,60-62) on the same gradients: same update, same EMA, no host sync in the fused path."""
    import copy
    from vaw_b200.models.dit import DiT
    from vaw_b200.optim import FusedAdamW
    torch.manual_seed(0)
    m = DiT(image_size=16, patch_size=2, in_channels=4, hidden_size=128, depth=2, num_heads=2, class_dropout_prob=0.0,
            num_classes=10).to("cuda")
    dezero(m)
    x = torch.randn(4, 4, 16, 16, device="cuda"); t = torch.rand(4, device="cuda") * 999
    y = torch.randint(0, 10, (4,), device="cuda")
    opt = FusedAdamW(m, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.01, ema_decay=0.99)
    ref_p = {k: p.detach().clone() for k, p in m.named_parameters()}
    ref_params = [torch.nn.Parameter(v.clone()) for v in ref_p.values()]
    ref_opt = torch.optim.AdamW(ref_params, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.01, eps=1e-8)
    ema_ref = [v.clone() for v in ref_p.values()]
    for it in range(3):
        out, _ = m(x, t, y)
        (out.float() ** 2).sum().backward()          # large gradients: the clip is active
        for rp, p in zip(ref_params, m.parameters()):   # frozen tensors (pos_embed) have no gradient on either side
            rp.grad = None if p.grad is None else p.grad.detach().clone()
        total = torch.nn.utils.clip_grad_norm_(ref_params, 0.5)
        ref_opt.step()
        with torch.no_grad():
            for e, rp in zip(ema_ref, ref_params):
                e.copy_(e * 0.99 + rp.detach() * 0.01)
        opt.step(max_grad_norm=0.5)
        opt.zero_grad()
        norm, coef = opt.grad_norm.tolist()
        assert abs(norm - float(total)) <= 1e-4 * float(total)
        assert coef < 1.0 and abs(coef - 0.5 / (float(total) + 1e-6)) < 1e-5
    ema_sd = opt.ema_state_dict()
    for (k, p), rp, e in zip(m.named_parameters(), ref_params, ema_ref):
        torch.testing.assert_close(p.detach(), rp.detach(), rtol=2e-5, atol=2e-6, msg=k)
        torch.testing.assert_close(ema_sd[k], e, rtol=2e-5, atol=2e-6, msg=k)
