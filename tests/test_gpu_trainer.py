"""GPU tier: the reference's Trainer.train_step (tools/trainer.py:68-150), restated call for call, driving THIS repo's
objects the way main.py:341-355 builds them:

    model      = DiT(...)                         ; ema_model = copy.deepcopy(model)              main.py:343-344
    model      = DataParallel(model, device_ids=) ; ema_model = DataParallel(ema_model, ...)      main.py:347-348 (DDP)
    optimizer  = FusedAdamW(model, lr, betas, weight_decay, eps)                                  main.py:354
    scheduler  = LambdaLR(optimizer, lr_lambda)                                                   main.py:355
    scaler     = GradScaler()                                                                     trainer.py:40

and then per step: `with sync_context: with autocast(): loss_dict = training_losses(model, images, features,
model_kwargs=...); loss = loss_dict["loss"].mean() / accum; scaler.scale(loss).backward()` ... `scaler.unscale_`,
`clip_grad_norm_(model.parameters())`, `scaler.step`, `scaler.update`, `optimizer.zero_grad`, `scheduler.step`,
`ema(model, ema_model, decay)` over the two state_dicts.  The reference cannot travel to the GPU box, so its loop body is
written out here (each line cites trainer.py); the checker is the oracle model + torch.optim.AdamW + LambdaLR run
through the identical sequence."""
import copy
import warnings
from contextlib import nullcontext

import numpy as np
import pytest
import torch
import torch.nn as nn

from gpu_util import dezero, relerr
from oracle import diffusion as odiff
from oracle.dit import dit_forward
from vaw_b200.models.dit import DiT
from vaw_b200.optim import DataParallel, FusedAdamW
from vaw_b200.tools import gaussian_diffusion as gd

pytestmark = pytest.mark.gpu
DEV = "cuda"


def ema(source, target, decay):
    """tools/trainer.py:12-18, verbatim semantics: iterate the state_dicts, write through `.data`."""
    with torch.no_grad():
        source_dict, target_dict = source.state_dict(), target.state_dict()
        for key in source_dict.keys():
            target_dict[key].data.copy_(target_dict[key].data * decay + source_dict[key].data * (1 - decay))


class RefShapedTrainer:
    """The body of tools/trainer.py:28-150 with data loading replaced by a list of micro-batches."""

    def __init__(self, args, model, ema_model, optimizer, scheduler, diffusion):
        self.args, self.model, self.ema_model = args, model, ema_model
        self.optimizer, self.scheduler, self.diffusion = optimizer, scheduler, diffusion
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from torch.cuda.amp import GradScaler                       # trainer.py:3
            self.scaler = GradScaler() if args.amp else None            # trainer.py:40

    def _compute_loss(self, images, labels, features):                  # trainer.py:55-58
        model_kwargs = {"y": labels} if self.args.class_cond else {}
        return self.diffusion.training_losses(self.model, images, features, model_kwargs=model_kwargs)

    def train_step(self, micro_batches):
        from torch.cuda.amp import autocast
        self.model.train()                                              # :69
        accum = max(1, self.args.grad_accumulation)                     # :73
        total = 0.0
        for k in range(accum):                                          # :79
            images, labels = micro_batches[k]
            if self.args.parallel and accum > 1 and k < accum - 1:      # :94-99
                sync_context = self.model.no_sync()
            else:
                sync_context = nullcontext()
            with sync_context:                                          # :103
                if self.args.amp:
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        with autocast():                                # :105
                            loss_dict = self._compute_loss(images, labels, None)
                            loss = loss_dict["loss"].mean() / accum     # :107
                    self.scaler.scale(loss).backward()                  # :108
                else:
                    loss_dict = self._compute_loss(images, labels, None)
                    loss = loss_dict["loss"].mean() / accum
                    loss.backward()                                     # :112
            total += loss.item()                                        # :114
            if (k + 1) % accum == 0:                                    # :123
                if self.args.amp:
                    if self.args.grad_clip:
                        self.scaler.unscale_(self.optimizer)            # :126
                        nn.utils.clip_grad_norm_(self.model.parameters(), self.args.grad_clip)   # :60-62
                    self.scaler.step(self.optimizer)                    # :128
                    self.scaler.update()                                # :129
                else:
                    if self.args.grad_clip:
                        nn.utils.clip_grad_norm_(self.model.parameters(), self.args.grad_clip)
                    self.optimizer.step()                               # :132
                self.optimizer.zero_grad()                              # :133
        self.scheduler.step()                                           # :135
        ema(self.model, self.ema_model, self.args.ema_decay)            # :137-138 (rank 0)
        return total


class OracleRun:
    """Same sequence with the oracle DiT (bf16 autocast), torch.optim.AdamW and LambdaLR - the reference's own stack."""

    def __init__(self, state, cfg, lr, betas, wd, lam, clip, decay):
        self.sd = {k: v.detach().clone().float().requires_grad_(k != "pos_embed") for k, v in state.items()}
        self.ema = {k: v.detach().clone() for k, v in self.sd.items()}
        self.cfg, self.clip, self.decay = cfg, clip, decay
        self.params = [v for v in self.sd.values() if v.requires_grad]
        self.opt = torch.optim.AdamW(self.params, lr=lr, betas=betas, weight_decay=wd, eps=1e-8)
        self.sched = torch.optim.lr_scheduler.LambdaLR(self.opt, lam)
        self.tb = odiff.tables(odiff.named_beta_schedule("cosine", 1000))

    def step(self, micro, ts, noises):
        accum, total = len(micro), 0.0
        for (x, y), t, eps in zip(micro, ts, noises):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                fn = lambda xt, tt: dit_forward(self.sd, xt.to(DEV), tt.to(DEV), y, **self.cfg)[0].float().cpu()
                terms = odiff.training_losses_torch(self.tb, "EPSILON", "lambda", fn, x.cpu(), t.cpu(), eps.cpu())
            loss = terms["loss"].mean() / accum
            loss.backward()
            total += loss.item()
        if self.clip:
            nn.utils.clip_grad_norm_(self.params, self.clip)
        self.opt.step()
        self.opt.zero_grad()
        self.sched.step()
        with torch.no_grad():
            for k in self.sd:
                self.ema[k].copy_(self.ema[k] * self.decay + self.sd[k].detach() * (1 - self.decay))
        return total


@pytest.mark.parametrize("amp,wrap,accum", [(True, True, 2), (False, False, 1), (True, False, 1)])
def test_trainer_shaped_steps_vs_oracle(amp, wrap, accum):
    from types import SimpleNamespace
    torch.manual_seed(0)
    B, img, steps = 4, 16, 3
    cfg = dict(patch_size=2, num_heads=2, depth=2)
    model = DiT(image_size=img, patch_size=2, in_channels=4, hidden_size=128, depth=2, num_heads=2,
                class_dropout_prob=0.0, num_classes=10).to(DEV)
    dezero(model)
    init = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ema_model = copy.deepcopy(model).to(DEV)                                    # main.py:344
    if wrap:   # main.py:347-348 with DDP -> DataParallel (single process here: no collective, same code path otherwise)
        model = DataParallel(model, device_ids=[0], output_device=0)
        ema_model = DataParallel(ema_model, device_ids=[0], output_device=0)
    lr, betas, wd, clip, decay = 2e-3, (0.9, 0.95), 0.01, 0.5, 0.9
    lam = lambda s: 0.5 ** s
    optimizer = FusedAdamW(model, lr=lr, betas=betas, weight_decay=wd, eps=1e-8)          # main.py:354
    scheduler = torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda=lam)               # main.py:355
    diffusion = gd.create_gaussian_diffusion(noise_schedule="cosine", mean_type="epsilon", weight_type="lambda")
    args = SimpleNamespace(amp=amp, grad_accumulation=accum, grad_clip=clip, parallel=wrap, class_cond=True,
                           ema_decay=decay)
    tr = RefShapedTrainer(args, model, ema_model, optimizer, scheduler, diffusion)
    ref = OracleRun(init, cfg, lr, betas, wd, lam, clip, decay)

    g = torch.Generator(device=DEV).manual_seed(3)
    for step in range(steps):
        micro = [(torch.randn(B, 4, img, img, device=DEV, generator=g), torch.randint(0, 10, (B,), device=DEV, generator=g))
                 for _ in range(accum)]
        # the trainer draws noise and t inside training_losses (t=None, noise=None): replay the device generator for the
        # oracle in the reference's order, noise first, then t (gaussian_diffusion.py:849-852)
        state = torch.cuda.get_rng_state()
        ts, noises = [], []
        for x, _ in micro:
            noises.append(torch.randn_like(x))
            ts.append(torch.randint(0, 1000, (B,), device=DEV))
        torch.cuda.set_rng_state(state)
        got = tr.train_step(micro)
        want = ref.step(micro, ts, noises)
        assert abs(got - want) / abs(want) < 2e-2, (step, got, want)
        assert abs(optimizer.param_groups[0]["lr"] - lr * lam(step + 1)) < 1e-12          # LambdaLR drives the fused step
    if amp:
        assert tr.scaler.get_scale() == 65536.0                                           # no overflow, no back-off
    core = model.module if wrap else model
    ema_core = ema_model.module if wrap else ema_model
    num = den = 0.0
    for k, p in core.named_parameters():
        if not p.requires_grad:
            assert torch.equal(p.detach(), init[k])
            continue
        got, want = p.detach().float().cpu(), ref.sd[k].detach().cpu()
        if p.dim() >= 2:
            assert relerr(got, want) < 2e-2, k
        num += float(((got - want) ** 2).sum())
        den += float(((want - init[k].cpu()) ** 2).sum())
        e_got, e_want = dict(ema_core.named_parameters())[k].detach().float().cpu(), ref.ema[k].cpu()
        if p.dim() >= 2:
            assert relerr(e_got, e_want) < 2e-2, ("ema", k)
    # the three updates of the whole model as one vector (Adam turns ~zero gradients into +-lr noise, see test_gpu_step)
    assert (num / den) ** 0.5 < 0.2, (num / den) ** 0.5
    # the EMA model (a deepcopy updated through .data) is used in eval mode: its forward must see the EMA weights
    ema_core.eval()
    x = torch.randn(2, 4, img, img, device=DEV); t = torch.full((2,), 500.0, device=DEV); y = torch.tensor([1, 2], device=DEV)
    with torch.no_grad():
        o = ema_model(x, t, y)[0]
        sd = {k: v.detach() for k, v in ema_core.state_dict().items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            o_ref = dit_forward(sd, x, t, y, **cfg)[0]
    assert relerr(o, o_ref) < 2e-2


def test_gradscaler_skips_the_fused_step_on_overflow_and_backs_off():
    """trainer.py:126-129 with an inf in the gradients: scaler.step must leave parameters, moments and the bf16 shadow
    untouched, scaler.update must halve the scale, and the next clean step must behave like step 1 (bias correction)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from torch.cuda.amp import GradScaler
        scaler = GradScaler(init_scale=1024.0)
    torch.manual_seed(0)
    m = DiT(image_size=16, patch_size=2, in_channels=4, hidden_size=128, depth=2, num_heads=2, class_dropout_prob=0.0,
            num_classes=10).to(DEV).train()
    dezero(m)
    opt = FusedAdamW(m, lr=1e-3, betas=(0.9, 0.95))
    x = torch.randn(4, 4, 16, 16, device=DEV); t = torch.rand(4, device=DEV) * 999; y = torch.randint(0, 10, (4,), device=DEV)
    ref_params = [torch.nn.Parameter(p.detach().clone()) for p in m.parameters() if p.requires_grad]
    ref_opt = torch.optim.AdamW(ref_params, lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0)
    for it, poison in enumerate((True, False, False)):
        loss = (m(x, t, y)[0].float() ** 2).mean()
        scaler.scale(loss).backward()
        if poison:
            m.blocks[0].mlp.fc1.weight.grad[0, 0] = float("inf")
        before = m._flat.detach().clone()
        scale = scaler.get_scale()
        if not poison:
            for rp, p in zip(ref_params, [p for p in m.parameters() if p.requires_grad]):
                rp.grad = p.grad.detach().clone() / scale
            ref_opt.step()
        scaler.step(opt)            # no unscale_ before: the kernel unscales with the scaler's device-side 1/scale
        scaler.update()
        opt.zero_grad()
        if poison:
            assert torch.equal(m._flat.detach(), before) and opt.step_count == 0
            assert scaler.get_scale() == scale * 0.5
            assert float(opt.m.abs().sum()) == 0.0
        else:
            assert not torch.equal(m._flat.detach(), before) and opt.step_count == it
    for rp, p in zip(ref_params, [p for p in m.parameters() if p.requires_grad]):
        torch.testing.assert_close(p.detach(), rp.detach(), rtol=2e-5, atol=2e-6)


def test_fused_adamw_is_a_torch_optimizer_with_checkpointable_state():
    """main.py:354-355 / tools/utils.py:93-120: isinstance(Optimizer), real param_groups, state_dict keys of torch's AdamW
    (step / exp_avg / exp_avg_sq) that load into torch.optim.AdamW and back."""
    torch.manual_seed(0)
    m = DiT(image_size=16, patch_size=2, in_channels=4, hidden_size=128, depth=2, num_heads=2, class_dropout_prob=0.0,
            num_classes=10).to(DEV).train()
    dezero(m)
    decay, no_decay = [], []
    for k, p in m.named_parameters():
        if p.requires_grad:
            (no_decay if p.dim() < 2 else decay).append(p)
    opt = FusedAdamW(m, lr=1e-3, betas=(0.9, 0.95), params=[{"params": decay, "weight_decay": 0.1},
                                                           {"params": no_decay, "weight_decay": 0.0}])
    assert isinstance(opt, torch.optim.Optimizer) and len(opt.param_groups) == 2
    assert sum(len(g["params"]) for g in opt.param_groups) == len(decay) + len(no_decay)
    twin = [torch.nn.Parameter(p.detach().clone()) for p in decay + no_decay]
    ref = torch.optim.AdamW([{"params": twin[:len(decay)], "weight_decay": 0.1},
                             {"params": twin[len(decay):], "weight_decay": 0.0}], lr=1e-3, betas=(0.9, 0.95))
    x = torch.randn(4, 4, 16, 16, device=DEV); t = torch.rand(4, device=DEV) * 999; y = torch.randint(0, 10, (4,), device=DEV)

    def one(o, r):
        (m(x, t, y)[0].float() ** 2).mean().backward()
        for tw, p in zip(twin, decay + no_decay):
            tw.grad = p.grad.detach().clone()
        o.step(); o.zero_grad(); r.step()

    one(opt, ref); one(opt, ref)
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 2.0
    for tw, p in zip(twin, decay + no_decay):
        torch.testing.assert_close(p.detach(), tw.detach(), rtol=2e-5, atol=2e-6)
    # the state loads into torch's AdamW (same layout) ...
    ref2 = torch.optim.AdamW([{"params": twin[:len(decay)], "weight_decay": 0.1},
                              {"params": twin[len(decay):], "weight_decay": 0.0}], lr=1e-3, betas=(0.9, 0.95))
    ref2.load_state_dict(copy.deepcopy(sd))
    # ... and into a fresh FusedAdamW, which then continues exactly like the uninterrupted one
    opt2 = FusedAdamW(m, lr=1e-3, betas=(0.9, 0.95), params=[{"params": decay, "weight_decay": 0.1},
                                                            {"params": no_decay, "weight_decay": 0.0}])
    opt2.load_state_dict(copy.deepcopy(sd))
    assert opt2.step_count == 2
    one(opt2, ref2)
    for tw, p in zip(twin, decay + no_decay):
        torch.testing.assert_close(p.detach(), tw.detach(), rtol=2e-5, atol=2e-6)
    with pytest.raises(Exception):
        FusedAdamW(m, params=[torch.nn.Parameter(torch.zeros(3, device=DEV))]).step()
