"""GPU tier, needs >= 2 GPUs (skipped otherwise): data-parallel parity over NCCL.
  * per-rank gradients after the bucketed all-reduce == gradients of one process on the concatenated batch
  * the loss-aware sampler's history is identical on all ranks and equals the single-process update applied to the
    rank-ordered concatenation of (t, loss) (reference resample.py:76-79)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _uvit_worker(rank, world, port, q):
    """U-ViT under the data-parallel wrapper (BASELINE config 4 is defined on 2/4/8 GPUs): per-block buckets gated by the
    events vaw_uvit_backward_ev records, gradients == single-process gradients of the concatenated batch, and after a
    FusedAdamW step every rank holds bit-identical parameters."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from gpu_util import relerr
        from vaw_b200.models.uvit import UViT
        from vaw_b200.optim import DataParallel, FusedAdamW
        from vaw_b200.tools import gaussian_diffusion as gd
        torch.manual_seed(200 + rank)
        m = UViT(image_size=32, patch_size=4, in_channels=3, embed_dim=128, depth=5, num_heads=2, num_classes=10).to(dev).train()
        net = DataParallel(m, device_ids=[rank], output_device=rank)
        assert net._events is not None and len(net._events) == m.num_blocks + 1
        B = 4
        g = torch.Generator().manual_seed(9)
        X = torch.randn(world * B, 3, 32, 32, generator=g); Y = torch.randint(0, 10, (world * B,), generator=g)
        E = torch.randn(world * B, 3, 32, 32, generator=g); T = torch.randint(0, 1000, (world * B,), generator=g)
        d = gd.create_gaussian_diffusion(noise_schedule="linear")
        sl = slice(rank * B, (rank + 1) * B)
        terms = d.training_losses(net, X[sl].to(dev), None, t=T[sl].to(dev), model_kwargs={"y": Y[sl].to(dev)}, noise=E[sl].to(dev))
        terms["loss"].mean().backward()
        torch.cuda.synchronize()
        dp_grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
        with net.no_sync():
            for p in m.parameters():
                p.grad = None
            allt = d.training_losses(net, X.to(dev), None, t=T.to(dev), model_kwargs={"y": Y.to(dev)}, noise=E.to(dev))
            allt["loss"].mean().backward()
        worst = max(relerr(dp_grads[k], p.grad) for k, p in m.named_parameters())
        assert worst < 2e-3, f"U-ViT DP gradient mismatch {worst}"
        # one optimizer step on the all-reduced gradients: bit-identical parameters everywhere
        for k, p in m.named_parameters():
            p.grad.copy_(dp_grads[k])
        opt = FusedAdamW(net, lr=1e-3, betas=(0.9, 0.95))
        opt.step()
        chk = m._flat.detach().view(torch.int32).to(torch.int64).sum().reshape(1)
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        assert all(torch.equal(allc[0], c) for c in allc), "parameters diverged across ranks after the step"
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


def _shard_worker(rank, world, port, q):
    """DataParallel(shard_optimizer=True): reduce-scatter + 1/W AdamW + bf16 all-gather must leave every rank with the
    parameters (fp32 master after gather_master, bf16 shadow, moments of the owned slices) that all-reduce + replicated
    AdamW produces - bit for bit at two ranks (a two-term average has one rounding whichever collective computes it)."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from gpu_util import dezero
        from vaw_b200.models.dit import DiT
        from vaw_b200.optim import DataParallel, FusedAdamW
        from vaw_b200.tools import gaussian_diffusion as gd
        d = gd.create_gaussian_diffusion(noise_schedule="cosine")
        B = 4
        g = torch.Generator().manual_seed(11)
        data = [(torch.randn(world * B, 4, 16, 16, generator=g), torch.randint(0, 10, (world * B,), generator=g),
                 torch.randn(world * B, 4, 16, 16, generator=g), torch.randint(0, 1000, (world * B,), generator=g))
                for _ in range(3)]
        sl = slice(rank * B, (rank + 1) * B)
        results = {}
        for shard in (False, True):
            torch.manual_seed(5)
            m = DiT(image_size=16, patch_size=2, in_channels=4, hidden_size=128, depth=3, num_heads=2,
                    class_dropout_prob=0.0, num_classes=10, learn_align=False).to(dev).train()
            dezero(m)
            net = DataParallel(m, device_ids=[rank], shard_optimizer=shard)
            assert (net._shard_sync is not None) == shard
            opt = FusedAdamW(net, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.01)
            losses = []
            for X, Y, E, T in data:
                terms = d.training_losses(net, X[sl].to(dev), None, t=T[sl].to(dev), model_kwargs={"y": Y[sl].to(dev)},
                                          noise=E[sl].to(dev))
                terms["loss"].mean().backward()
                opt.step(); opt.zero_grad()
                losses.append(terms["loss"].detach().clone())
            torch.cuda.synchronize(dev)   # the sharded mode's bf16 all-gather runs on its own stream until the next forward
            shadow = m._shadow.clone()
            sd = {k: v.detach().clone() for k, v in net.state_dict().items()}     # gathers the fp32 master when sharded
            results[shard] = (losses, shadow, sd, m._flat.detach().clone())
        la, sa, da, fa = results[False]
        lb, sb, db, fb = results[True]
        assert all(torch.equal(x, y) for x, y in zip(la, lb)), "losses differ between the two modes"
        assert torch.equal(sa, sb), "bf16 shadows differ"
        assert torch.equal(fa, fb), "fp32 master differs after gather_master"
        assert all(torch.equal(da[k], db[k]) for k in da)
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()[-1800:]))
    finally:
        dist.destroy_process_group()


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from gpu_util import dezero, relerr
        from oracle import resample as ors
        from oracle.train_step import synthetic_history
        from vaw_b200.models.dit import DiT
        from vaw_b200.optim import DataParallel
        from vaw_b200.tools import gaussian_diffusion as gd, resample as rs
        torch.manual_seed(100 + rank)  # different init per rank: the wrapper must broadcast rank 0's weights
        m = DiT(image_size=16, patch_size=2, in_channels=4, hidden_size=128, depth=3, num_heads=2, class_dropout_prob=0.0,
                num_classes=10).to(dev).train()
        dezero(m)
        net = DataParallel(m)
        w0 = m.blocks[0].mlp.fc1.weight.detach().clone()
        ws = [torch.empty_like(w0) for _ in range(world)]
        dist.all_gather(ws, w0)
        assert all(torch.equal(ws[0], x) for x in ws), "weights not broadcast"
        B = 4
        g = torch.Generator().manual_seed(7)
        X = torch.randn(world * B, 4, 16, 16, generator=g); Y = torch.randint(0, 10, (world * B,), generator=g)
        E = torch.randn(world * B, 4, 16, 16, generator=g); T = torch.randint(0, 1000, (world * B,), generator=g)
        d = gd.create_gaussian_diffusion(noise_schedule="cosine")
        sl = slice(rank * B, (rank + 1) * B)
        terms = d.training_losses(net, X[sl].to(dev), None, t=T[sl].to(dev), model_kwargs={"y": Y[sl].to(dev)}, noise=E[sl].to(dev))
        terms["loss"].mean().backward()
        torch.cuda.synchronize()
        dp_grads = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.requires_grad}
        # single-process reference on the concatenated batch (same weights), no all-reduce
        with net.no_sync():
            for p in m.parameters():
                p.grad = None
            terms_all = d.training_losses(net, X.to(dev), None, t=T.to(dev), model_kwargs={"y": Y.to(dev)}, noise=E.to(dev))
            terms_all["loss"].mean().backward()
        worst = max(relerr(dp_grads[k], p.grad) for k, p in m.named_parameters() if p.requires_grad)
        assert worst < 2e-3, f"DP gradient mismatch {worst}"
        # sampler: identical history on every rank == single-process update with the rank-ordered concatenation
        s = rs.LossSecondMomentResampler(d)
        hist, counts = synthetic_history(0)
        s.load_history(hist, counts, dev)
        s.update_with_local_losses(T[sl].to(dev), terms["loss"].detach())
        all_losses = [torch.empty(B, device=dev) for _ in range(world)]
        dist.all_gather(all_losses, terms["loss"].detach())
        h_ref, c_ref = ors.update_history(hist.copy(), counts.copy(), T.tolist(), torch.cat(all_losses).cpu().tolist())
        assert np.array_equal(s._loss_history, h_ref) and np.array_equal(s._loss_counts, c_ref)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


def _run(target, port_base):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = port_base + os.getpid() % 300
    procs = [ctx.Process(target=target, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_data_parallel_parity():
    _run(_worker, 29600)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_uvit_data_parallel_parity():
    _run(_uvit_worker, 30000)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_optimizer_equals_all_reduce_path():
    _run(_shard_worker, 30400)
