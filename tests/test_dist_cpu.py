"""CPU tier, world_size 2 over gloo: the host-side logic of the multi-GPU path (rank-ordered (t, loss) gather,
bucketed gradient averaging with a no_sync window, per-rank seeds)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vaw_b200.parallel import FlatGradSync, gather_tloss, shard_seed
    from oracle import resample as ors
    try:
        # --- equal batch sizes: rank-ordered concatenation, fp32 losses bit-preserved
        rng = np.random.RandomState(shard_seed(42, rank))
        ts = torch.from_numpy(rng.randint(0, 20, size=6))
        ls = torch.from_numpy(rng.rand(6).astype(np.float32))
        gt, gl = gather_tloss(ts, ls)
        exp_t, exp_l = [], []
        for r in range(world):
            rr = np.random.RandomState(shard_seed(42, r))
            exp_t.append(rr.randint(0, 20, size=6)); exp_l.append(rr.rand(6).astype(np.float32))
        assert np.array_equal(gt.numpy(), np.concatenate(exp_t)) and np.array_equal(gl.numpy(), np.concatenate(exp_l))
        # every rank derives the identical history from the gathered entries (reference resample.py:76-79)
        h, c = ors.update_history(np.zeros((1000, 10)), np.zeros(1000, dtype=int), gt.tolist(), gl.tolist())
        hs = torch.from_numpy(h.copy()); dist.broadcast(hs, 0)
        assert np.array_equal(hs.numpy(), h)
        # --- ragged batch sizes: padding entries carry t = -1
        n = 3 + 2 * rank
        gt, gl = gather_tloss(torch.arange(n), torch.full((n,), float(rank)), ragged=True)
        valid = gt >= 0
        assert gt.numel() == world * (3 + 2 * (world - 1)) and int(valid.sum()) == sum(3 + 2 * r for r in range(world))
        assert gt[valid].tolist() == [i for r in range(world) for i in range(3 + 2 * r)]
        # --- bucketed gradient averaging + no_sync window
        g = torch.arange(10, dtype=torch.float32) * (rank + 1)
        sync = FlatGradSync(g, [[(6, 8), (8, 10)], (0, 6)])
        with sync.no_sync():
            sync.launch()
        assert torch.equal(g, torch.arange(10, dtype=torch.float32) * (rank + 1))
        sync.launch(); sync.wait()
        mean_scale = sum(r + 1 for r in range(world)) / world
        assert torch.allclose(g, torch.arange(10, dtype=torch.float32) * mean_scale)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
