"""CPU tier: the range bookkeeping of the sharded-optimizer data-parallel mode (no process group, no GPU).
DiT.block_shard_ranges() must tile the flat parameter buffer exactly once, its reduce-scattered ranges must split into
float4-aligned slices for every world size the benchmark runs at, and the optimizer's range intersection helper must
behave like set intersection on half-open intervals."""
import pytest
import torch

from vaw_b200.models.dit import DiT
from vaw_b200.optim import _intersect


def _model(hidden, depth, heads, learn_align=False):
    return DiT(image_size=32, patch_size=2, in_channels=4, hidden_size=hidden, depth=depth, num_heads=heads,
               class_dropout_prob=0.1, num_classes=1000, learn_align=learn_align,
               **(dict(encoder_depth=2, z_dims=768, projector_dim=2048) if learn_align else {}))


@pytest.mark.parametrize("hidden,depth,heads,align", [(1152, 28, 16, False), (384, 12, 6, False), (1152, 3, 16, True),
                                                       (128, 2, 2, False)])
def test_block_shard_ranges_tile_the_parameter_buffer(hidden, depth, heads, align):
    m = _model(hidden, depth, heads, align)
    big, small, tail, total = m.block_shard_ranges()
    assert len(big) == len(small) == depth
    per_block, tail2, total2 = m.block_grad_ranges()
    assert total == total2 and len(per_block) == depth
    # every element of every tensor is covered exactly once (the 64-element alignment gaps between tensors may or may
    # not be covered, never twice)
    cover = torch.zeros(total, dtype=torch.int8)
    for rs in list(big) + list(small) + [tail]:
        for b, e in rs:
            assert 0 <= b < e <= total
            cover[b:e] += 1
    assert int(cover.max()) == 1
    off, num, names = m._layout()
    for o, n in zip(off, num):
        if n:
            assert bool((cover[o:o + n] == 1).all())
    # the same tensors belong to block i in both bucketings
    for i in range(depth):
        a = torch.zeros(total, dtype=torch.bool)
        b_ = torch.zeros(total, dtype=torch.bool)
        for lo, hi in list(big[i]) + list(small[i]):
            a[lo:hi] = True
        for lo, hi in per_block[i]:
            b_[lo:hi] = True
        for o, n in zip(off, num):
            if n:
                assert bool(a[o:o + n].all()) == bool(b_[o:o + n].all())
    # reduce-scattered ranges: float4-aligned slices at every world size of the scaling run
    for rs in big:
        for b, e in rs:
            assert b % 4 == 0
            for world in (2, 4, 8):
                assert (e - b) % (4 * world) == 0, (b, e, world)
    lo, hi = m.stacked_range()
    inside = [(b, e) for rs in big for b, e in rs if lo <= b and e <= hi]
    assert len(inside) == depth and sum(e - b for b, e in inside) == depth * 6 * hidden * hidden


def test_intersect_is_interval_intersection():
    assert _intersect([(0, 10), (20, 30)], [(5, 25)]) == [(5, 10), (20, 25)]
    assert _intersect([(0, 10)], [(10, 20)]) == []
    assert _intersect([(0, 100)], [(3, 7), (50, 60)]) == [(3, 7), (50, 60)]
    assert _intersect([], [(0, 5)]) == []
    got = _intersect([(0, 8), (8, 16)], [(4, 12)])
    assert sum(e - b for b, e in got) == 8 and got[0][0] == 4 and got[-1][1] == 12
