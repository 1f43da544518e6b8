"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's timestep importance sampling.

Follows /root/reference/tools/resample.py: ScheduleSampler.sample :43-59, UniformSampler :62-68,
LossSecondMomentResampler.weights :142-149, update_with_all_losses :151-159, _warmed_up :161-162.

`np.random.choice(T, B, p=p)` and `np.sum` / `np.mean` are spelled out as the explicit fp64 operation sequences
numpy executes, because that sequence is what the CUDA kernel has to reproduce bit for bit:
    cdf = cumsum(p) (sequential adds); cdf /= cdf[-1]; u = random_sample(B); idx = searchsorted(cdf, u, 'right')
    pairwise sum: recurse on n/2 rounded down to a multiple of 8 until n <= 128; leaf = 8 strided accumulators
tests/test_oracle_golden.py pins every function here against numpy itself and against the executed reference.
"""
from __future__ import annotations

import numpy as np


def np_pairwise_sum(a):
    """numpy's pairwise summation of a contiguous float64 vector, written out."""
    a = np.asarray(a, dtype=np.float64)
    n = a.shape[0]
    if n < 8:
        res = np.float64(0.0)
        for i in range(n):
            res = res + a[i]
        return res
    if n <= 128:
        r = [a[j] for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = r[j] + a[i + j]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + a[i]
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return np_pairwise_sum(a[:n2]) + np_pairwise_sum(a[n2:])


def second_moment_weights(history, counts, history_per_term=10, uniform_prob=0.001):
    """LossSecondMomentResampler.weights()."""
    history = np.asarray(history, dtype=np.float64)
    T = history.shape[0]
    if not (np.asarray(counts) == history_per_term).all():
        return np.ones([T], dtype=np.float64)
    w = np.empty(T, dtype=np.float64)
    for i in range(T):
        w[i] = np.sqrt(np_pairwise_sum(history[i] * history[i]) / history.shape[1])
    w = w / np_pairwise_sum(w)
    w = w * (1 - uniform_prob)
    w = w + uniform_prob / T
    return w


def sample_from_weights(w, u):
    """ScheduleSampler.sample given the uniform draws u (float64 [B]) -> (idx int64 [B], weights float32 [B], p, cdf)."""
    w = np.asarray(w, dtype=np.float64)
    T = w.shape[0]
    p = w / np_pairwise_sum(w)
    cdf = np.empty(T, dtype=np.float64)
    c = np.float64(0.0)
    for i in range(T):
        c = p[0] if i == 0 else c + p[i]
        cdf[i] = c
    cdf = cdf / cdf[-1]
    idx = np.empty(len(u), dtype=np.int64)
    for b, ub in enumerate(np.asarray(u, dtype=np.float64)):
        lo, hi = 0, T
        while lo < hi:  # first i with cdf[i] > u  == searchsorted(side='right')
            mid = (lo + hi) // 2
            if cdf[mid] <= ub:
                lo = mid + 1
            else:
                hi = mid
        idx[b] = lo
    weights = (1 / (T * p[idx])).astype(np.float32)
    return idx, weights, p, cdf


def sample(w, batch_size):
    """Draws from numpy's global RNG exactly like np.random.choice(T, size=(B,), p=p) does."""
    return sample_from_weights(w, np.random.random_sample(batch_size))[:2]


def update_history(history, counts, ts, losses, history_per_term=10):
    """update_with_all_losses: sequential append-or-shift per entry, in place."""
    for t, loss in zip(ts, losses):
        t = int(t)
        if counts[t] == history_per_term:
            history[t, :-1] = history[t, 1:]
            history[t, -1] = loss
        else:
            history[t, counts[t]] = loss
            counts[t] += 1
    return history, counts
