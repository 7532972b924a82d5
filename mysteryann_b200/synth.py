"""Synthetic "cross-modal low-rank" data of the BASELINE.json shapes.

base  = z  @ A   + 0.1 * eps                 (z ~ N(0, I_R), R = 32)
query = z' @ A_q + 0.1 * eps + shift         (A_q = A + gap * N/sqrt(R), shift = gap * N(0, I_D), gap = 0.5)

`make_numpy` is the single-stream numpy generator of the survey probes (small sizes, CPU tests);
`make_torch` draws the same distributions chunk by chunk on a CUDA device for the 10M/100M shapes
(different random stream, same distribution).
"""
from __future__ import annotations

import math

import numpy as np

SEED = 20240430


def make_numpy(n, n_train, n_test, dim=200, rank=32, gap=0.5, seed=SEED, normalize=False):
    rng = np.random.default_rng(seed)
    f32 = np.float32
    A = (rng.standard_normal((rank, dim)) / math.sqrt(rank)).astype(f32)
    Aq = (A + gap * rng.standard_normal((rank, dim)) / math.sqrt(rank)).astype(f32)
    shift = (gap * rng.standard_normal(dim)).astype(f32)

    def draw(m, M, sh):
        z = rng.standard_normal((m, rank)).astype(f32)
        e = rng.standard_normal((m, dim)).astype(f32)
        x = z @ M + f32(0.1) * e
        if sh is not None:
            x = x + sh
        x = x.astype(f32)
        if normalize:
            x /= np.linalg.norm(x, axis=1, keepdims=True)
        return np.ascontiguousarray(x)

    base = draw(n, A, None)
    train = draw(n_train, Aq, shift)
    test = draw(n_test, Aq, shift)
    return base, train, test


def make_torch(n, n_train, n_test, dim=200, rank=32, gap=0.5, seed=SEED, normalize=False, device="cuda",
               chunk=1 << 20):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    kw = dict(device=device, dtype=torch.float32, generator=g)
    A = torch.randn(rank, dim, **kw) / math.sqrt(rank)
    Aq = A + gap * torch.randn(rank, dim, **kw) / math.sqrt(rank)
    shift = gap * torch.randn(dim, **kw)

    def draw(m, M, sh):
        out = torch.empty(m, dim, device=device, dtype=torch.float32)
        for s in range(0, m, chunk):
            e = min(m, s + chunk)
            z = torch.randn(e - s, rank, **kw)
            x = z @ M
            x.add_(torch.randn(e - s, dim, **kw), alpha=0.1)
            if sh is not None:
                x.add_(sh)
            if normalize:
                x /= x.norm(dim=1, keepdim=True)
            out[s:e] = x
        return out

    return draw(n, A, None), draw(n_train, Aq, shift), draw(n_test, Aq, shift)
