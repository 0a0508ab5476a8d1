"""Synthetic planted-partition (SBM) edge lists of the shapes BASELINE.json names.

Output convention is the reference's: `links` is a K x 2 int64 array of 1-based vertex ids that
holds BOTH directions of every undirected edge, no self loops, no multi-edges -- i.e. what
`simplegraph()` (sbm.jl:414-425) returns.
"""
from __future__ import annotations

import numpy as np


def planted_partition(n: int, q: int, avg_degree: float, epsilon: float = 0.1, seed: int = 0,
                      connected_only: bool = False):
    """Sparse planted partition: q equal groups, c_out = epsilon * c_in, mean degree `avg_degree`.

    Edges are drawn as Poisson(c_rs / n) pair counts (the generator of the reference's
    labeledSBMgenerator notebook does the same per block pair), then deduplicated.
    Returns (links, planted_labels) with 1-based links and 0-based labels.
    """
    rng = np.random.default_rng(seed)
    labels = (np.arange(n) * q // n).astype(np.int64)
    # avg_degree = (c_in + (q-1) c_out)/q
    c_in = avg_degree * q / (1.0 + (q - 1) * epsilon)
    c_out = epsilon * c_in
    sizes = np.bincount(labels, minlength=q)
    starts = np.concatenate(([0], np.cumsum(sizes)))
    src_all, dst_all = [], []
    for r in range(q):
        for s in range(r, q):
            c = c_in if r == s else c_out
            npairs = sizes[r] * sizes[s] / (2.0 if r == s else 1.0)
            m = rng.poisson(c / n * npairs)
            a = rng.integers(starts[r], starts[r + 1], size=m)
            b = rng.integers(starts[s], starts[s + 1], size=m)
            src_all.append(a)
            dst_all.append(b)
    a = np.concatenate(src_all)
    b = np.concatenate(dst_all)
    keep = a != b
    a, b = a[keep], b[keep]
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    key = np.unique(lo * n + hi)
    lo, hi = key // n, key % n
    if connected_only:
        lo, hi, labels, n = _giant_component(lo, hi, labels, n)
    links = np.empty((2 * lo.size, 2), dtype=np.int64)
    links[: lo.size, 0] = lo + 1
    links[: lo.size, 1] = hi + 1
    links[lo.size:, 0] = hi + 1
    links[lo.size:, 1] = lo + 1
    return links, labels


def _giant_component(lo, hi, labels, n):
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    A = coo_matrix((np.ones(lo.size, dtype=np.int8), (lo, hi)), shape=(n, n))
    _, comp = connected_components(A, directed=False)
    big = np.argmax(np.bincount(comp))
    keep_v = comp == big
    newid = np.cumsum(keep_v) - 1
    m = keep_v[lo] & keep_v[hi]
    return newid[lo[m]], newid[hi[m]], labels[keep_v], int(keep_v.sum())


def random_messages(n_links: int, q: int, seed: int = 0) -> np.ndarray:
    """`rand(1,B)` normalised, one row per link (sbm.jl:319-323)."""
    rng = np.random.default_rng(seed)
    m = rng.random((n_links, q))
    return m / m.sum(axis=1, keepdims=True)
