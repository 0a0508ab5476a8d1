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


def planted_partition_device(n: int, q: int, avg_degree: float, epsilon: float = 0.1, seed: int = 0,
                             device="cuda", shuffle: bool = False):
    """The same ensemble as `planted_partition`, drawn with torch on a GPU for graphs too large to generate on the
    host in reasonable time (N = 50M, degree 20: 10^9 directed links).  The number of edge draws is fixed at
    n * avg_degree / 2 (instead of Poisson counts per block pair) and each draw is within-group with probability
    1 / (1 + (q-1) epsilon).  `shuffle` renumbers the vertices at random, so that a contiguous vertex partition cuts
    (P-1)/P of the links (worst case) instead of following the planted groups.

    Returns (links_cm, labels): links_cm is a (2, K) int64 CUDA tensor = the K x 2 `links` matrix in column-major
    order (1-based, both directions, no self loops or multi-edges); labels is an (n,) int64 CUDA tensor."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    m = int(n * avg_degree / 2)
    p_in = 1.0 / (1.0 + (q - 1) * epsilon)
    u = torch.randint(0, n, (m,), generator=g, device=dev, dtype=torch.int64)
    gv = u * q // n
    if q > 1:
        other = torch.rand(m, generator=g, device=dev) >= p_in
        shift = torch.randint(1, q, (m,), generator=g, device=dev, dtype=torch.int64)
        gv = torch.where(other, (gv + shift) % q, gv)
        del other, shift
    start = (gv * n + q - 1) // q            # first vertex with label gv under labels = i * q // n
    size = ((gv + 1) * n + q - 1) // q - start
    del gv
    v = start + (torch.rand(m, generator=g, device=dev, dtype=torch.float64) * size).to(torch.int64).clamp_(max=n)
    v = torch.minimum(v, start + size - 1)
    del start, size
    keep = u != v
    lo, hi = torch.minimum(u, v)[keep], torch.maximum(u, v)[keep]
    del u, v, keep
    key = torch.unique(lo * n + hi)
    del lo, hi
    lo, hi = key // n, key % n
    del key
    labels = torch.arange(n, device=dev, dtype=torch.int64) * q // n
    if shuffle:
        perm = torch.randperm(n, generator=g, device=dev)
        lo, hi = perm[lo], perm[hi]
        newlab = torch.empty_like(labels)
        newlab[perm] = labels
        labels = newlab
        del perm
    E = lo.numel()
    links = torch.empty((2, 2 * E), dtype=torch.int64, device=dev)
    links[0, :E] = lo + 1
    links[1, :E] = hi + 1
    links[0, E:] = hi + 1
    links[1, E:] = lo + 1
    return links, labels


def labeled_planted_partition(n: int, B: int, c_by_label, kinds, epsilon: float = 0.1, seed: int = 0):
    """Labeled planted partition in the spirit of the reference's `labeledSBMedgelist_Nblock=[300,300,300]_c=[5,3]_...`
    sample: B equal groups; label alpha (1-based in the output) has mean degree c_by_label[alpha] and is assortative
    ('a': c_out = epsilon c_in) or disassortative ('d': c_in = epsilon c_out).  A vertex pair carries at most one label.
    Returns (links K x 3 int64 with both directions of every edge, planted labels 0..B-1)."""
    rng = np.random.default_rng(seed)
    lab = (np.arange(n) * B // n).astype(np.int64)
    starts = (np.arange(B + 1) * n + B - 1) // B
    los, his, als = [], [], []
    for a, (c, kind) in enumerate(zip(c_by_label, kinds)):
        w_in, w_out = (1.0, epsilon) if kind == "a" else (epsilon, 1.0)
        p_in = w_in / (w_in + (B - 1) * w_out) if B > 1 else 1.0
        m = int(rng.poisson(c * n / 2))
        u = rng.integers(0, n, size=m)
        gv = lab[u].copy()
        if B > 1:
            other = rng.random(m) >= p_in
            gv[other] = (gv[other] + rng.integers(1, B, size=int(other.sum()))) % B
        v = starts[gv] + (rng.random(m) * (starts[gv + 1] - starts[gv])).astype(np.int64)
        v = np.minimum(v, starts[gv + 1] - 1)
        keep = u != v
        los.append(np.minimum(u, v)[keep])
        his.append(np.maximum(u, v)[keep])
        als.append(np.full(int(keep.sum()), a + 1, dtype=np.int64))
    lo, hi, al = np.concatenate(los), np.concatenate(his), np.concatenate(als)
    _, first = np.unique(lo * n + hi, return_index=True)
    lo, hi, al = lo[first], hi[first], al[first]
    links = np.empty((2 * lo.size, 3), dtype=np.int64)
    links[: lo.size] = np.column_stack([lo + 1, hi + 1, al])
    links[lo.size:] = np.column_stack([hi + 1, lo + 1, al])
    return links, lab
