"""Spectral initial condition of sbm.jl (`--init=normalizedLaplacian`, the default): host-side, as in the reference.

`Laplacian` (sbm.jl:377-388), `Kmeans` (sbm.jl:215-244) and `CrsbyKmeans` (sbm.jl:246-258).  ARPACK (`eigs`) through
scipy; the k-means is the reference's own Lloyd iteration from a random labelling, not a library call."""
from __future__ import annotations

import numpy as np


def laplacian(links, Ntot: int, B: int):
    """sbm.jl:377-388: the B leading eigenvectors of I - D^-1/2 A D^-1/2 (shift-and-largest-magnitude, as upstream)."""
    from scipy.sparse import coo_matrix, diags, identity
    from scipy.sparse.linalg import eigs
    shift = 1000
    K = links.shape[0]
    A = coo_matrix((np.ones(K), (links[:, 0] - 1, links[:, 1] - 1)), shape=(Ntot, Ntot)).tocsr()
    deg = np.asarray(A.sum(axis=1)).ravel()
    Dh = diags(1.0 / np.sqrt(deg))
    L = identity(Ntot, format="csr") - Dh @ A @ Dh - shift * identity(Ntot, format="csr")
    vals, V = eigs(L, k=B, which="LM")
    return vals + shift, np.real(V)


def kmeans(V: np.ndarray, B: int, rng: np.random.Generator):
    """sbm.jl:215-244."""
    Ntot = V.shape[0]
    itmax, threshold, distprev = 64, 0.0001, 10.0
    labels = rng.integers(1, B + 1, size=Ntot)
    distance = np.zeros((Ntot, B))
    for _ in range(itmax):
        for k in range(1, B + 1):
            Vk = V[labels == k]
            with np.errstate(invalid="ignore"):
                mean = Vk.mean(axis=0) if len(Vk) else np.full(V.shape[1], np.nan)
            distance[:, k - 1] = ((V - mean) ** 2).sum(axis=1)
        d = np.where(np.isnan(distance), np.inf, distance)
        labels = np.argmin(d, axis=1) + 1
        m = float(np.nanmean(distance))
        if abs(m - distprev) < threshold:
            break
        distprev = m
    return labels


def crs_by_kmeans(links, labels, Ntot: int, B: int):
    """sbm.jl:246-258."""
    counts = np.bincount(labels - 1, minlength=B).astype(np.float64)
    gr = np.where(counts == 0, 1.0 / Ntot, counts / Ntot)
    Crs = np.zeros((B, B))
    np.add.at(Crs, (labels[links[:, 0] - 1] - 1, labels[links[:, 1] - 1] - 1), 1.0)
    return gr, Crs / (Ntot * np.outer(gr, gr))


def laplacian_initial(links, Ntot: int, B: int, rng: np.random.Generator):
    """sbm.jl:279-282."""
    links = np.asarray(links, dtype=np.int64)
    if B == 1:
        return np.ones(1), np.array([[links.shape[0] / Ntot]])
    _, V = laplacian(links, Ntot, B)
    labels = kmeans(V, B, rng)
    return crs_by_kmeans(links, labels, Ntot, B)
