"""Shared builders for the parity tests."""
import numpy as np

from graphbix_b200.synth import planted_partition


def planted_case(n, q, deg, eps, seed, informative=True):
    """(n, links, psi0, labels): connected planted partition and initial messages in `links` order.

    With `informative`, the messages lean towards the planted labels, so the synchronous and asynchronous
    schedules fall into the same labelled fixed point (random messages + a label-symmetric Crs can break
    the permutation symmetry differently)."""
    links, lab = planted_partition(n, q, deg, eps, seed=seed, connected_only=True)
    n = int(links.max())
    K = links.shape[0]
    rng = np.random.default_rng(seed + 1000)
    psi0 = rng.random((K, q)) * (0.5 if informative else 1.0)
    if informative:
        psi0[np.arange(K), lab[links[:, 0] - 1]] += 1.0
    psi0 /= psi0.sum(1, keepdims=True)
    return n, links, psi0, lab


def assortative_init(q, cin=20.0, cout=0.01):
    return np.ones(q) / q, cin * np.eye(q) + cout * np.ones((q, q))


def sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def best_permutation(P_ref, P):
    """Column permutation `perm` of P that matches P_ref best (Hungarian on the overlap): P[:, perm] ~ P_ref."""
    from scipy.optimize import linear_sum_assignment
    _, c = linear_sum_assignment(-(P_ref.T @ P))
    return c


def big_case(g):
    """Regenerates the inputs of a tests/golden/sbmbig_em_*.npz fixture from its seed and checks their digests."""
    n, links, psi0, lab = planted_case(int(g["n_req"]), int(g["q"]), float(g["deg"]), float(g["eps"]), int(g["seed"]))
    assert n == int(g["n"]) and sha(links) == str(g["links_sha"]) and sha(psi0) == str(g["psi0_sha"]), \
        "the synthetic-graph generator drifted: regenerate tests/golden (make_golden.py big)"
    return n, links, psi0, lab
