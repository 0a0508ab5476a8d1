"""Known-answer test of the oracle's BP equations: on a tree belief propagation is exact.

At the fixed point of EM() with the affinity matrix frozen (learning rate 0) the marginals must equal the marginals of
the pairwise model the messages describe,

    P(sigma)  ~  prod_i gr[sigma_i] exp(-theta_i h[sigma_i])  prod_{(i,j) in E} Crs[sigma_i, sigma_j],

computed by brute-force enumeration of all q^N labelings (h = update_h() is the mean-field term of the non-edges,
sbm.jl:70-73, and enters as an external field), and the Bethe free energy must equal -log Z of that model exactly
(sbm.jl:150-213).  This pins cavity products, normalisation, the external field and the free-energy bookkeeping of
oracle/sbm_oracle.py (and of the C transcription, and of the labeled oracle in its one-label limit) against an
independent calculation that shares no code with them."""
import itertools
import math

import numpy as np
import pytest

from oracle import labeled_oracle as L
from oracle import sbm_oracle as O


def random_tree(n, rng):
    edges = [(int(rng.integers(0, v)), v) for v in range(1, n)]
    und = np.array([(a + 1, b + 1) for a, b in edges], dtype=np.int64)
    return np.vstack([und, und[:, ::-1]])


def exact_marginals(n, q, links, gr, C, field):
    """Brute force over q^n labelings; field[i, s] is the exponent subtracted at vertex i (theta_i h_s)."""
    states = np.array(list(itertools.product(range(q), repeat=n)))
    logw = np.log(gr)[states].sum(axis=1) - field[np.arange(n)[None, :], states].sum(axis=1)
    und = links[links[:, 0] < links[:, 1]] - 1
    for a, b in und:
        logw += np.log(C[states[:, a], states[:, b]])
    m = logw.max()
    w = np.exp(logw - m)
    Z = w.sum()
    marg = np.zeros((n, q))
    for s in range(q):
        marg[:, s] = ((states == s) * w[:, None]).sum(axis=0) / Z
    return marg, m + math.log(Z)


@pytest.mark.parametrize("q,seed,prior", [(2, 1, True), (3, 2, True), (3, 7, False), (4, 3, False)])
def test_bp_is_exact_on_a_tree(q, seed, prior):
    rng = np.random.default_rng(seed)
    n = 9
    links = random_tree(n, rng)
    K = len(links)
    A = rng.uniform(0.5, 3.0, size=(q, q))
    C = 0.5 * (A + A.T) + 2.0 * np.eye(q)
    gr0 = rng.dirichlet(np.ones(q) * 5)
    psi0 = rng.random((K, q))
    psi0 /= psi0.sum(axis=1, keepdims=True)
    r = O.EM(n, q, links, 1e9, (gr0, C), False, 2000, prior, 0.0, rng=rng, PSIcav0=psi0, BPconvthreshold=1e-14,
             float32_logbiprob=False)
    assert r.cnv
    np.testing.assert_allclose(r.Crs, C, rtol=0, atol=0)                      # learning rate 0: Crs untouched
    np.testing.assert_allclose(r.gr, r.PSI.mean(axis=0), rtol=1e-10)          # gr = update_gr(PSI) at the end, sbm.jl:365
    if prior:
        gr_run = r.gr                # prior learning keeps gr = mean marginal all along (sbm.jl:120-122)
    else:                            # otherwise gr stays the mean of the INITIAL marginals (sbm.jl:341): a 0-iteration run
        gr_run = O.EM(n, q, links, 1e9, (gr0, C), False, 0, False, 0.0, PSIcav0=psi0, float32_logbiprob=False).gr
    h = r.PSI.sum(axis=0) @ C / n                                             # update_h with theta = 1
    marg, logZ = exact_marginals(n, q, links, gr_run, C, np.tile(h, (n, 1)))
    assert np.abs(marg - r.PSI).max() < 1e-9
    L_edges = K // 2
    if prior:
        # Bethe free energy: sum_i log Z_i - sum_(ij) log Z_ij' = log Z, with Z_ij' = N Z_ij (the oracle's Z_ij carries
        # 1/N); freeenergy() evaluates Z_i with the final gr, which is the gr of the sweeps only with prior learning
        bethe = -n * (r.FE + 0.5 * float(r.gr @ C @ r.gr)) - L_edges * math.log(n)
        assert abs(bethe - logZ) < 1e-8
    # the C transcription and the labeled oracle (one label, affinity = Crs / N) land on the same fixed point
    from oracle import sbm_oracle_c as OC
    if OC.available():
        rc = OC.EM(n, q, links, 1e9, (gr0, C), False, 2000, prior, 0.0, PSIcav0=psi0, BPconvthreshold=1e-14,
                   float32_logbiprob=False)
        assert np.abs(rc.PSI - marg).max() < 1e-9
    links3 = np.column_stack([links, np.ones(K, dtype=np.int64)])
    rl = L.EM(n, [L_edges], q, 1, links3, 500, 1e-14, False, False, False, 0.3, 0, ["a"], 1 / q, rng=rng, PSIcav0=psi0,
              affinity0=(C / n)[None], gr0=gr_run)
    assert rl.cnv
    margl, _ = exact_marginals(n, q, links, gr_run, C, np.tile(rl.PSI.sum(axis=0) @ C / n, (n, 1)))
    assert np.abs(rl.PSI - margl).max() < 1e-9


def test_bp_is_exact_on_a_tree_with_degree_correction():
    rng = np.random.default_rng(11)
    n, q = 9, 2
    links = random_tree(n, rng)
    K = len(links)
    C = np.array([[3.0, 0.7], [0.7, 2.5]])
    psi0 = rng.random((K, q))
    psi0 /= psi0.sum(axis=1, keepdims=True)
    r = O.EM(n, q, links, 1e9, (np.array([0.6, 0.4]), C), True, 500, True, 0.0, rng=rng, PSIcav0=psi0,
             BPconvthreshold=1e-14, float32_logbiprob=False)
    if not r.cnv:
        pytest.skip("the discrete update_theta cycles on this tree (the reference would too)")
    h = (r.theta[:, None] * r.PSI).sum(axis=0) @ C / n                        # update_h, sbm.jl:70-73
    marg, _ = exact_marginals(n, q, links, r.gr, C, r.theta[:, None] * h[None, :])
    assert np.abs(marg - r.PSI).max() < 1e-9                                  # theta_i theta_j factors cancel in the marginals


@pytest.mark.parametrize("q,dc,seed", [(2, False, 3), (3, False, 6), (4, False, 2)])
def test_mod_oracle_bp_is_exact_on_a_tree(q, dc, seed):
    """mod.jl's model on a tree: P(sigma) ~ prod_i gr[sigma_i] exp(-alpha beta d_i h[sigma_i]) prod_edges
    (1 + (e^beta - 1) delta(sigma_i, sigma_j)) with d_i = 1 without degree correction (mod.jl:38-52); omegas frozen.
    (With degree correction the hub's self-field alpha beta d_i h makes BP oscillate on a ten-vertex tree, for the
    reference's schedule too, so only the uncorrected model is enumerated.)"""
    from oracle import mod_oracle as M
    rng = np.random.default_rng(seed)
    n = 10
    # a tree with a hub, so that the mean excess degree of mod.jl:185 is > 1 and beta is finite
    und = [(1, v) for v in range(2, 7)] + [(int(rng.integers(2, v)), v) for v in range(7, n + 1)]
    und = np.array(und, dtype=np.int64)
    links = np.vstack([und, und[:, ::-1]])
    K = len(links)
    gr = rng.dirichlet(np.ones(q) * 5)
    psi0 = rng.random((K, q))
    psi0 /= psi0.sum(axis=1, keepdims=True)
    r = M.EM(gr, n, q, links, 1, 3000, 0.0, False, False, dc, 0.3, rng=rng, PSIcav0=psi0, BPconvthreshold=1e-14)
    assert r.cnv and np.isfinite(r.beta)
    deg = np.bincount(links[:, 0] - 1, minlength=n).astype(np.float64)
    d = deg if dc else np.ones(n)
    h = d @ r.PSI                                                              # update_h, mod.jl:22-24
    C = np.ones((q, q)) + (math.exp(r.beta) - 1) * np.eye(q)
    marg, _ = exact_marginals(n, q, links, gr, C, r.alpha * r.beta * d[:, None] * h[None, :])
    assert np.abs(marg - r.PSI).max() < 1e-9
