"""Generates tests/golden/{sbm,mod,lsbm}_em_*.npz: inputs and converged outputs of the ORACLE (oracle/sbm_oracle.py, the
restatement of the reference's EM(), src/sbm.jl:28-370) on small planted-partition graphs.

The reference itself cannot run here (no Julia), so these are oracle outputs, not reference outputs;
they pin the oracle against regressions and let the GPU tests run without re-running the slow Python loops.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from helpers import assortative_init, planted_case  # noqa: E402
from oracle import sbm_oracle as O  # noqa: E402

CASES = [
    # name, n, q, deg, eps, seed, dc, prior, init
    ("q2_dc_prior", 260, 2, 8.0, 0.1, 42, True, True, "assortative"),
    ("q3_nodc_prior", 260, 3, 8.0, 0.1, 43, False, True, "assortative"),
    ("q3_nodc_noprior", 260, 3, 8.0, 0.1, 44, False, False, "assortative"),
    ("q4_dc_prior", 300, 4, 9.0, 0.08, 45, True, True, "assortative"),
    ("q3_dc_noprior", 260, 3, 8.0, 0.1, 46, True, False, "assortative"),
    ("q2_disassortative", 240, 2, 8.0, 8.0, 47, False, True, "disassortative"),
]


MOD_CASES = [
    # name, n, q, deg, eps, seed, dc, learning, prior
    ("q2_dc_learn", 260, 2, 8.0, 0.1, 62, True, True, False),
    ("q3_nodc_learn_prior", 260, 3, 8.0, 0.1, 63, False, True, True),
    ("q3_dc_learn", 260, 3, 8.0, 0.1, 63, True, True, False),
    ("q4_dc_nolearn", 260, 4, 8.0, 0.1, 64, True, False, False),
]


def main_mod():
    from oracle import mod_oracle as M
    for name, n, q, deg, eps, seed, dc, learning, prior in MOD_CASES:
        n, links, psi0, lab = planted_case(n, q, deg, eps, seed)
        gr = np.ones(q) / q
        r = M.EM(gr, n, q, links, 1, 4000, 0.0, learning, prior, dc, 0.3, rng=np.random.default_rng(seed),
                 PSIcav0=psi0, BPconvthreshold=1e-12)
        print("mod", name, "n", n, "K", links.shape[0], "cnv", r.cnv, "itr", r.itrs_done, "FE", r.FE)
        if not r.cnv:
            print("   -> oracle did not converge, skipped")
            continue
        np.savez_compressed(
            os.path.join(HERE, f"mod_em_{name}.npz"), n=n, q=q, links=links, psi0=psi0, gr0=gr, dc=dc,
            learning=learning, prior=prior, lr=0.3, tol=1e-12, PSI=r.PSI, PSIcav=r.PSIcav, gr=r.gr,
            params=np.array([r.excm, r.alpha, r.beta, r.omegain, r.omegaout]),
            scalars=np.array([r.FE, r.CVBayes, r.CVGP, r.CVGT, r.CVMAP, r.varCVBayes, r.varCVGP, r.varCVGT,
                              r.varCVMAP]), itrnum=r.itrnum)


LABELED_CASES = [
    # name, n, B, mean degree per label, kind per label, seed, dc
    ("B2_G2_dc", 240, 2, [5.0, 4.0], ["a", "d"], 3, True),
    ("B2_G2_nodc", 240, 2, [5.0, 4.0], ["a", "d"], 4, False),
    ("B3_G2_dc", 300, 3, [6.0, 4.0], ["a", "a"], 5, True),
]


def main_labeled():
    """Fixtures for the labeled SBM of the reference's notebook (DESIGN.md row (f'); its CUDA path is the next row)."""
    from oracle import labeled_oracle as L
    for name, n, B, cs, kinds, seed, dc in LABELED_CASES:
        links, lab = L.labeled_planted_partition(n, B, cs, kinds, eps=0.08, seed=seed)
        n = int(links[:, :2].max())
        G = len(cs)
        und = links[links[:, 0] < links[:, 1]]
        Larray = [int((und[:, 2] == a + 1).sum()) for a in range(G)]
        rng = np.random.default_rng(seed + 100)
        K = len(links)
        psi0 = rng.random((K, B)) * 0.5
        psi0[np.arange(K), lab[links[:, 0] - 1]] += 1.0
        psi0 /= psi0.sum(1, keepdims=True)
        aff0 = L.initial_affinity(n, K, Larray, B, G, dc, 0, kinds, 1 / B)
        r = L.EM(n, Larray, B, G, links, 4000, 1e-12, True, True, dc, 0.3, 0, kinds, 1 / B, rng=rng, PSIcav0=psi0)
        print("labeled", name, "n", n, "K", K, "cnv", r.cnv, "itr", r.itrs_done, "FE", r.FE)
        if not r.cnv:
            print("   -> oracle did not converge, skipped")
            continue
        np.savez_compressed(
            os.path.join(HERE, f"lsbm_em_{name}.npz"), n=n, B=B, G=G, links=links, Larray=np.array(Larray), psi0=psi0,
            affinity0=aff0, dc=dc, lr=0.3, tol=1e-12, labels=lab, gr=r.gr, affinity=r.affinity, PSI=r.PSI,
            PSIcav=r.PSIcav, scalars=np.array([r.FE, r.CVBayes, r.CVGP, r.CVGT, r.CVMAP, r.varCVBayes, r.varCVGP,
                                               r.varCVGT, r.varCVMAP]), itrnum=r.itrnum)


def main():
    for name, n, q, deg, eps, seed, dc, prior, init in CASES:
        n, links, psi0, lab = planted_case(n, q, deg, eps, seed)
        if init == "assortative":
            gr0, C0 = assortative_init(q)
        else:
            gr0, C0 = np.ones(q) / q, -9.9 * np.eye(q) + 10 * np.ones((q, q))
        r = O.EM(n, q, links, n, (gr0, C0), dc, 4000, prior, 0.3, rng=np.random.default_rng(seed),
                 PSIcav0=psi0, BPconvthreshold=1e-12)
        print(name, "n", n, "K", links.shape[0], "cnv", r.cnv, "itr", r.itrs_done, "BPconv", r.BPconv, "FE", r.FE)
        if not r.cnv:
            print("   -> oracle did not converge, skipped")
            continue
        np.savez_compressed(
            os.path.join(HERE, f"sbm_em_{name}.npz"), n=n, q=q, links=links, psi0=psi0, gr0=gr0, C0=C0,
            dc=dc, prior=prior, lr=0.3, tol=1e-12, labels=lab, gr=r.gr, Crs=r.Crs, PSI=r.PSI, PSIcav=r.PSIcav,
            theta=r.theta, scalars=np.array([r.FE, r.CVBayes, r.CVGP, r.CVGT, r.CVMAP, r.varCVBayes, r.varCVGP,
                                             r.varCVGT, r.varCVMAP]), itrnum=r.itrnum)


BIG_CASES = [
    # name, n, q, deg, eps, seed, dc, prior  -- the headline q = 8 through the C oracle (same schedule as sbm_oracle.py)
    ("q8_nodc_prior", 6000, 8, 14.0, 0.05, 71, False, True),
    ("q8_dc_prior", 6000, 8, 14.0, 0.05, 72, True, True),
    ("q8_dc_noprior", 6000, 8, 14.0, 0.05, 74, True, False),
]


def digest(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main_big():
    """q = 8 fixed points (n = 6000) from the C oracle.  Inputs are regenerated from the seed by the tests
    (helpers.planted_case); their SHA-256 is stored so that a drifting generator fails loudly instead of silently."""
    from oracle import sbm_oracle_c as OC
    for name, n, q, deg, eps, seed, dc, prior in BIG_CASES:
        n2, links, psi0, lab = planted_case(n, q, deg, eps, seed)
        gr0, C0 = assortative_init(q)
        r = OC.EM(n2, q, links, n2, (gr0, C0), dc, 4000, prior, 0.3, PSIcav0=psi0, BPconvthreshold=1e-12,
                  float32_logbiprob=True)
        print("big", name, "n", n2, "K", links.shape[0], "cnv", r.cnv, "itr", r.itrs_done, "FE", r.FE)
        assert r.cnv
        np.savez_compressed(
            os.path.join(HERE, f"sbmbig_em_{name}.npz"), n_req=n, n=n2, q=q, deg=deg, eps=eps, seed=seed, dc=dc,
            prior=prior, lr=0.3, tol=1e-12, links_sha=digest(links), psi0_sha=digest(psi0), gr0=gr0, C0=C0,
            gr=r.gr, Crs=r.Crs, PSI=r.PSI, theta=r.theta,
            scalars=np.array([r.FE, r.CVBayes, r.CVGP, r.CVGT, r.CVMAP, r.varCVBayes, r.varCVGP, r.varCVGT,
                              r.varCVMAP]), itrnum=r.itrnum)


DEFAULT_SEEDS = [81, 82, 83, 84, 85, 86]


def default_messages(K, q, seed):
    """rand(1,B) rows normalised to 1, sbm.jl:319-323."""
    psi0 = np.random.default_rng(seed).random((K, q))
    return psi0 / psi0.sum(axis=1, keepdims=True)


def main_default():
    """The reference's default conditions (sbm.jl:293-295 `uniformAssortative`, sbm.jl:319-323 uniform random messages,
    BPconvthreshold = 1e-6, itrmax = 256, no degree correction) on a planted 4-group graph, one run per seed, plus one
    run of the first seed to 1e-12.  A label-symmetric start lets every schedule (and every randperm of the reference)
    break the symmetry its own way, so what is comparable is what the reference's driver compares: the free energies
    of the `initnum` samples and the best of them (sbm.jl:912-943)."""
    from oracle import sbm_oracle_c as OC
    n, q, deg, eps = 3000, 4, 10.0, 0.05
    gr0, C0 = assortative_init(q)
    n2, links, _, lab = planted_case(n, q, deg, eps, 81, informative=False)       # ONE graph ...
    K = links.shape[0]
    FE, CVB, cnv, itr, PSIs, shas = [], [], [], [], [], []
    for seed in DEFAULT_SEEDS:                                                      # ... several random starts
        psi0 = default_messages(K, q, seed)
        r = OC.EM(n2, q, links, n2, (gr0, C0), False, 256, True, 0.3, PSIcav0=psi0, BPconvthreshold=1e-6,
                  float32_logbiprob=True, seed=seed)
        print("default", seed, "n", n2, "cnv", r.cnv, "itr", r.itrs_done, "FE", r.FE)
        FE.append(r.FE); CVB.append(r.CVBayes); cnv.append(r.cnv); itr.append(r.itrnum)
        PSIs.append(r.PSI.astype(np.float32)); shas.append(digest(psi0))
    psi0 = default_messages(K, q, DEFAULT_SEEDS[0])
    r = OC.EM(n2, q, links, n2, (gr0, C0), False, 4000, True, 0.3, PSIcav0=psi0, BPconvthreshold=1e-12,
              float32_logbiprob=True, seed=DEFAULT_SEEDS[0])
    print("default tight", "cnv", r.cnv, "itr", r.itrs_done, "FE", r.FE)
    assert r.cnv
    np.savez_compressed(
        os.path.join(HERE, "sbmdefault_q4.npz"), n_req=n, n=n2, q=q, deg=deg, eps=eps, graph_seed=81,
        seeds=np.array(DEFAULT_SEEDS), links_sha=digest(links), gr0=gr0, C0=C0, FE=np.array(FE), CVBayes=np.array(CVB),
        cnv=np.array(cnv), itrnum=np.array(itr), PSI32=np.stack(PSIs), shas=np.array(shas),
        tight_PSI=r.PSI, tight_gr=r.gr, tight_Crs=r.Crs,
        tight_scalars=np.array([r.FE, r.CVBayes, r.CVGP, r.CVGT, r.CVMAP, r.varCVBayes, r.varCVGP, r.varCVGT,
                                r.varCVMAP]))


def main_fixture():
    """The one data set the reference ships: src/labeled_sbm/labeledSBMedgelist_Nblock=[300,300,300]_c=[5,3]_X=[0.1,0.9]_1.txt
    (900 vertices, three planted blocks of 300, label 1 disassortative, label 2 assortative), copied verbatim to
    tests/golden/reference_labeledSBMedgelist_900.txt.  Oracle fixed point for B = 3 with the notebook's settings
    (dc, learning, priorlearning, assortativity ["d","a"], Omega = 1/B, no random edge discarding) from uniform
    random messages, run to 1e-12."""
    from graphbix_b200.labeled import initial_affinity, labeledsimplegraph
    from graphbix_b200.labeled_main import edges_per_label, read_labeled_edgelist
    from oracle import labeled_oracle as L
    links, ids = read_labeled_edgelist(os.path.join(HERE, "reference_labeledSBMedgelist_900.txt"))
    s = labeledsimplegraph(links)
    n, K, B, G = len(ids), len(s), 3, 2
    Larray = edges_per_label(s)
    rng = np.random.default_rng(0)
    psi0 = rng.random((K, B))
    psi0 /= psi0.sum(1, keepdims=True)
    aff0 = initial_affinity(n, K, Larray, B, G, True, 0, ["d", "a"], 1 / B)
    r = L.EM(n, Larray, B, G, s, 4000, 1e-12, True, True, True, 0.3, 0, ["d", "a"], 1 / B, rng=rng, PSIcav0=psi0)
    print("fixture B=3", "n", n, "K", K, "cnv", r.cnv, "itr", r.itrs_done, "FE", r.FE, "CVBayes", r.CVBayes)
    assert r.cnv
    np.savez_compressed(
        os.path.join(HERE, "lsbm_reference_fixture_B3.npz"), n=n, B=B, G=G, Larray=np.array(Larray), ids=np.array(ids),
        psi0_sha=digest(psi0), links_sha=digest(s), gr=r.gr, affinity=r.affinity, PSI=r.PSI,
        scalars=np.array([r.FE, r.CVBayes, r.CVGP, r.CVGT, r.CVMAP, r.varCVBayes, r.varCVGP, r.varCVGT, r.varCVMAP]),
        itrnum=r.itrnum)


if __name__ == "__main__":
    if "fixture" in sys.argv[1:]:
        main_fixture()
    elif "big" in sys.argv[1:]:
        main_big()
    elif "default" in sys.argv[1:]:
        main_default()
    elif "labeled" in sys.argv[1:]:
        main_labeled()
    elif "mod" in sys.argv[1:]:
        main_mod()
    elif "sbm" in sys.argv[1:]:
        main()
    else:
        main()
        main_mod()
        main_labeled()
        main_big()
        main_default()
        main_fixture()
