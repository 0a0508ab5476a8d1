"""The stated equivalence (BASELINE.json north_star): from the same initial messages and parameters the CUDA
path (synchronous sweeps) and the oracle (the reference's asynchronous sweeps, oracle/sbm_oracle.py) reach
the same fixed point: marginals within 1e-6 max-abs, gr and Crs within 1e-6 relative, CV errors within
1e-5 relative, identical assignments.  Oracle outputs come from tests/golden/ (make_golden.py)."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sbm_em_*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[7:-4] for p in GOLDEN])
def test_em_fixed_point_matches_oracle(gpu_required, path):
    from graphbix_b200.engine import EM
    g = np.load(path)
    n, q = int(g["n"]), int(g["q"])
    out, eng, res = EM(n, q, g["links"], n, (g["gr0"], g["C0"]), bool(g["dc"]), 4000, bool(g["prior"]),
                       float(g["lr"]), PSIcav0=g["psi0"], BPconvthreshold=float(g["tol"]), return_engine=True)
    gr, Crs, PSI = out[0], out[1], out[2]
    assert out[13], "CUDA path did not converge"
    assert res.nonfinite == 0
    assert np.abs(PSI - g["PSI"]).max() < 1e-6                       # marginals: 1e-6 max-abs per vertex
    np.testing.assert_allclose(gr.ravel(), g["gr"], rtol=1e-6)        # gamma: 1e-6 relative
    np.testing.assert_allclose(Crs, g["Crs"], rtol=1e-6)              # c_ab: 1e-6 relative
    np.testing.assert_allclose(np.array(out[3:12]), g["scalars"], rtol=1e-5)   # FE, CV errors, variances
    assert np.array_equal(eng.assignment(), np.argmax(g["PSI"], axis=1) + 1)   # identical assignments
    np.testing.assert_allclose(eng.messages(), g["PSIcav"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(eng.state()["theta"], g["theta"], rtol=1e-6)
    eng.close()


@pytest.mark.parametrize("q,seed,prior", [(2, 1, True), (3, 2, True), (4, 3, False), (8, 4, False)])
def test_cuda_marginals_are_exact_on_a_tree(gpu_required, q, seed, prior):
    """Independent of the oracle: on a tree BP is exact, so the CUDA path's fixed point (affinity frozen, prior learning
    on) must equal brute-force enumeration of prod_i gr e^{-h} prod_edges Crs, and its Bethe free energy -log Z."""
    import math

    from test_oracle_exact_tree import exact_marginals, random_tree
    from graphbix_b200.engine import EM, SbmEngine
    rng = np.random.default_rng(seed)
    n = 9 if q < 8 else 6
    links = random_tree(n, rng)
    K = len(links)
    A = rng.uniform(0.5, 3.0, size=(q, q))
    C = 0.5 * (A + A.T) + 2.0 * np.eye(q)
    gr0 = rng.dirichlet(np.ones(q) * 5)
    psi0 = rng.random((K, q))
    psi0 /= psi0.sum(axis=1, keepdims=True)
    out, eng, res = EM(n, q, links, 1e9, (gr0, C), False, 3000, prior, 0.0, PSIcav0=psi0, BPconvthreshold=1e-14,
                       return_engine=True)
    gr, Crs, PSI, FE = out[0].ravel(), out[1], out[2], out[3]
    assert out[13] and res.nonfinite == 0
    np.testing.assert_allclose(Crs, C, rtol=1e-15)
    if not prior:   # gr stays the mean of the initial marginals through the sweeps (sbm.jl:341): read it after init
        e0 = SbmEngine(n, links, q)
        e0.set_options(degreecorrection=0, priorlearning=0, learningrate=0.0)
        e0.init(gr0, C, psi0)
        gr = e0.state()["gr"]
        e0.close()
    h = PSI.sum(axis=0) @ C / n
    marg, logZ = exact_marginals(n, q, links, gr, C, np.tile(h, (n, 1)))
    assert np.abs(marg - PSI).max() < 1e-9
    if prior:
        bethe = -n * (FE + 0.5 * float(gr @ C @ gr)) - (K // 2) * math.log(n)
        assert abs(bethe - logZ) < 1e-8
    eng.close()


BIG = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sbmbig_em_*.npz")))


@pytest.mark.parametrize("path", BIG, ids=[os.path.basename(p)[10:-4] for p in BIG])
def test_q8_fixed_point_matches_c_oracle(gpu_required, path):
    """The headline q = 8 (n = 6000, degree 14; with and without degree correction / prior learning) against fixed
    points of the C oracle (the reference's asynchronous schedule), north_star tolerances."""
    from helpers import big_case
    from graphbix_b200.engine import EM
    g = np.load(path)
    n, links, psi0, _ = big_case(g)
    q = int(g["q"])
    out, eng, res = EM(n, q, links, n, (g["gr0"], g["C0"]), bool(g["dc"]), 4000, bool(g["prior"]), float(g["lr"]),
                       PSIcav0=psi0, BPconvthreshold=float(g["tol"]), return_engine=True)
    assert out[13] and res.nonfinite == 0
    assert np.abs(out[2] - g["PSI"]).max() < 1e-6
    np.testing.assert_allclose(out[0].ravel(), g["gr"].ravel(), rtol=1e-6)
    np.testing.assert_allclose(out[1], g["Crs"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(np.array(out[3:12]), g["scalars"], rtol=1e-5)
    assert np.array_equal(eng.assignment(), np.argmax(g["PSI"], axis=1) + 1)
    np.testing.assert_allclose(eng.state()["theta"], g["theta"], rtol=1e-6)
    eng.close()


def test_reference_default_conditions(gpu_required):
    """sbm.jl's defaults: `uniformAssortative` Crs, uniform random messages, BPconvthreshold 1e-6, itrmax 256.  A
    label-symmetric start lets every schedule (every randperm of the reference itself) break the symmetry its own way
    and the default threshold can fire on the way out of the symmetric saddle, so single runs are compared the way the
    reference's driver uses them (sbm.jl:912-943: the sample of lowest free energy wins): the best of the seeds must
    agree, and wherever both schedules land in the same optimum the marginals agree up to the label permutation."""
    from helpers import assortative_init, best_permutation, planted_case, sha
    from graphbix_b200.engine import EM
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sbmdefault_q4.npz"))
    q = int(g["q"])
    gr0, C0 = assortative_init(q)
    n, links, _, _ = planted_case(int(g["n_req"]), q, float(g["deg"]), float(g["eps"]), int(g["graph_seed"]), informative=False)
    assert sha(links) == str(g["links_sha"]), "generator drifted: make_golden.py default"
    K = links.shape[0]

    def messages(seed):
        psi0 = np.random.default_rng(int(seed)).random((K, q))
        return psi0 / psi0.sum(axis=1, keepdims=True)

    fe_gpu, same = [], 0
    for k, seed in enumerate(g["seeds"]):
        psi0 = messages(seed)
        assert sha(psi0) == str(g["shas"][k])
        out = EM(n, q, links, n, (gr0, C0), False, 256, True, 0.3, PSIcav0=psi0, BPconvthreshold=1e-6)
        fe_gpu.append(out[3])
        if abs(out[3] - g["FE"][k]) < 2e-4 * abs(g["FE"][k]) and out[13] and g["cnv"][k]:
            ref = g["PSI32"][k].astype(np.float64)
            perm = best_permutation(ref, out[2])
            # both stopped when the MEAN change fell below 1e-6: close on average, borderline vertices still move
            assert np.abs(out[2][:, perm] - ref).mean() < 2e-3
            assert (np.argmax(out[2][:, perm], 1) == np.argmax(ref, 1)).mean() > 0.99
            same += 1
    assert same >= 1
    # the planted optimum is the lowest free energy both find (the reference keeps the minimum over `initnum` samples)
    np.testing.assert_allclose(min(fe_gpu), g["FE"].min(), rtol=1e-5)
    # and run to 1e-12 from the same random start the fixed points agree to north_star's tolerances up to the labels
    out = EM(n, q, links, n, (gr0, C0), False, 4000, True, 0.3, PSIcav0=messages(g["seeds"][0]), BPconvthreshold=1e-12)
    assert out[13]
    perm = best_permutation(g["tight_PSI"], out[2])
    assert np.abs(out[2][:, perm] - g["tight_PSI"]).max() < 1e-6
    np.testing.assert_allclose(out[0].ravel()[perm], g["tight_gr"].ravel(), rtol=1e-6)
    np.testing.assert_allclose(out[1][np.ix_(perm, perm)], g["tight_Crs"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(np.array(out[3:12]), g["tight_scalars"], rtol=1e-5)
