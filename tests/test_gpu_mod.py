"""mod.jl path on the GPU: sweep-by-sweep against the numpy statement of the synchronous schedule
(tests/sync_model_mod.py), and converged results against the oracle's goldens (oracle/mod_oracle.py)."""
import glob
import os

import numpy as np
import pytest

from helpers import planted_case
from sync_model_mod import SyncMOD

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["push", "pull"])
def sweep_kind(request, monkeypatch):
    """Every test of this file runs with both sweep kernels (push: GBX_PULL=0, pull: GBX_PULL=1; where a run cannot use
    the pull kernels -- damping, q = 9..12 -- the library falls back to push)."""
    monkeypatch.setenv("GBX_PULL", "1" if request.param == "pull" else "0")
    return request.param

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "mod_em_*.npz")))


@pytest.mark.parametrize("q", [2, 3, 8])
@pytest.mark.parametrize("dc,learning,prior", [(True, True, False), (False, True, True), (True, False, False)])
def test_mod_sweeps_match_sync_model(gpu_required, q, dc, learning, prior):
    from graphbix_b200.engine import ModEngine
    n, links, psi0, _ = planted_case(500, q, 9.0, 0.15, seed=70 + q)
    gr = np.ones(q) / q
    ref = SyncMOD(n, links, q, gr, psi0, dc=dc, learning=learning, prior=prior, lr=0.3)
    eng = ModEngine(n, links, q)
    eng.set_options(dc=int(dc), learning=int(learning), priorlearning=int(prior), learningrate=0.3)
    eng.init(gr, psi0)
    np.testing.assert_allclose(eng.state()["PSI"], ref.PSI, rtol=0, atol=1e-13)
    for it in range(6):
        cref = ref.iterate()
        cgpu = eng.iterate(1)
        st = eng.state()
        np.testing.assert_allclose(cgpu, cref, rtol=1e-9, atol=1e-16)
        np.testing.assert_allclose(st["PSI"], ref.PSI, rtol=0, atol=1e-12, err_msg=f"PSI it={it}")
        np.testing.assert_allclose(eng.messages(), ref.messages_in_link_order(), rtol=0, atol=1e-12)
        np.testing.assert_allclose([st["alpha"], st["beta"], st["omegain"], st["omegaout"]],
                                   [ref.alpha, ref.beta, ref.omegain, ref.omegaout], rtol=1e-11)
        np.testing.assert_allclose(st["gr"], ref.gr, rtol=1e-12)
    res = eng.finalize()
    f = ref.finalize()
    for k in ("FE", "CVBayes", "CVGP", "CVGT", "CVMAP", "varCVBayes", "varCVGP", "varCVGT", "varCVMAP"):
        np.testing.assert_allclose(getattr(res, k), f[k], rtol=1e-10, err_msg=k)
    assert res.nonfinite == 0


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[7:-4] for p in GOLDEN])
def test_mod_em_fixed_point_matches_oracle(gpu_required, path):
    from graphbix_b200.engine import EM_mod
    g = np.load(path)
    n, q = int(g["n"]), int(g["q"])
    out, eng, res = EM_mod(g["gr0"], n, q, g["links"], 1, 4000, 0.0, bool(g["learning"]), bool(g["prior"]),
                           bool(g["dc"]), float(g["lr"]), PSIcav0=g["psi0"], BPconvthreshold=float(g["tol"]),
                           return_engine=True)
    assert out[15], "CUDA path did not converge"
    assert np.abs(out[5] - g["PSI"]).max() < 1e-6
    np.testing.assert_allclose(np.array(out[0:5]), g["params"], rtol=1e-6)          # excm, alpha, beta, omegas
    np.testing.assert_allclose(np.array(out[6:15]), g["scalars"], rtol=1e-5)        # FE, CV errors, variances
    np.testing.assert_allclose(eng.messages(), g["PSIcav"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(eng.state()["gr"], g["gr"], rtol=1e-6)
    assert np.array_equal(eng.assignment(), np.argmax(g["PSI"], axis=1) + 1)
    eng.close()


@pytest.mark.parametrize("q,seed", [(2, 3), (3, 6), (4, 2)])
def test_mod_cuda_marginals_are_exact_on_a_tree(gpu_required, q, seed):
    """Independent of the oracle: on a tree BP is exact, so the CUDA fixed point of mod.jl's model (omegas frozen, no
    degree correction) equals brute-force enumeration of prod_i gr e^{-alpha beta h} prod_edges (1 + (e^beta - 1) delta)."""
    import math

    from test_oracle_exact_tree import exact_marginals
    from graphbix_b200.engine import EM_mod
    rng = np.random.default_rng(seed)
    n = 10
    und = [(1, v) for v in range(2, 7)] + [(int(rng.integers(2, v)), v) for v in range(7, n + 1)]
    und = np.array(und, dtype=np.int64)
    links = np.vstack([und, und[:, ::-1]])
    K = len(links)
    gr = rng.dirichlet(np.ones(q) * 5)
    psi0 = rng.random((K, q))
    psi0 /= psi0.sum(axis=1, keepdims=True)
    # ten vertices and a global field: the synchronous schedule needs damping here (same fixed point)
    out = EM_mod(gr, n, q, links, 1, 5000, 0.0, False, False, False, 0.3, PSIcav0=psi0, BPconvthreshold=1e-14, damping=0.5)
    alpha, beta, PSI = out[1], out[2], out[5]
    if not out[15]:
        pytest.skip("synchronous BP oscillates on this ten-vertex tree even with damping 0.5")
    assert math.isfinite(beta)
    h = PSI.sum(axis=0)
    C = np.ones((q, q)) + (math.exp(beta) - 1) * np.eye(q)
    marg, _ = exact_marginals(n, q, links, gr, C, alpha * beta * np.tile(h, (n, 1)))
    assert np.abs(marg - PSI).max() < 1e-9
