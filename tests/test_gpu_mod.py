"""mod.jl path on the GPU: sweep-by-sweep against the numpy statement of the synchronous schedule
(tests/sync_model_mod.py), and converged results against the oracle's goldens (oracle/mod_oracle.py)."""
import glob
import os

import numpy as np
import pytest

from helpers import planted_case
from sync_model_mod import SyncMOD

pytestmark = pytest.mark.gpu

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
