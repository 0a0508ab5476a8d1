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
