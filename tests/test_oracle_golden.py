"""CPU: the oracle reproduces its committed golden outputs, and the numpy statement of the synchronous
schedule (tests/sync_model.py, what the CUDA kernels implement) reaches the same fixed points."""
import glob
import os

import numpy as np
import pytest

from oracle import sbm_oracle as O
from sync_model import SyncSBM

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sbm_em_*.npz")))
IDS = [os.path.basename(p)[7:-4] for p in GOLDEN]


def test_golden_present():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN[:2], ids=IDS[:2])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    n, q = int(g["n"]), int(g["q"])
    r = O.EM(n, q, g["links"], n, (g["gr0"], g["C0"]), bool(g["dc"]), 4000, bool(g["prior"]), float(g["lr"]),
             rng=np.random.default_rng(12345), PSIcav0=g["psi0"], BPconvthreshold=float(g["tol"]))
    assert r.cnv
    # a different sweep order (other permutations) must land on the same fixed point
    assert np.abs(r.PSI - g["PSI"]).max() < 1e-9
    np.testing.assert_allclose(r.Crs, g["Crs"], rtol=1e-9)
    np.testing.assert_allclose(np.array([r.FE, r.CVBayes, r.CVGP, r.CVGT, r.CVMAP]), g["scalars"][:5], rtol=1e-8)


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_sync_schedule_reaches_oracle_fixed_point(path):
    g = np.load(path)
    n, q = int(g["n"]), int(g["q"])
    s = SyncSBM(n, g["links"], q, g["gr0"], g["C0"], g["psi0"], dc=bool(g["dc"]), prior=bool(g["prior"]),
                lr=float(g["lr"]))
    assert s.run(4000, float(g["tol"]))
    f = s.finalize()
    assert np.abs(s.PSI - g["PSI"]).max() < 1e-6
    np.testing.assert_allclose(s.gr, g["gr"], rtol=1e-6)
    np.testing.assert_allclose(s.C, g["Crs"], rtol=1e-6)
    got = np.array([f[k] for k in ("FE", "CVBayes", "CVGP", "CVGT", "CVMAP", "varCVBayes", "varCVGP", "varCVGT",
                                   "varCVMAP")])
    np.testing.assert_allclose(got, g["scalars"], rtol=1e-5)
    np.testing.assert_allclose(s.messages_in_link_order(), g["PSIcav"], rtol=0, atol=1e-6)
