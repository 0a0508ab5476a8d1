"""CPU: mod.jl oracle reproduces its goldens under another sweep order; the synchronous statement reaches them."""
import glob
import os

import numpy as np
import pytest

from oracle import mod_oracle as M
from sync_model_mod import SyncMOD

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "mod_em_*.npz")))
IDS = [os.path.basename(p)[7:-4] for p in GOLDEN]
KEYS = ("FE", "CVBayes", "CVGP", "CVGT", "CVMAP", "varCVBayes", "varCVGP", "varCVGT", "varCVMAP")


def test_golden_present():
    assert len(GOLDEN) >= 4


@pytest.mark.parametrize("path", GOLDEN[-1:], ids=IDS[-1:])
def test_mod_oracle_reproduces_golden(path):
    g = np.load(path)
    n, q = int(g["n"]), int(g["q"])
    r = M.EM(g["gr0"], n, q, g["links"], 1, 4000, 0.0, bool(g["learning"]), bool(g["prior"]), bool(g["dc"]),
             float(g["lr"]), rng=np.random.default_rng(4242), PSIcav0=g["psi0"], BPconvthreshold=float(g["tol"]))
    assert r.cnv
    assert np.abs(r.PSI - g["PSI"]).max() < 1e-9
    np.testing.assert_allclose([getattr(r, k) for k in KEYS], g["scalars"], rtol=1e-8)


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_mod_sync_schedule_reaches_oracle_fixed_point(path):
    g = np.load(path)
    n, q = int(g["n"]), int(g["q"])
    s = SyncMOD(n, g["links"], q, g["gr0"], g["psi0"], dc=bool(g["dc"]), learning=bool(g["learning"]),
                prior=bool(g["prior"]), lr=float(g["lr"]))
    assert s.run(4000, float(g["tol"]))
    f = s.finalize()
    assert np.abs(s.PSI - g["PSI"]).max() < 1e-6
    np.testing.assert_allclose([s.excm, s.alpha, s.beta, s.omegain, s.omegaout], g["params"], rtol=1e-6)
    np.testing.assert_allclose([f[k] for k in KEYS], g["scalars"], rtol=1e-5)
