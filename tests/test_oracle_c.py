"""CPU: the two independent restatements of the reference's EM() (oracle/sbm_oracle.py, oracle/sbm_oracle.c)
give the same numbers when they sweep the vertices in the same order, and the C one reproduces the goldens."""
import glob
import os

import numpy as np
import pytest

from oracle import sbm_oracle as O
from oracle import sbm_oracle_c as OC

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sbm_em_*.npz")))
IDS = [os.path.basename(p)[7:-4] for p in GOLDEN]


@pytest.mark.parametrize("dc,prior", [(False, True), (True, True), (True, False), (False, False)])
def test_c_and_python_oracles_agree_sweep_for_sweep(dc, prior):
    g = np.load(GOLDEN[1])
    n, q = int(g["n"]), int(g["q"])
    rng = np.random.default_rng(7)
    perms = [rng.permutation(n) for _ in range(5)]
    kw = dict(PSIcav0=g["psi0"], BPconvthreshold=0.0, perms=perms)
    a = O.EM(n, q, g["links"], n, (g["gr0"], g["C0"]), dc, 5, prior, 0.3, **kw)
    b = OC.EM(n, q, g["links"], n, (g["gr0"], g["C0"]), dc, 5, prior, 0.3, **kw)
    assert np.abs(a.PSI - b.PSI).max() < 1e-12
    assert np.abs(a.PSIcav - b.PSIcav).max() < 1e-12
    np.testing.assert_allclose(b.Crs, a.Crs, rtol=1e-12)
    np.testing.assert_allclose(b.gr, a.gr, rtol=1e-12)
    np.testing.assert_allclose(b.theta, a.theta, rtol=1e-12)
    np.testing.assert_allclose([b.FE, b.CVBayes, b.CVGP, b.CVGT, b.CVMAP, b.varCVBayes, b.varCVGP, b.varCVGT, b.varCVMAP],
                               [a.FE, a.CVBayes, a.CVGP, a.CVGT, a.CVMAP, a.varCVBayes, a.varCVGP, a.varCVGT, a.varCVMAP],
                               rtol=1e-11)
    assert abs(a.BPconv - b.BPconv) < 1e-14


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_c_oracle_reproduces_golden(path):
    g = np.load(path)
    n, q = int(g["n"]), int(g["q"])
    r = OC.EM(n, q, g["links"], n, (g["gr0"], g["C0"]), bool(g["dc"]), 4000, bool(g["prior"]), float(g["lr"]),
              PSIcav0=g["psi0"], BPconvthreshold=float(g["tol"]), seed=99)
    assert r.cnv
    assert np.abs(r.PSI - g["PSI"]).max() < 1e-9
    np.testing.assert_allclose(r.Crs, g["Crs"], rtol=1e-9)
    np.testing.assert_allclose([r.FE, r.CVBayes, r.CVGP, r.CVGT, r.CVMAP], g["scalars"][:5], rtol=1e-8)


def test_q1_matches_the_references_closed_form():
    """sbm.jl:477-492 `UniformGraphError` (degree-uncorrected branch): 1 - log(K / (N*N - 1)).  With one group and
    no degree correction the EM fixed point is Crs = K/N, so CVBayes = 1 - log(K / N^2): the same up to the
    reference's N*N-1 (operator precedence in `Ntot*Ntot-1`)."""
    g = np.load(GOLDEN[1])
    n, links = int(g["n"]), g["links"]
    K = links.shape[0]
    psi0 = np.ones((K, 1))
    # q = 1: PSI = 1 from the start, so BPconv = 0 and the loop stops after one iteration (sbm.jl:348); with
    # learning rate 1 that single M-step already lands on the fixed point Crs = K/N.
    r = OC.EM(n, 1, links, n, (np.ones(1), np.array([[3.0]])), False, 400, True, 1.0, PSIcav0=psi0,
              BPconvthreshold=1e-14)
    assert r.cnv and r.itrnum == 1
    np.testing.assert_allclose(r.Crs[0, 0], K / n, rtol=1e-9)
    np.testing.assert_allclose(r.CVBayes, 1 - np.log(K / n ** 2), rtol=1e-6)
    nullerror = 1 - np.log(K / (n * n - 1))
    assert abs(r.CVBayes - nullerror) < 1e-4
    np.testing.assert_allclose(r.CVGP, r.CVBayes, rtol=1e-6)
    np.testing.assert_allclose(r.CVMAP, r.CVBayes, rtol=1e-6)
