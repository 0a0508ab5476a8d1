"""BASELINE.json's full sizes, checked through size-independent properties (the oracle cannot run them in seconds),
and a live comparison with the C oracle at the largest size it finishes in a few seconds."""
import numpy as np
import pytest

from helpers import assortative_init, planted_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big_graph():
    from graphbix_b200.synth import planted_partition
    links, lab = planted_partition(1_000_000, 8, 16.0, 0.1, seed=100)
    return 1_000_000, links, lab


def test_configs2_properties_at_full_size(gpu_required, big_graph):
    """sbm.jl, N = 1M, average degree 16, q = 8 (configs[2]): invariants of every EM iteration."""
    from graphbix_b200.engine import SbmEngine
    n, links, _ = big_graph
    q = 8
    gr0, C0 = assortative_init(q)
    deg = np.bincount(links[:, 0] - 1, minlength=n).astype(np.float64)

    def run(iters):
        eng = SbmEngine(n, links, q)
        eng.set_options(degreecorrection=1, priorlearning=1, learningrate=0.3)
        eng.init(gr0, C0, None, seed=77)
        st0 = eng.state()
        conv = eng.iterate(iters)
        return eng, st0, conv

    eng, st0, conv = run(1)
    st1 = eng.state()
    # residual = sum_i sum_s |dPSI_i(s)| / N (sbm.jl:123), recomputed on the host from the two states
    np.testing.assert_allclose(conv, np.abs(st1["PSI"] - st0["PSI"]).sum() / n, rtol=1e-9)
    # gr follows the mean marginal (sbm.jl:120-122 adds the change of the mean; gr started as the mean)
    np.testing.assert_allclose(st1["gr"], st1["PSI"].mean(axis=0), rtol=1e-9)
    eng.iterate(2)
    st = eng.state()
    msg = eng.messages()
    assert np.isfinite(msg).all() and msg.min() >= 0.0
    np.testing.assert_allclose(msg.sum(axis=1), 1.0, rtol=0, atol=1e-12)          # normalize_logprob, sbm.jl:109
    np.testing.assert_allclose(st["PSI"].sum(axis=1), 1.0, rtol=0, atol=1e-12)    # sbm.jl:118
    np.testing.assert_allclose(st["gr"].sum(), 1.0, rtol=1e-12)
    assert np.array_equal(st["Crs"], st["Crs"].T)                                  # update_Crs symmetrises, sbm.jl:139
    assert (st["Crs"] >= 0).all() and (st["Crs"] <= n).all()                       # clamp to [0, maxCrs], sbm.jl:140-146
    # update_theta (sbm.jl:47-64): degree over the mean degree of the vertex's MAP block
    block = np.argmax(st["PSI"], axis=1)
    drmean = np.bincount(block, weights=deg, minlength=q) / np.maximum(np.bincount(block, minlength=q), 1)
    np.testing.assert_allclose(st["theta"], deg / drmean[block], rtol=1e-12)
    assert np.array_equal(eng.assignment(), block + 1)
    # h = sum_i theta_i PSI_i Crs / N enters the NEXT sweep; here: free energy pipeline runs and is finite
    res = eng.finalize()
    vals = [res.FE, res.CVBayes, res.CVGP, res.CVGT, res.CVMAP, res.varCVBayes, res.varCVGP, res.varCVGT, res.varCVMAP]
    assert np.isfinite(vals).all() and res.nonfinite == 0
    # Jensen: Bayes prediction error <= Gibbs prediction error; training error <= prediction error
    assert res.CVBayes <= res.CVGP + 1e-12 and res.CVGT <= res.CVGP + 1e-12
    eng.close()
    # bit-reproducible: fixed-order reductions, no floating-point atomics
    eng2, _, _ = run(3)
    st2 = eng2.state()
    assert np.array_equal(st2["PSI"], st["PSI"]) and np.array_equal(st2["Crs"], st["Crs"])
    eng2.close()


def test_configs1_mod_properties_at_full_size(gpu_required):
    """mod.jl, N = 100k, average degree 10 (configs[1]): q = 2..10, every run converges to normalised marginals and
    the planted 4 groups are recovered at q = 4."""
    from graphbix_b200.engine import EM_mod
    from graphbix_b200.synth import planted_partition
    links, lab = planted_partition(100_000, 4, 10.0, 0.1, seed=31, connected_only=True)
    n = int(links.max())
    lab = lab[:n] if len(lab) >= n else lab
    cv = {}
    for q in range(2, 11):
        out = EM_mod(np.ones(q) / q, n, q, links, 1, 256, 0.0, True, True, False, 0.3, seed=5 + q)
        PSI = out[5]
        assert out[15], f"q={q} did not converge"
        np.testing.assert_allclose(PSI.sum(axis=1), 1.0, rtol=0, atol=1e-12)
        assert np.isfinite(out[6:15]).all()
        assert out[3] > out[4] > 0            # omega_in > omega_out: assortative structure found
        cv[q] = out[7]
        if q == 4:
            block = np.argmax(PSI, axis=1)
            conf = np.zeros((4, 4))
            np.add.at(conf, (lab[:n], block), 1)
            assert conf.max(axis=1).sum() / n > 0.95
    assert cv[4] < cv[2] and cv[4] < cv[3]  # the Bayes prediction error keeps falling up to the planted q
    assert cv[10] > cv[4] - 0.02              # ... and does not improve materially beyond it


def test_live_c_oracle_parity_n5000(gpu_required):
    """The reference's asynchronous schedule (C oracle, built by oracle/Makefile) against the CUDA path on a graph of
    5000 vertices: north_star tolerances (marginals 1e-6, gr / Crs 1e-6 relative, CV errors 1e-5 relative)."""
    from oracle import sbm_oracle_c as OC
    if not OC.available():
        pytest.skip("C oracle not built (python -c 'import __graft_entry__ as g; g.build()')")
    from graphbix_b200.engine import EM
    q = 4
    n, links, psi0, _ = planted_case(5000, q, 10.0, 0.1, seed=17)
    gr0, C0 = assortative_init(q)
    ref = OC.EM(n, q, links, n, (gr0, C0), False, 2000, True, 0.3, PSIcav0=psi0, BPconvthreshold=1e-12,
                float32_logbiprob=False)
    assert ref.cnv
    out, eng, res = EM(n, q, links, n, (gr0, C0), False, 4000, True, 0.3, PSIcav0=psi0, BPconvthreshold=1e-12,
                       return_engine=True)
    assert out[13] and res.nonfinite == 0
    assert np.abs(out[2] - ref.PSI).max() < 1e-6
    np.testing.assert_allclose(out[0].ravel(), ref.gr.ravel(), rtol=1e-6)
    np.testing.assert_allclose(out[1], ref.Crs, rtol=1e-6)
    np.testing.assert_allclose([out[3], out[4], out[5], out[6], out[7]],
                               [ref.FE, ref.CVBayes, ref.CVGP, ref.CVGT, ref.CVMAP], rtol=1e-5)
    assert np.array_equal(eng.assignment(), np.argmax(ref.PSI, axis=1) + 1)
    eng.close()
