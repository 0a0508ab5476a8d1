"""CUDA path vs the numpy statement of the same synchronous schedule, sweep by sweep (tight tolerances)."""
import numpy as np
import pytest

from helpers import assortative_init, planted_case
from sync_model import SyncSBM

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["push", "pull"])
def sweep_kind(request, monkeypatch):
    """Every test of this file runs with both sweep kernels (push: GBX_PULL=0, pull: GBX_PULL=1; where a run cannot use
    the pull kernels -- damping, q = 9..12 -- the library falls back to push)."""
    monkeypatch.setenv("GBX_PULL", "1" if request.param == "pull" else "0")
    return request.param


def _run_pair(n, links, q, psi0, gr0, C0, iters, **opt):
    from graphbix_b200.engine import SbmEngine
    dc = opt.get("dc", True)
    prior = opt.get("prior", True)
    damping = opt.get("damping", 1.0)
    ref = SyncSBM(n, links, q, gr0, C0, psi0, dc=dc, prior=prior, lr=0.3, damping=damping)
    eng = SbmEngine(n, links, q)
    eng.set_options(degreecorrection=int(dc), priorlearning=int(prior), learningrate=0.3, damping=damping)
    eng.init(gr0, C0, psi0)
    st = eng.state()
    np.testing.assert_allclose(st["PSI"], ref.PSI, rtol=0, atol=1e-13)
    np.testing.assert_allclose(st["gr"], ref.gr, rtol=1e-13)
    for it in range(iters):
        cref = ref.iterate()
        cgpu = eng.iterate(1)
        st = eng.state()
        np.testing.assert_allclose(cgpu, cref, rtol=1e-9, atol=1e-16, err_msg=f"conv it={it}")
        np.testing.assert_allclose(st["PSI"], ref.PSI, rtol=0, atol=1e-12, err_msg=f"PSI it={it}")
        np.testing.assert_allclose(eng.messages(), ref.messages_in_link_order(), rtol=0, atol=1e-12,
                                   err_msg=f"messages it={it}")
        np.testing.assert_allclose(st["Crs"], ref.C, rtol=1e-11, err_msg=f"Crs it={it}")
        np.testing.assert_allclose(st["gr"], ref.gr, rtol=1e-11, err_msg=f"gr it={it}")
        np.testing.assert_allclose(st["h"], ref.h, rtol=1e-11, err_msg=f"h it={it}")
        if dc:
            np.testing.assert_allclose(st["theta"], ref.theta, rtol=1e-12, err_msg=f"theta it={it}")
    return eng, ref


@pytest.mark.parametrize("q", [2, 3, 8])
@pytest.mark.parametrize("dc,prior", [(True, True), (False, True), (True, False), (False, False)])
def test_sweeps_match_sync_model(gpu_required, q, dc, prior):
    n, links, psi0, _ = planted_case(600, q, 9.0, 0.15, seed=q)
    gr0, C0 = assortative_init(q)
    eng, ref = _run_pair(n, links, q, psi0, gr0, C0, 6, dc=dc, prior=prior)
    res = eng.finalize()
    f = ref.finalize()
    for k in ("FE", "CVBayes", "CVGP", "CVGT", "CVMAP", "varCVBayes", "varCVGP", "varCVGT", "varCVMAP"):
        np.testing.assert_allclose(getattr(res, k), f[k], rtol=1e-10, err_msg=k)
    assert res.nonfinite == 0


def test_damped_sweeps_match_sync_model(gpu_required):
    n, links, psi0, _ = planted_case(500, 3, 8.0, 0.1, seed=11)
    gr0, C0 = assortative_init(3)
    _run_pair(n, links, 3, psi0, gr0, C0, 5, damping=0.5)


def _with_hubs(n, q, seed):
    """Planted partition plus two hubs of degree > 256 (own kernel) and a degree-0-free tail."""
    n0, links, psi0, lab = planted_case(n, q, 6.0, 0.1, seed=seed)
    rng = np.random.default_rng(seed)
    und = links[links[:, 0] < links[:, 1]]
    extra = []
    for hub, deg in ((1, 300), (n0 // 2, 700)):
        nb = rng.choice(np.arange(1, n0 + 1), size=deg, replace=False)
        extra += [(min(hub, v), max(hub, v)) for v in nb if v != hub]
    und = np.unique(np.vstack([und, np.array(extra)]), axis=0)
    links = np.vstack([und, und[:, ::-1]])
    K = links.shape[0]
    psi0 = rng.random((K, q))
    psi0[np.arange(K), lab[links[:, 0] - 1]] += 1.0
    psi0 /= psi0.sum(1, keepdims=True)
    return n0, links, psi0


@pytest.mark.parametrize("q", [2, 5, 8, 16])
def test_hub_vertices(gpu_required, q):
    n, links, psi0 = _with_hubs(1500, q, seed=5)
    deg = np.bincount(links[:, 0])
    assert deg.max() > 256
    gr0, C0 = assortative_init(q)
    _run_pair(n, links, q, psi0, gr0, C0, 4, dc=True, prior=True)


@pytest.mark.parametrize("q,dc", [(3, False), (3, True), (5, True), (8, True)])
def test_clamp_regime(gpu_required, q, dc):
    """Tiny affinities push unnormalised values below 1e-8, where the absolute clamp of sbm.jl:42 binds (with degree
    correction the theta factors move the true magnitudes the clamp is applied to)."""
    n, links, psi0, _ = planted_case(400, q, 10.0, 0.1, seed=21)
    gr0 = np.ones(q) / q
    C0 = 1e-3 * np.eye(q) + 1e-5 * np.ones((q, q))
    from graphbix_b200.engine import SbmEngine
    ref = SyncSBM(n, links, q, gr0, C0, psi0, dc=dc, prior=True, lr=0.0)
    eng = SbmEngine(n, links, q)
    eng.set_options(degreecorrection=int(dc), priorlearning=1, learningrate=0.0)
    eng.init(gr0, C0, psi0)
    for it in range(3):
        ref.iterate()
        eng.iterate(1)
        np.testing.assert_allclose(eng.state()["PSI"], ref.PSI, rtol=0, atol=1e-12)
        np.testing.assert_allclose(eng.messages(), ref.messages_in_link_order(), rtol=0, atol=1e-12)


@pytest.mark.parametrize("q", [2, 3, 8])
@pytest.mark.parametrize("log2c", [20, 30, 45])
def test_mixed_plain_and_exponent_tracked_products(gpu_required, q, log2c):
    """The aggregation multiplies a vertex's edge factors directly when degree <= 900 / max|log2 factor| and tracks
    exponents otherwise.  With affinities around 2^log2c and degrees spread over 10..55 both kinds of vertex share
    warps (lanes of one warp in different branches)."""
    n, links, psi0, _ = planted_case(700, q, 30.0, 0.2, seed=40 + q)
    deg = np.bincount(links[:, 0])
    assert deg.min() < 900 / log2c < deg.max() or log2c == 20
    gr0 = np.ones(q) / q
    C0 = (2.0 ** log2c) * np.eye(q) + 1.0
    from graphbix_b200.engine import SbmEngine
    ref = SyncSBM(n, links, q, gr0, C0, psi0, dc=False, prior=True, lr=0.0, maxCrs=2.0 ** 60)
    eng = SbmEngine(n, links, q)
    eng.set_options(degreecorrection=0, priorlearning=1, learningrate=0.0, maxCrs=2.0 ** 60)
    eng.init(gr0, C0, psi0)
    for it in range(4):
        ref.iterate()
        eng.iterate(1)
        np.testing.assert_allclose(eng.state()["PSI"], ref.PSI, rtol=0, atol=1e-12, err_msg=f"PSI it={it}")
        np.testing.assert_allclose(eng.messages(), ref.messages_in_link_order(), rtol=0, atol=1e-12)


@pytest.mark.parametrize("q", [2, 3, 4, 6, 8, 16])
@pytest.mark.parametrize("dc", [True, False])
def test_several_tiles_per_cta(gpu_required, q, dc):
    """More tiles than resident CTAs: every CTA walks several tiles through its double-buffered staging, and its
    M-step accumulators and partial sums carry over from tile to tile."""
    n, links, psi0, _ = planted_case(24000 if q < 16 else 16000, q, 10.0, 0.15, seed=70 + q)
    gr0, C0 = assortative_init(q)
    _run_pair(n, links, q, psi0, gr0, C0, 3, dc=dc, prior=True)


@pytest.mark.parametrize("q,dc", [(2, True), (4, False), (8, True), (8, False), (16, True)])
def test_pull_and_push_sweeps_agree(gpu_required, monkeypatch, q, dc):
    """Two sweep kernels: push (new messages scattered into the receiver's rows; GBX_PULL=0) and pull (a vertex keeps its
    outgoing messages and rebuilds the incoming ones from per-vertex records, gbx_sbm_pull.cuh; GBX_PULL=1).  Same
    schedule, same formulas: the two agree to rounding sweep after sweep, and so do free energy and CV errors."""
    from graphbix_b200.engine import SbmEngine
    n, links, psi0, _ = planted_case(3000, q, 12.0, 0.1, seed=90 + q)
    gr0, C0 = assortative_init(q)
    eng = {}
    for kind in ("pull", "push"):
        monkeypatch.setenv("GBX_PULL", "1" if kind == "pull" else "0")
        e = SbmEngine(n, links, q)
        e.set_options(degreecorrection=int(dc))
        e.init(gr0, C0, psi0)
        assert e.sweep_kind == kind
        eng[kind] = e
    monkeypatch.delenv("GBX_PULL")
    for it in range(8):
        ca, cb = eng["pull"].iterate(1), eng["push"].iterate(1)
        sa, sb = eng["pull"].state(), eng["push"].state()
        np.testing.assert_allclose(ca, cb, rtol=1e-9, atol=1e-16)
        np.testing.assert_allclose(sa["PSI"], sb["PSI"], rtol=0, atol=1e-12, err_msg=f"PSI it={it}")
        np.testing.assert_allclose(sa["Crs"], sb["Crs"], rtol=1e-11)
        np.testing.assert_allclose(sa["theta"], sb["theta"], rtol=1e-13)
        np.testing.assert_allclose(eng["pull"].messages(), eng["push"].messages(), rtol=0, atol=1e-12)
    ra, rb = eng["pull"].finalize(), eng["push"].finalize()
    for k in ("FE", "CVBayes", "CVGP", "CVGT", "CVMAP", "varCVBayes", "varCVGP", "varCVGT", "varCVMAP"):
        np.testing.assert_allclose(getattr(ra, k), getattr(rb, k), rtol=1e-10, err_msg=k)
    for e in eng.values():
        e.close()


@pytest.mark.parametrize("q,dc", [(3, True), (8, False), (8, True)])
def test_exact_mstep_pairing(gpu_required, q, dc):
    """gbx_sbm_options::mstep_exact = 1: update_Crs pairs the NEW messages of both directions (what sbm.jl:128-136 sees
    when it runs after BP()), sweep by sweep equal to the numpy statement of that schedule; and at an itrmax exit of a
    run that has not converged the default (statistics fused into the sweep: new outgoing x consumed incoming) differs
    from it by no more than a few residuals -- at the fixed point the two coincide."""
    from graphbix_b200.engine import SbmEngine
    n, links, psi0, _ = planted_case(2000, q, 9.0, 0.15, seed=30 + q, informative=False)
    gr0, C0 = assortative_init(q)
    ref = SyncSBM(n, links, q, gr0, C0, psi0, dc=dc, prior=True, lr=0.3, mstep_exact=True)
    exact, fused = SbmEngine(n, links, q), SbmEngine(n, links, q)
    exact.set_options(degreecorrection=int(dc), mstep_exact=1)
    fused.set_options(degreecorrection=int(dc), mstep_exact=0)
    exact.init(gr0, C0, psi0)
    fused.init(gr0, C0, psi0)
    its = 14
    for it in range(its):
        ref.iterate()
        ce, cf = exact.iterate(1), fused.iterate(1)
        st = exact.state()
        np.testing.assert_allclose(st["PSI"], ref.PSI, rtol=0, atol=1e-11, err_msg=f"PSI it={it}")
        np.testing.assert_allclose(st["Crs"], ref.C, rtol=1e-10, err_msg=f"Crs it={it}")
    assert max(ce, cf) > 1e-6                                    # an itrmax exit: neither run has converged
    Ce, Cf = exact.state()["Crs"], fused.state()["Crs"]
    assert np.abs(Ce - Cf).max() / np.abs(Ce).max() <= 3.0 * max(ce, cf)
    exact.close()
    fused.close()
    # ... and from a start that leans to the planted labels (one basin for both) they converge to the same parameters
    n, links, psi0, _ = planted_case(2000, q, 9.0, 0.15, seed=30 + q, informative=True)
    out = []
    for flag in (1, 0):
        eng = SbmEngine(n, links, q)
        eng.set_options(degreecorrection=int(dc), mstep_exact=flag, itrmax=3000, bp_conv_threshold=1e-12)
        eng.init(gr0, C0, psi0)
        res = eng.em()
        out.append((res, eng.state()["Crs"]))
        eng.close()
    if out[0][0].cnv and out[1][0].cnv:
        np.testing.assert_allclose(out[0][1], out[1][1], rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(out[0][0].FE, out[1][0].FE, rtol=1e-8)
    else:
        assert dc                   # only the degree-corrected theta cycle (update_theta's argmax) keeps a run from converging
