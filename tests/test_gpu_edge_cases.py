"""Edge cases of the C ABI on a GPU: malformed input is refused with the documented codes, degenerate graphs run."""
import ctypes as C

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


def _create(lib, n, links, q):
    h = C.c_void_p()
    links = np.asfortranarray(np.asarray(links, dtype=np.int64).reshape(-1, 2))
    rc = lib.gbx_sbm_create(C.byref(h), 0, n, links.shape[0], links.ctypes.data_as(C.POINTER(C.c_int64)), q)
    msg = lib.gbx_last_error().decode() if rc != 0 else ""
    if rc == 0:
        lib.gbx_sbm_destroy(h)
    return rc, msg


def test_malformed_links_are_refused(gpu_required):
    from graphbix_b200 import _lib
    lib = _lib.load()
    ok = [(1, 2), (2, 1), (2, 3), (3, 2)]
    assert _create(lib, 3, ok, 2)[0] == 0
    rc, msg = _create(lib, 3, [(1, 2), (2, 1), (2, 2), (2, 2)], 2)
    assert rc == -1 and "self loop" in msg
    rc, msg = _create(lib, 3, [(1, 2), (2, 1), (1, 2), (2, 1)], 2)
    assert rc == -1 and "duplicate" in msg
    rc, msg = _create(lib, 3, [(1, 2), (2, 3), (3, 2), (2, 3)], 2)
    assert rc == -1
    rc, msg = _create(lib, 3, [(1, 2), (2, 3)], 2)
    assert rc == -1 and "reverse" in msg                       # not bidirected
    rc, msg = _create(lib, 3, [(1, 4), (4, 1)], 2)
    assert rc == -1 and "out of" in msg                        # vertex id beyond n
    rc, msg = _create(lib, 3, [(0, 1), (1, 0)], 2)
    assert rc == -1                                            # ids are 1-based
    assert _create(lib, 3, ok, 0)[0] == -1 and _create(lib, 3, ok, 17)[0] == -1     # q in 1..16
    assert _create(lib, 0, ok, 2)[0] == -1


def test_call_order_is_enforced(gpu_required):
    from graphbix_b200 import _lib
    from graphbix_b200._lib import SbmResult
    lib = _lib.load()
    h = C.c_void_p()
    links = np.asfortranarray(np.array([(1, 2), (2, 1)], dtype=np.int64))
    assert lib.gbx_sbm_create(C.byref(h), 0, 2, 2, links.ctypes.data_as(C.POINTER(C.c_int64)), 2) == 0
    res = SbmResult()
    assert lib.gbx_sbm_em(h, C.byref(res)) == -3               # GBX_ERR_STATE: em before init
    assert lib.gbx_sbm_iterate(h, 1, None) == -3
    assert lib.gbx_sbm_step(h, 1, 0) == -3
    lib.gbx_sbm_destroy(h)
    assert lib.gbx_sbm_destroy(None) == 0


def test_isolated_vertices_and_tiny_graphs(gpu_required):
    """Vertices without links keep the prior as their marginal (logunnormalizedMessage with an empty neighbour list,
    sbm.jl:83-98); a two-vertex graph and a graph whose last vertices are isolated run through EM and the free energy."""
    from graphbix_b200.engine import SbmEngine
    q = 3
    n0, links, psi0, _ = planted_case(300, q, 6.0, 0.1, seed=9)
    n = n0 + 7                                                  # seven isolated vertices at the end
    gr0, C0 = assortative_init(q)
    ref = SyncSBM(n, links, q, gr0, C0, psi0, dc=False, prior=True, lr=0.3)
    eng = SbmEngine(n, links, q)
    eng.set_options(degreecorrection=0)
    eng.init(gr0, C0, psi0)
    for it in range(4):
        ref.iterate()
        eng.iterate(1)
        st = eng.state()
        np.testing.assert_allclose(st["PSI"], ref.PSI, rtol=0, atol=1e-12)
        np.testing.assert_allclose(st["Crs"], ref.C, rtol=1e-11)
    np.testing.assert_allclose(st["PSI"][n0:], np.tile(st["PSI"][n0], (7, 1)), rtol=0, atol=0)   # all equal: prior only
    res = eng.finalize()
    assert np.isfinite([res.FE, res.CVBayes, res.CVGP, res.CVGT, res.CVMAP]).all()
    eng.close()
    # two vertices, one edge
    eng = SbmEngine(2, np.array([(1, 2), (2, 1)]), 2)
    eng.set_options(degreecorrection=1, itrmax=50)
    eng.init(np.ones(2) / 2, np.array([[2.0, 0.5], [0.5, 2.0]]), None, seed=1)
    res = eng.em()
    st = eng.state()
    assert np.isfinite(st["PSI"]).all() and np.allclose(st["PSI"].sum(axis=1), 1.0)
    eng.close()


@pytest.mark.parametrize("q", [1, 16])
def test_extreme_q(gpu_required, q):
    from graphbix_b200.engine import SbmEngine
    n, links, psi0, _ = planted_case(400, max(q, 2), 8.0, 0.2, seed=13)
    psi0 = psi0[:, :q] / psi0[:, :q].sum(axis=1, keepdims=True)
    gr0, C0 = assortative_init(q)
    ref = SyncSBM(n, links, q, gr0, C0, psi0, dc=True, prior=True, lr=0.3)
    eng = SbmEngine(n, links, q)
    eng.init(gr0, C0, psi0)
    for it in range(3):
        ref.iterate()
        eng.iterate(1)
        np.testing.assert_allclose(eng.state()["PSI"], ref.PSI, rtol=0, atol=1e-12)
        np.testing.assert_allclose(eng.messages(), ref.messages_in_link_order(), rtol=0, atol=1e-12)
    eng.close()


@pytest.mark.parametrize("model", ["sbm", "mod"])
def test_queued_iterations_stop_exactly_where_the_one_by_one_loop_stops(gpu_required, monkeypatch, model):
    """Small graphs queue 8 EM iterations before the host reads the residual; the device freezes the state at the
    converging iteration.  Same itrnum and bit-identical state as the loop that synchronises every iteration."""
    from graphbix_b200.engine import EM, EM_mod
    q = 3
    n, links, psi0, _ = planted_case(500, q, 8.0, 0.1, seed=31)
    gr0, C0 = assortative_init(q)
    outs = []
    for batch in ("1", "8", "5"):
        monkeypatch.setenv("GBX_EM_BATCH", batch)
        if model == "sbm":
            out = EM(n, q, links, n, (gr0, C0), False, 300, True, 0.3, PSIcav0=psi0)
            outs.append((out[14], out[13], out[2], out[1], np.array(out[3:12])))
        else:
            out = EM_mod(np.ones(q) / q, n, q, links, 1, 300, 0.0, True, True, False, 0.3, PSIcav0=psi0)
            outs.append((out[16], out[15], out[5], np.array([out[1], out[2], out[3], out[4]]), np.array(out[6:15])))
    assert outs[0][1] and 1 < outs[0][0] < 300
    for o in outs[1:]:
        assert o[0] == outs[0][0] and o[1] == outs[0][1]
        assert np.array_equal(o[2], outs[0][2]) and np.array_equal(o[3], outs[0][3]) and np.array_equal(o[4], outs[0][4])


@pytest.mark.parametrize("model", ["sbm", "mod"])
def test_graph_replay_equals_plain_launches(gpu_required, monkeypatch, model):
    """Small graphs: from the second batch of eight iterations on the EM loop replays a CUDA graph of two iterations whose
    kernels read their iteration number from the device (SbmState::it_now).  Same results, bit for bit, as plain launches
    (GBX_GRAPH=0): the iteration at which the loop stops, the marginals and the free energy."""
    from graphbix_b200.engine import EM, EM_mod
    from graphbix_b200.synth import planted_partition
    links, _ = planted_partition(4000, 3, 9.0, 0.15, seed=17, connected_only=True)
    n = int(links.max())
    outs = []
    for graph in ("1", "0"):
        monkeypatch.setenv("GBX_GRAPH", graph)
        if model == "sbm":
            o = EM(n, 3, links, n, "uniformAssortative", True, 200, True, 0.3, seed=4)
            outs.append((o[2], o[1], o[3], o[13], o[14]))
        else:
            o = EM_mod(np.ones(3) / 3, n, 3, links, 1, 200, 0.0, True, True, True, 0.3, seed=4)
            outs.append((o[5], np.array(o[1:5]), o[6], o[15], o[16]))
    a, b = outs
    assert a[3] == b[3] and a[4] == b[4]
    assert a[4] == 0 or a[4] > 16 or not a[3]          # (the run is long enough for the replay to have been used)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
