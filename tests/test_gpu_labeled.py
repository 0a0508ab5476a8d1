"""Labeled SBM (the reference's notebook model; gbx_lsbm_* / graphbix_b200.labeled): the CUDA path against the numpy
statement of its synchronous schedule sweep by sweep, and against the oracle's fixed points (tests/golden/lsbm_em_*.npz,
asynchronous schedule) with north_star's tolerances."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from sync_model_labeled import SyncLabeled

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "lsbm_em_*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[8:-4] for p in GOLDEN])
@pytest.mark.parametrize("red", [False, True])
def test_labeled_sweeps_match_sync_model(gpu_required, path, red):
    from graphbix_b200.labeled import LabeledEngine
    g = np.load(path)
    n, B, G, dc = int(g["n"]), int(g["B"]), int(g["G"]), bool(g["dc"])
    gr0 = np.ones(B) / B
    ref = SyncLabeled(n, g["links"], B, G, gr0, g["affinity0"], g["psi0"], dc=dc, red=red)
    eng = LabeledEngine(n, g["links"], B, G)
    eng.set_options(dc=int(dc), REDfrac=0.5 if red else 0.0)
    eng.init(gr0, g["affinity0"], g["psi0"])
    st = eng.state()
    np.testing.assert_allclose(st["PSI"], ref.PSI, rtol=0, atol=1e-12)
    np.testing.assert_allclose(st["h"], ref.h, rtol=1e-11)
    for it in range(6):
        cref = ref.iterate()
        cgpu = eng.iterate(1)
        st = eng.state()
        np.testing.assert_allclose(cgpu, cref, rtol=1e-8, atol=1e-15, err_msg=f"conv it={it}")
        np.testing.assert_allclose(st["PSI"], ref.PSI, rtol=0, atol=1e-12, err_msg=f"PSI it={it}")
        np.testing.assert_allclose(eng.messages(), ref.psi, rtol=0, atol=1e-12, err_msg=f"messages it={it}")
        np.testing.assert_allclose(st["gr"], ref.gr, rtol=1e-11, err_msg=f"gr it={it}")
        np.testing.assert_allclose(st["affinity"], ref.aff, rtol=1e-10, err_msg=f"affinity it={it}")
    res = eng.finalize()
    f = ref.finalize()
    for k in ("FE", "CVBayes", "CVGP", "CVGT", "CVMAP", "varCVBayes", "varCVGP", "varCVGT", "varCVMAP"):
        np.testing.assert_allclose(getattr(res, k), f[k], rtol=1e-9, err_msg=k)
    assert res.nonfinite == 0
    eng.close()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[8:-4] for p in GOLDEN])
def test_labeled_fixed_point_matches_oracle(gpu_required, path):
    from graphbix_b200.labeled import EM_labeled
    g = np.load(path)
    n, B, G, dc = int(g["n"]), int(g["B"]), int(g["G"]), bool(g["dc"])
    out = EM_labeled(n, list(g["Larray"]), B, G, g["links"], 4000, float(g["tol"]), True, True, dc, float(g["lr"]), 0,
                     ["a", "d"], 1 / B, PSIcav0=g["psi0"], affinity0=g["affinity0"])
    gr, aff, PSI = out[0], out[1], out[2]
    assert out[12], "CUDA path did not converge"
    assert np.abs(PSI - g["PSI"]).max() < 1e-6                          # marginals: 1e-6 max-abs per vertex
    np.testing.assert_allclose(gr, g["gr"], rtol=1e-6)
    np.testing.assert_allclose(aff, g["affinity"], rtol=1e-6)
    np.testing.assert_allclose(np.array(out[3:12]), g["scalars"], rtol=1e-5)     # FE, CV errors, variances
    assert np.array_equal(np.argmax(PSI, axis=1), np.argmax(g["PSI"], axis=1))    # identical assignments


def test_labeled_input_checks(gpu_required):
    from graphbix_b200 import _lib
    lib = _lib.load()

    def create(links3, G):
        h = C.c_void_p()
        a = np.asfortranarray(np.array(links3, dtype=np.int64))
        rc = lib.gbx_lsbm_create(C.byref(h), 0, 3, a.shape[0], a.ctypes.data_as(C.POINTER(C.c_int64)), 2, G)
        msg = lib.gbx_last_error().decode() if rc else ""
        if rc == 0:
            lib.gbx_lsbm_destroy(h)
        return rc, msg

    assert create([(1, 2, 1), (2, 1, 1), (2, 3, 2), (3, 2, 2)], 2)[0] == 0
    rc, msg = create([(1, 2, 1), (2, 1, 1), (2, 3, 3), (3, 2, 3)], 2)
    assert rc == -1 and "label" in msg
    rc, msg = create([(1, 2, 1), (2, 1, 2)], 2)
    assert rc == -1 and "directions" in msg
    assert create([(1, 2, 1), (2, 1, 1)], 5)[0] == -1                  # more labels than GBX_MAX_LABELS


def test_labeled_queued_iterations_are_exact(gpu_required, monkeypatch):
    """gbx_lsbm_em queues 8 iterations on small graphs; the device freezes the state at the converging one."""
    from graphbix_b200.labeled import EM_labeled
    g = np.load(GOLDEN[0])
    n, B, G, dc = int(g["n"]), int(g["B"]), int(g["G"]), bool(g["dc"])
    outs = []
    for batch in ("1", "8", "3"):
        monkeypatch.setenv("GBX_EM_BATCH", batch)
        outs.append(EM_labeled(n, list(g["Larray"]), B, G, g["links"], 2000, 1e-9, True, True, dc, 0.3, 0, ["a", "d"], 1 / B,
                               PSIcav0=g["psi0"], affinity0=g["affinity0"]))
    assert outs[0][12] and 1 < outs[0][13] < 2000
    for o in outs[1:]:
        assert o[13] == outs[0][13] and o[12]
        # shared-memory atomics make the sums order dependent: equal to rounding, not bit for bit
        np.testing.assert_allclose(o[2], outs[0][2], rtol=0, atol=1e-13)
        np.testing.assert_allclose(o[1], outs[0][1], rtol=1e-12)
        np.testing.assert_allclose(o[3:12], outs[0][3:12], rtol=1e-11)


def test_labeled_main_program_mirror(gpu_required, tmp_path):
    """graphbix_b200.labeled_main: the notebook's MAIN on a planted two-label graph; q* = 2 and the planted groups."""
    from graphbix_b200.labeled_main import run
    from graphbix_b200.synth import labeled_planted_partition
    links, lab = labeled_planted_partition(400, 2, [6.0, 4.0], ["d", "a"], 0.05, seed=8)
    und = links[links[:, 0] < links[:, 1]]
    path = tmp_path / "edges.txt"
    path.write_text("".join(f"{i} {j} {a}\n" for i, j, a in und))
    out = run(str(path), "t", Barray=[2, 3], samples=2, dc=True, REDfrac=0.0, assortativity=["d", "a"], seed=3,
              outdir=str(tmp_path), log=lambda *a: None)
    for f in ("summary_t.txt", "assessment_t.txt", "assignment_t.txt"):
        assert (tmp_path / f).stat().st_size > 0
    assert "q* = " in (tmp_path / "summary_t.txt").read_text()
    assert out["CV"]["Bayes"][0, 0] < 1.2 * out["CV"]["Bayes"][1, 0]
    a = np.loadtxt(tmp_path / "assignment_t.txt", dtype=np.int64)
    ids, block2 = a[:, 0], a[:, 1]
    agree = (block2 - 1 == lab[ids - 1]).mean()
    assert max(agree, 1 - agree) > 0.9


def test_reference_fixture_three_blocks_and_qstar(gpu_required):
    """The data set the reference ships with its notebook (tests/golden/reference_labeledSBMedgelist_900.txt): with the
    notebook's settings EM_labeled lands on the oracle's fixed point for B = 3, recovers the three planted blocks of 300,
    and the Bayes prediction error picks q* = 3 (a clear drop from B = 2, no significant gain at B = 4)."""
    from helpers import best_permutation
    from test_oracle_labeled import _reference_fixture
    from graphbix_b200.labeled import EM_labeled
    s, n, Larray, truth, g, initial_affinity = _reference_fixture()
    K = len(s)
    L = K // 2
    cv, se = {}, {}
    for B in (2, 3, 4):
        best = None
        for seed in range(3):                                  # notebook: `samples` runs, lowest free energy wins
            rng = np.random.default_rng(seed)
            psi0 = rng.random((K, B))
            psi0 /= psi0.sum(1, keepdims=True)
            tol = 1e-12 if (B == 3 and seed == 0) else 1e-6
            out = EM_labeled(n, Larray, B, 2, s, 4000 if tol < 1e-9 else 512, tol, True, True, True, 0.3, 0, ["d", "a"],
                             1 / B, PSIcav0=psi0)
            if B == 3 and seed == 0:
                assert out[12]
                perm = best_permutation(g["PSI"], out[2])
                assert np.abs(out[2][:, perm] - g["PSI"]).max() < 1e-6
                np.testing.assert_allclose(out[0].ravel()[perm], g["gr"].ravel(), rtol=1e-6)
                np.testing.assert_allclose(out[1][:, perm][:, :, perm], g["affinity"], rtol=1e-6, atol=1e-14)
                np.testing.assert_allclose(np.array(out[3:12]), g["scalars"], rtol=1e-5)
                blk = np.argmax(out[2], axis=1)
                conf = np.zeros((3, 3))
                np.add.at(conf, (truth, blk), 1)
                assert conf.max(axis=1).sum() / n > 0.94           # the three blocks of 300
                assert sorted(np.bincount(blk, minlength=3)) [0] > 250
            if out[12] and (best is None or out[3] < best[3]):
                best = out
        assert best is not None
        cv[B], se[B] = best[4], np.sqrt(best[8] / L)
    assert cv[3] < cv[2] - 2 * se[2]                               # B = 3 is clearly better than B = 2 ...
    assert cv[4] > cv[3] - 2 * se[3]                               # ... and B = 4 is no significant improvement: q* = 3


@pytest.mark.parametrize("B", [2, 3])
def test_labeled_batch_equals_one_at_a_time(gpu_required, B):
    """gbx_lsbm_create_batch: several graphs of different sizes in ONE handle (one launch set per EM iteration for all of
    them; graphs converge at different iterations and drop out of the live list) give what a handle of their own gives:
    same itrnum, same state and messages up to the rounding of the atomics."""
    from graphbix_b200.labeled import EM_labeled, EM_labeled_batch
    from graphbix_b200.synth import labeled_planted_partition
    graphs, singles, psi0 = [], [], []
    for k in range(7):
        n = 600 + 100 * k
        links, _ = labeled_planted_partition(n, B, [5.0, 3.0], ["a", "d"], 0.05 + 0.1 * k, seed=40 + k)
        und = links[links[:, 0] < links[:, 1]]
        Larray = [int((und[:, 2] == a + 1).sum()) for a in range(2)]
        graphs.append((n, Larray, B, 2, links))
        m = np.random.default_rng(k).random((links.shape[0], B))
        psi0.append(m / m.sum(axis=1, keepdims=True))
        singles.append(EM_labeled(n, Larray, B, 2, links, 300, 1e-8, True, True, True, 0.3, 0, ["a", "d"], 1 / B,
                                  PSIcav0=psi0[-1], return_engine=True))
    batch, eng, res = EM_labeled_batch(graphs, 300, 1e-8, True, True, True, 0.3, 0, ["a", "d"], PSIcav0=psi0,
                                       return_engine=True)
    assert len(batch) == 7 and eng.count == 7
    assert len({b[13] for b in batch}) > 1                         # they did not all stop at the same iteration
    for g, ((s, seng, sres), b) in enumerate(zip(singles, batch)):
        assert s[12] == b[12] and s[13] == b[13]
        assert sres.itrs_done == res[g].itrs_done
        np.testing.assert_allclose(b[2], s[2], rtol=0, atol=1e-11)
        np.testing.assert_allclose(b[1], s[1], rtol=1e-9)
        np.testing.assert_allclose(b[3:12], s[3:12], rtol=1e-8)
        np.testing.assert_allclose(eng.messages(g), seng.messages(), rtol=0, atol=1e-11)
        seng.close()
    eng.close()


def test_labeled_batch_rejects_bad_input(gpu_required):
    from graphbix_b200.labeled import LabeledBatch
    from graphbix_b200._lib import GbxError
    links = np.array([[1, 2, 1], [2, 1, 1]], dtype=np.int64)
    with pytest.raises(GbxError):
        LabeledBatch([(2, links), (2, np.array([[1, 3, 1], [3, 1, 1]], dtype=np.int64))], 2, 1)   # vertex 3 of a 2-vertex graph
    with pytest.raises(GbxError):
        LabeledBatch([(2, links), (2, np.array([[1, 2, 1], [2, 1, 2]], dtype=np.int64))], 2, 2)   # directions disagree in label


def test_labeled_device_random_messages_same_alone_and_in_batch(gpu_required):
    """gbx_lsbm_init_random: a graph starts from the same device-drawn messages in a handle of its own and inside a batch."""
    from graphbix_b200.labeled import EM_labeled, EM_labeled_batch
    from graphbix_b200.synth import labeled_planted_partition
    graphs = []
    for k in range(3):
        links, _ = labeled_planted_partition(500 + 50 * k, 2, [5.0, 3.0], ["a", "d"], 0.1, seed=7 + k)
        und = links[links[:, 0] < links[:, 1]]
        graphs.append((500 + 50 * k, [int((und[:, 2] == a + 1).sum()) for a in range(2)], 2, 2, links))
    batch = EM_labeled_batch(graphs, 200, 1e-8, True, True, True, 0.3, 0, ["a", "d"], seeds=[11, 12, 13])
    for k, g in enumerate(graphs):
        s = EM_labeled(*g[:4], g[4], 200, 1e-8, True, True, True, 0.3, 0, ["a", "d"], 0.5, seed=11 + k)
        assert s[13] == batch[k][13] and s[12] == batch[k][12]
        np.testing.assert_allclose(batch[k][2], s[2], rtol=0, atol=1e-11)
        assert abs(s[2].sum(axis=1) - 1).max() < 1e-12


def test_labeled_runs_are_bit_reproducible(gpu_required):
    """No floating-point atomics on the labeled path (per-block partial sums, fixed-order reductions): the same batch run
    twice gives identical bits."""
    from graphbix_b200.labeled import EM_labeled_batch
    from graphbix_b200.synth import labeled_planted_partition
    graphs = []
    for k in range(5):
        links, _ = labeled_planted_partition(2000 + 300 * k, 3, [5.0, 3.0], ["a", "d"], 0.15, seed=70 + k)
        und = links[links[:, 0] < links[:, 1]]
        graphs.append((2000 + 300 * k, [int((und[:, 2] == a + 1).sum()) for a in range(2)], 3, 2, links))
    runs = [EM_labeled_batch(graphs, 60, 1e-9, True, True, True, 0.3, 0, ["a", "d"], seeds=[1, 2, 3, 4, 5]) for _ in range(2)]
    for a, b in zip(*runs):
        assert np.array_equal(a[2], b[2]) and np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
        assert a[3:12] == b[3:12] and a[13] == b[13]
