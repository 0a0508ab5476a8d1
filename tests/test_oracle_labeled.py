"""oracle/labeled_oracle.py (the labeled SBM of the reference's notebook; DESIGN.md row (f'), CUDA path not built yet).

The reference has no fixtures for it, so the transcription is pinned against the sbm.jl oracle: with one label, no
degree correction and the parameters frozen, both describe the same model (Crs = Ntot * affinity) and must reach the
same BP fixed point and the same cross-validation errors."""
import numpy as np

from helpers import assortative_init, planted_case
from oracle import labeled_oracle as L
from oracle import sbm_oracle as O


def test_one_label_reduces_to_sbm_oracle():
    q = 3
    n, links, psi0, _ = planted_case(150, q, 7.0, 0.15, seed=4)
    gr0, C0 = assortative_init(q, cin=12.0, cout=0.8)
    rng = np.random.default_rng(1)
    perms = [rng.permutation(n) for _ in range(400)]
    r = O.EM(n, q, links, n, (gr0, C0), False, 400, False, 0.0, PSIcav0=psi0, BPconvthreshold=1e-12, perms=perms,
             float32_logbiprob=False)
    assert r.cnv
    # without prior learning sbm.jl keeps gr = mean of the initial marginals (sbm.jl:341) through the loop and only
    # refreshes it at the very end (sbm.jl:365); a run of zero iterations returns exactly that vector
    gr_run = O.EM(n, q, links, n, (gr0, C0), False, 0, False, 0.0, PSIcav0=psi0, float32_logbiprob=False).gr
    links3 = np.column_stack([links, np.ones(len(links), dtype=np.int64)])
    Ltot = len(links) // 2
    lr = L.EM(n, [Ltot], q, 1, links3, 400, 1e-12, False, False, False, 0.3, 0, ["a"], 1 / q, PSIcav0=psi0,
              affinity0=(r.Crs / n)[None], gr0=gr_run, perms=perms)
    assert lr.cnv
    assert np.abs(lr.PSI - r.PSI).max() < 1e-8
    assert np.abs(lr.PSIcav - r.PSIcav).max() < 1e-8
    np.testing.assert_allclose([lr.CVBayes, lr.CVGP, lr.CVGT, lr.CVMAP], [r.CVBayes, r.CVGP, r.CVGT, r.CVMAP], rtol=1e-8)
    np.testing.assert_allclose([lr.varCVBayes, lr.varCVGP], [r.varCVBayes, r.varCVGP], rtol=1e-6)


def test_two_labels_recover_planted_groups_and_learn_their_structure():
    B = 2
    links, lab = L.labeled_planted_partition(240, B, [5.0, 4.0], ["a", "d"], eps=0.08, seed=3)
    n = int(links[:, :2].max())
    und = links[links[:, 0] < links[:, 1]]
    Larray = [int((und[:, 2] == a + 1).sum()) for a in range(2)]
    rng = np.random.default_rng(5)
    K = len(links)
    psi0 = rng.random((K, B)) * 0.5
    psi0[np.arange(K), lab[links[:, 0] - 1]] += 1.0
    psi0 /= psi0.sum(1, keepdims=True)
    r = L.EM(n, Larray, B, 2, links, 300, 1e-7, True, True, True, 0.3, 0, ["a", "d"], 1 / B, rng=rng, PSIcav0=psi0)
    assert r.cnv
    np.testing.assert_allclose(r.PSI.sum(axis=1), 1.0, atol=1e-12)
    np.testing.assert_allclose(r.PSIcav.sum(axis=1), 1.0, atol=1e-12)
    np.testing.assert_allclose(r.gr.sum(), 1.0, atol=1e-9)
    for a in range(2):
        np.testing.assert_allclose(r.affinity[a], r.affinity[a].T, rtol=1e-12)
    block = np.argmax(r.PSI, axis=1)
    agree = (block == lab[:n]).mean()
    assert max(agree, 1 - agree) > 0.9
    # label 1 was planted assortative, label 2 disassortative: the learned matrices say so
    assert np.diag(r.affinity[0]).min() > r.affinity[0][0, 1]
    assert np.diag(r.affinity[1]).max() < r.affinity[1][0, 1]
    assert np.isfinite([r.FE, r.CVBayes, r.CVGP, r.CVGT, r.CVMAP]).all()
    assert r.CVBayes <= r.CVGP + 1e-12


def test_labeled_golden_fixtures_reproduce():
    """tests/golden/lsbm_em_*.npz (make_golden.py labeled) are the fixed points the CUDA path of this model will be
    checked against; the oracle must still reproduce them (different sweep order, same fixed point)."""
    import glob
    import os
    paths = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "lsbm_em_*.npz")))
    assert len(paths) >= 3
    g = np.load(paths[0])
    n, B, G = int(g["n"]), int(g["B"]), int(g["G"])
    r = L.EM(n, list(g["Larray"]), B, G, g["links"], 4000, 1e-12, True, True, bool(g["dc"]), float(g["lr"]), 0,
             ["a", "d"], 1 / B, rng=np.random.default_rng(999), PSIcav0=g["psi0"], affinity0=g["affinity0"])
    assert r.cnv
    assert np.abs(r.PSI - g["PSI"]).max() < 1e-8
    np.testing.assert_allclose(r.affinity, g["affinity"], rtol=1e-7)
    np.testing.assert_allclose([r.FE, r.CVBayes, r.CVGP, r.CVGT, r.CVMAP], g["scalars"][:5], rtol=1e-8)


def _reference_fixture():
    """The edge list the reference ships (src/labeled_sbm/labeledSBMedgelist_Nblock=[300,300,300]_c=[5,3]_X=[0.1,0.9]_1.txt,
    copied verbatim): 900 vertices, planted blocks 1-300 / 301-600 / 601-900, label 1 disassortative, label 2 assortative."""
    import os
    from graphbix_b200.labeled import initial_affinity, labeledsimplegraph
    from graphbix_b200.labeled_main import edges_per_label, read_labeled_edgelist
    here = os.path.join(os.path.dirname(__file__), "golden")
    links, ids = read_labeled_edgelist(os.path.join(here, "reference_labeledSBMedgelist_900.txt"))
    s = labeledsimplegraph(links)
    truth = (np.asarray(ids) - 1) // 300
    return s, len(ids), edges_per_label(s), truth, np.load(os.path.join(here, "lsbm_reference_fixture_B3.npz")), initial_affinity


def test_reference_fixture_oracle_golden_and_sync_schedule():
    from helpers import best_permutation, sha
    from sync_model_labeled import SyncLabeled
    s, n, Larray, truth, g, initial_affinity = _reference_fixture()
    assert n == 900 and list(np.bincount(truth)) == [300, 300, 300] and sha(s) == str(g["links_sha"])
    assert Larray == list(g["Larray"]) == [2246, 1358]
    # the oracle's fixed point recovers the planted blocks
    blk = np.argmax(g["PSI"], axis=1)
    conf = np.zeros((3, 3))
    np.add.at(conf, (truth, blk), 1)
    assert conf.max(axis=1).sum() / n > 0.94
    # and the synchronous schedule (what gbx_lsbm.cu runs) reaches it from the same random start
    K, B = len(s), 3
    rng = np.random.default_rng(0)
    psi0 = rng.random((K, B))
    psi0 /= psi0.sum(1, keepdims=True)
    assert sha(psi0) == str(g["psi0_sha"])
    m = SyncLabeled(n, s, B, 2, np.ones(B) / B, initial_affinity(n, K, Larray, B, 2, True, 0, ["d", "a"], 1 / B), psi0)
    for _ in range(600):
        if m.iterate() < 1e-12:
            break
    perm = best_permutation(g["PSI"], m.PSI)
    assert np.abs(m.PSI[:, perm] - g["PSI"]).max() < 1e-6
    f = m.finalize()
    np.testing.assert_allclose([f["FE"], f["CVBayes"], f["CVGP"], f["CVGT"], f["CVMAP"]], g["scalars"][:5], rtol=1e-5)
