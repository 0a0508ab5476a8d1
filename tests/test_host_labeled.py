"""Host-side pieces of the labeled-SBM mirror (graphbix_b200/labeled.py, labeled_main.py): no GPU needed."""
import numpy as np

from graphbix_b200.labeled import initial_affinity, labeledsimplegraph
from graphbix_b200.labeled_main import (edges_per_label, random_edge_discarding, read_labeled_edgelist, relabel_psi,
                                        uniform_graph_error)
from graphbix_b200.synth import labeled_planted_partition


def test_labeledsimplegraph_and_edgelist_reader(tmp_path):
    p = tmp_path / "e.txt"
    p.write_text("10 20 -1\n20 30 1\n30 10 1\n10 20 -1\n30 30 1\n")
    links, ids = read_labeled_edgelist(str(p))
    assert list(ids) == [10, 20, 30] and links.shape == (5, 3)
    s = labeledsimplegraph(links)
    assert len(s) == 6 and set(s[:, 2]) == {1, 2}                       # labels -1, 1 -> 1, 2; self loop and duplicate gone
    assert {tuple(r) for r in s} == {(1, 2, 1), (2, 1, 1), (2, 3, 2), (3, 2, 2), (3, 1, 2), (1, 3, 2)}
    assert edges_per_label(s) == [1, 2]


def test_initial_affinity_matches_notebook_formula():
    aff = initial_affinity(100, 800, [250, 150], 2, 2, True, 0, ["d", "a"], 0.5)
    cm = 2 * np.array([250, 150]) / 100
    d0, d1 = -0.99 * cm[0] / 0.5, 0.99 * cm[1] / 0.5
    np.testing.assert_allclose(aff[0], ((cm[0] - d0 * 0.5) * np.ones((2, 2)) + d0 * np.eye(2)) / 800)
    np.testing.assert_allclose(aff[1], ((cm[1] - d1 * 0.5) * np.ones((2, 2)) + d1 * np.eye(2)) / 800)
    assert aff[0][0, 0] < aff[0][0, 1] and aff[1][0, 0] > aff[1][0, 1]   # disassortative, assortative
    red = initial_affinity(100, 800, [250, 100, 50], 2, 3, False, 0.5, ["d", "a"], 0.5)
    np.testing.assert_allclose(red[2], (2 * 50 / 100) * np.ones((2, 2)) / 100)   # flat matrix for the neutral label


def test_random_edge_discarding_and_null_error():
    links, _ = labeled_planted_partition(300, 2, [5.0, 3.0], ["a", "d"], 0.1, seed=2)
    L = edges_per_label(links)
    rng = np.random.default_rng(0)
    l2, L2 = random_edge_discarding(links, 2, L, 1, 0.5, rng)
    assert len(l2) == len(links) and L2[2] == -(-L[0] // 2) and L2[0] == L[0] - L2[2] and L2[1] == L[1]
    assert {tuple(r[:2]) for r in l2} == {tuple(r[:2]) for r in links}
    key = {(a, b): c for a, b, c in l2}
    assert all(key[(b, a)] == c for (a, b), c in key.items())          # both directions keep the same label
    n = int(links[:, :2].max())
    e_dc = uniform_graph_error(l2, n, L2, True, 3, sum(L), 0.5)
    e_nodc = uniform_graph_error(l2, n, L2, False, 3, sum(L), 0.5)
    assert np.isfinite([e_dc, e_nodc]).all() and e_nodc > 0


def test_relabel_psi_orders_groups_by_size():
    PSI = np.array([[0.1, 0.9], [0.2, 0.8], [0.7, 0.3]])
    aff = np.arange(8, dtype=float).reshape(2, 2, 2)
    P, g, A = relabel_psi(PSI, np.array([0.3, 0.7]), aff)
    assert np.array_equal(P, PSI[:, ::-1]) and np.array_equal(g, [0.7, 0.3])
    assert np.array_equal(A[0], aff[0][::-1][:, ::-1])
