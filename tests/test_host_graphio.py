"""CPU: host-side mirror of the reference's graph preparation and output formatting (graphbix_b200/graphio.py,
spectral.py, sbm.py helpers)."""
import os

import numpy as np
import pytest

from graphbix_b200 import graphio as gio
from graphbix_b200 import spectral
from graphbix_b200.synth import planted_partition


def test_simplegraph_matches_reference_semantics():
    # sbm.jl:414-425: bidirect, unique rows (first occurrence order), no self loops
    links = np.array([[1, 2], [2, 1], [2, 3], [3, 3], [1, 2], [4, 1]])
    out = gio.simplegraph(links)
    assert out.tolist() == [[1, 2], [2, 1], [2, 3], [4, 1], [3, 2], [1, 4]]


def test_links_connected_renumbers_like_reference():
    # component {2,3,5} of a 5-vertex graph -> vertices 1,2,3 (sbm.jl:443-475)
    links = gio.simplegraph(np.array([[2, 3], [3, 5], [1, 4]]))
    cc = gio.connected_component_of(links, 5, 2)
    assert cc.tolist() == [2, 3, 5]
    n, out = gio.links_connected(links, 5, cc)
    assert n == 3
    assert sorted(map(tuple, out.tolist())) == [(1, 2), (2, 1), (2, 3), (3, 2)]


def test_giant_component_skips_small_roots():
    # vertex 1 is isolated from the big component: the search moves on to root 2 (sbm.jl:836-847)
    links = gio.simplegraph(np.array([[1, 9], [2, 3], [3, 4], [4, 5], [5, 6], [6, 7], [7, 8], [8, 2]]))
    n, out = gio.giant_component(links, 9, round(9 * 2 / 3), log=lambda *a: None)
    assert n == 7 and out.max() == 7 and out.shape[0] == 14


def test_uniform_graph_error_formulas():
    links = gio.simplegraph(np.array([[1, 2], [2, 3], [3, 1], [3, 4]]))
    K, N = links.shape[0], 4
    assert gio.uniform_graph_error(links, N, False) == pytest.approx(1 - np.log(K / (N * N - 1)))
    d = np.array([2, 2, 3, 1.0])
    assert gio.uniform_graph_error(links, N, True) == pytest.approx(1 + np.log(K) - 2 * np.sum(d * np.log(d)) / K)


def test_q_list_and_julia_number_formatting():
    assert gio.parse_q_list("2:3") == [2, 3]
    assert gio.parse_q_list("2:2:6") == [2, 4, 6]
    assert gio.parse_q_list("2,4,6") == [2, 4, 6]
    assert gio.julia_float(0.5) == "0.5" and gio.julia_float(1e-7) == "1.0e-7" and gio.julia_float(3.0) == "3.0"
    assert gio.julia_float(float("nan")) == "NaN"
    assert gio.julia_float16(0.1 / 3) == "0.03333" and gio.julia_float16(0.1) == "0.1"


def test_sortblock_aligns_labels_with_previous_partition():
    prev = np.array([1, 1, 1, 2, 2, 2])
    now = np.array([2, 2, 2, 1, 1, 3])
    out = gio.sortblock(now, prev)
    assert out[:3].tolist() == [1, 1, 1]          # the block overlapping previous block 1 becomes block 1
    assert set(out[3:].tolist()) == {2, 3}


def test_spectral_initial_condition_recovers_planted_blocks():
    links, lab = planted_partition(600, 3, 12.0, 0.02, seed=5, connected_only=True)
    n = int(links.max())
    gr, Crs = spectral.laplacian_initial(links, n, 3, np.random.default_rng(1))
    assert gr.shape == (3,) and Crs.shape == (3, 3)
    assert abs(gr.sum() - 1) < 1e-12 and np.allclose(Crs, Crs.T)
    # sum_rs gr_r Crs_rs gr_s = average degree (sbm.jl:256: Crs = counts / (N gr' gr))
    assert gr @ Crs @ gr == pytest.approx(links.shape[0] / n)
    assert np.diag(Crs).min() > 3 * (Crs - np.diag(np.diag(Crs))).max()   # assortative structure found


def test_smap_writer_format(tmp_path):
    from graphbix_b200.sbm import alluvial_diagram, assignment_probs
    block = np.array([2, 1, 2, 1])
    maxpsi = np.array([0.9, 0.6, 0.8, 0.99])
    alluvial_diagram(2, block, maxpsi, "toy", str(tmp_path))
    txt = open(os.path.join(tmp_path, "partition2_toy.smap")).read().splitlines()
    assert txt[0] == "*Undirected" and txt[1] == "*Modules 2"
    assert txt[2] == '1 "2,..." 0.05 0.005' and txt[3] == '2 "1,..." 0.05 0.005'
    assert txt[5] == "*Nodes 4"
    assert txt[6] == '1;1 "Node 2" 0.025' and txt[7] == '1:2 "Node 4" 0.025' and txt[8] == '2:1 "Node 1" 0.025'
    assert txt[10] == "*Links 4" and txt[11] == "1 1 0.025"
    assignment_probs(2, np.array([[0.25, 0.75], [1.0, 0.0]]), "toy", str(tmp_path))
    assert open(os.path.join(tmp_path, "ProbDistribution_partition2_toy.txt")).read() == "1 0.25 0.75\n2 1.0 0.0\n"
