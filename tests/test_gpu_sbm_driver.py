"""The host program (graphbix_b200.sbm, mirror of sbm.jl's main) end to end on BASELINE.json configs[0]:
planted 4-group SBM, N = 1000, average degree 8, q = 1..6."""
import os

import numpy as np
import pytest

from graphbix_b200.synth import planted_partition

pytestmark = pytest.mark.gpu


def test_sbm_driver_q_sweep(gpu_required, tmp_path, monkeypatch):
    from graphbix_b200 import sbm
    links, lab = planted_partition(1000, 4, 8.0, 0.05, seed=11)
    und = links[links[:, 0] < links[:, 1]]
    path = os.path.join(tmp_path, "edgelist.txt")
    np.savetxt(path, und, fmt="%d", delimiter=" ")
    monkeypatch.chdir(tmp_path)
    out = sbm.main([path, "--q=1:6", "--init=normalizedLaplacian", "--initnum=2", "--dc=false", "--seed=3",
                    "--dataset=planted4"])
    d = os.path.join(tmp_path, "output_planted4")
    for f in ("summary.txt", "assessment.txt", "assignment.txt", "partition4_planted4.smap",
              "ProbDistribution_partition4_planted4.txt"):
        assert os.path.exists(os.path.join(d, f)), f
    summary = open(os.path.join(d, "summary.txt")).read()
    assert "number of vertices (input): 1000" in summary and "q = 4: actual q = 4" in summary
    assess = open(os.path.join(d, "assessment.txt")).read()
    assert assess.startswith("Bethe free energy:\n(q, assessed value)\n1 ")
    assert "\nBayes prediction error:\n" in assess and "\nRescaled affinity matrix C:\n" in assess
    # model assessment: the Bayes prediction error is minimised at the planted q = 4 (or flat beyond it)
    cvb = out["CV"]["Bayes"][:, 0]
    assert cvb[3] < cvb[0] and cvb[3] < cvb[1] and cvb[3] < cvb[2]
    assert cvb[4] > cvb[3] - 0.02 and cvb[5] > cvb[3] - 0.02
    # the q = 4 assignment recovers the planted groups up to a permutation
    asg = np.loadtxt(os.path.join(d, "assignment.txt"), dtype=int)
    assert asg.shape[1] == 7 and asg[:, 0].tolist() == list(range(1, asg.shape[0] + 1))
    n = asg.shape[0]
    conf = np.zeros((4, 4), dtype=int)
    # vertices were renumbered by the giant-component step; compare through the confusion matrix of the largest blocks
    from graphbix_b200 import graphio as gio
    sg = gio.simplegraph(und)
    cc = gio.connected_component_of(sg, 1000, 1)
    np.add.at(conf, (lab[cc - 1], asg[:, 4] - 1), 1)
    assert conf.max(axis=1).sum() > 0.9 * n


def test_mod_driver_q_sweep(gpu_required, tmp_path, monkeypatch):
    """mod.jl's host program on a planted 3-community graph: files written, q* = 3 by the prediction errors."""
    from graphbix_b200 import mod
    links, lab = planted_partition(900, 3, 10.0, 0.05, seed=21, connected_only=True)
    und = links[links[:, 0] < links[:, 1]]
    path = os.path.join(tmp_path, "edgelist.txt")
    np.savetxt(path, und, fmt="%d", delimiter=" ")
    monkeypatch.chdir(tmp_path)
    out = mod.main([path, "--q=2:5", "--initnum=2", "--seed=5", "--alluvial=true"])
    for f in ("summary.txt", "assessment.txt", "assignment.txt", "partition3.smap"):
        assert os.path.exists(os.path.join(tmp_path, f)), f
    summary = open(os.path.join(tmp_path, "summary.txt")).read()
    assert "CV(Bayes Prediction): q* = " in summary and "mean excess degree = " in summary
    assess = open(os.path.join(tmp_path, "assessment.txt")).read()
    for sec in ("Modularity:", "MDL (map equation):", "alpha:", "beta:", "omega (affinity matrix elements):"):
        assert sec in assess
    cvb = out["CV"]["Bayes"][:, 0]
    assert cvb[1] < cvb[0]                      # q = 3 beats q = 2
    assert out["q_star"]["nbt"] in (0, 3)       # non-backtracking estimate: 3 informative eigenvalues (0 = all of them)
    asg = np.loadtxt(os.path.join(tmp_path, "assignment.txt"), dtype=int)
    conf = np.zeros((3, 3), dtype=int)
    np.add.at(conf, (lab, asg[:, 2] - 1), 1)
    assert conf.max(axis=1).sum() > 0.9 * asg.shape[0]
    assert out["modularity"][1] > 0.4
