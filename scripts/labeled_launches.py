"""A few EM iterations of the labeled SBM on one large graph (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphbix_b200.labeled import LabeledEngine, initial_affinity
from graphbix_b200.synth import labeled_planted_partition
nb, q, G = 1_000_000, 8, 2
links, _ = labeled_planted_partition(nb, q, [8.0, 8.0], ["a", "d"], 0.1, seed=5)
K = links.shape[0]
und = links[links[:, 0] < links[:, 1]]
Larray = [int((und[:, 2] == k + 1).sum()) for k in range(G)]
eng = LabeledEngine(nb, links, q, G)
rng = np.random.default_rng(3)
psi0 = rng.random((K, q)); psi0 /= psi0.sum(axis=1, keepdims=True)
eng.init(np.ones(q) / q, initial_affinity(nb, K, Larray, q, G, True, 0, ["a", "d"], 1 / q), psi0)
print(eng.iterate(3))
