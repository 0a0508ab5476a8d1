import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from graphbix_b200.engine import SbmEngine
from graphbix_b200.synth import planted_partition
n, q = int(sys.argv[1]), int(sys.argv[2])
links, _ = planted_partition(n, 8, 16.0, 0.1, seed=100)
outs = []
for rep in range(3):
    eng = SbmEngine(n, links, q)
    eng.set_options(degreecorrection=1)
    eng.init(np.ones(q) / q, 20 * np.eye(q) + 0.01, None, seed=3)
    eng.iterate(3)
    outs.append(eng.state()["PSI"].copy())
    eng.close()
print("max diff run0 vs run1:", np.abs(outs[0] - outs[1]).max(), " run0 vs run2:", np.abs(outs[0] - outs[2]).max(), flush=True)
