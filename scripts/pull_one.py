"""One configuration of the pull sweep for ncu: python scripts/pull_one.py MODEL Q [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("GBX_PULL", "1")
import numpy as np
import torch
from graphbix_b200.engine import ModEngine, SbmEngine
from graphbix_b200.synth import planted_partition
model, q = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
links, _ = planted_partition(n, 8, 16.0, 0.1, seed=100)
st = torch.cuda.current_stream()
if model == "sbm":
    eng = SbmEngine(n, links, q, stream=st)
    eng.set_options(degreecorrection=1)
    eng.init(np.ones(q) / q, 20 * np.eye(q) + 0.01, None, seed=3)
else:
    eng = ModEngine(n, links, q, stream=st)
    eng.set_options(dc=0)
    eng.init(np.ones(q) / q, None, seed=3)
eng.iterate(12, read_conv=False)
torch.cuda.synchronize()
print("ok", eng.lib.gbx_sbm_sweep_kind(eng._h))
