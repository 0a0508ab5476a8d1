"""mod.jl model: ms per EM iteration (BP sweep + omega / gr update) across q, N = 1M, degree 16, device-resident."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graphbix_b200.engine import ModEngine
from graphbix_b200.synth import planted_partition

n, deg = 1_000_000, 16.0
links, _ = planted_partition(n, 8, deg, 0.1, seed=100)
K = links.shape[0]
for q in (2, 3, 4, 5, 6, 8, 10):
    eng = ModEngine(n, links, q, stream=torch.cuda.current_stream())
    eng.set_options(dc=0)
    eng.init(np.ones(q) / q, None, seed=3)
    eng.iterate(5, read_conv=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.iterate(30, read_conv=False); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    b = K * (2 * 8 * q + 4) + n * (2 * 8 * q + 8)
    print(f"mod q={q}: {ms:.3f} ms per EM iteration  {K / ms * 1e-6:.2f} G upd/s  ~{b / ms * 1e-6:.0f} GB/s algorithmic")
    eng.close()
