import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphbix_b200.engine import EM
from graphbix_b200.synth import planted_partition
links, _ = planted_partition(1_000_000, 8, 16.0, 0.1, seed=100)
n = 1_000_000
lp = torch.from_numpy(np.asfortranarray(links)).pin_memory().numpy()
gr0 = np.ones(8) / 8; C0 = 20 * np.eye(8) + 0.01
for rep in range(4):
    t0 = time.perf_counter()
    out, eng, r = EM(n, 8, lp, n, (gr0, C0), True, 100, True, 0.3, seed=55, BPconvthreshold=0.0, return_engine=True)
    torch.cuda.synchronize()
    print(f"early={os.environ.get('GBX_EARLY_RESIDUAL','1')} rep={rep} total {time.perf_counter()-t0:.4f}s", {k: round(v, 4) for k, v in eng.timings.items()})
    eng.close()
