"""Where the time of a small EM() call goes (create / init / em / get_state), first call of a q vs warm."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graphbix_b200.engine import EM, EM_mod
from graphbix_b200.synth import planted_partition

for n in (1000, 100_000):
    links, _ = planted_partition(n, 4, 10.0 if n > 1000 else 8.0, 0.1, seed=31, connected_only=True)
    nn = int(links.max())
    for q in (4, 5):
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out, eng, r = EM_mod(np.ones(q) / q, nn, q, links, 1, 256, 0.0, True, True, False, 0.3, seed=5, return_engine=True)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            print(f"mod n={nn} q={q} rep={rep}: {1e3*(t1-t0):.1f} ms  itr={out[16]} timings={ {k: round(v*1e3,2) for k,v in eng.timings.items()} }")
            eng.close()
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out, eng, r = EM(nn, q, links, nn, "uniformAssortative", True, 256, True, 0.3, seed=5, return_engine=True)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            print(f"sbm n={nn} q={q} rep={rep}: {1e3*(t1-t0):.1f} ms  itr={out[14]} timings={ {k: round(v*1e3,2) for k,v in eng.timings.items()} }")
            eng.close()
