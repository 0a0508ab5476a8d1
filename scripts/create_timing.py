import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GBX_TIMING"] = "1"
import numpy as np, torch
from graphbix_b200.engine import SbmEngine
from graphbix_b200.synth import planted_partition
for n, deg in ((1000, 8.0), (1_000_000, 16.0)):
    links, _ = planted_partition(n, 8, deg, 0.1, seed=100)
    lp = torch.from_numpy(np.asfortranarray(links)).pin_memory().numpy()
    for rep in range(2):
        t0 = time.perf_counter(); e = SbmEngine(n, lp, 8); torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"n={n} rep={rep} create {1e3*(t1-t0):.2f} ms", file=sys.stderr)
        t0 = time.perf_counter(); e.close(); t1 = time.perf_counter()
        print(f"n={n} rep={rep} destroy {1e3*(t1-t0):.2f} ms", file=sys.stderr)
