"""Create (H2D of the edge list + graph build) and PSI read-back with PAGEABLE host arrays, plain copies against the
staged ones (GBX_STAGE_THREADS = 0 / n).  N = 1M, degree 16, q = 8: links 256 MB, PSI 64 MB."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graphbix_b200 import _lib
from graphbix_b200._lib import check
from graphbix_b200.engine import SbmEngine
from graphbix_b200.synth import planted_partition
links, _ = planted_partition(1_000_000, 8, 16.0, 0.1, seed=100)
n, q = 1_000_000, 8
lp = np.asfortranarray(links).copy(order="F")
pin = torch.from_numpy(np.asfortranarray(links)).pin_memory().numpy()
lib = _lib.load()
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
PSI = np.zeros((n, q)); th = np.zeros(n); gr = np.empty(q); Crs = np.empty((q, q)); h = np.empty(q)
gr0 = np.ones(q) / q; C0 = 20 * np.eye(q) + 0.01
for label, arr, nt in (("pinned", pin, 0), ("pageable plain", lp, 0), ("pageable staged 2", lp, 2), ("pageable staged 4", lp, 4),
                       ("pageable staged 6", lp, 6), ("pageable staged 8", lp, 8), ("pageable staged 12", lp, 12)):
    os.environ["GBX_STAGE_THREADS"] = str(nt)
    tc, tg = [], []
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e = SbmEngine(n, arr, q)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        e.init(gr0, C0, None, seed=3)
        e.iterate(1)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        check(lib.gbx_sbm_get_state(e._h, dp(gr), dp(Crs), dp(PSI), dp(th), dp(h)))
        t3 = time.perf_counter()
        e.close()
        tc.append((t1 - t0) * 1e3); tg.append((t3 - t2) * 1e3)
    print(f"{label:20s} create ms {[round(x, 1) for x in tc]}   get_state(PSI 64 MB pageable) ms {[round(x, 1) for x in tg]}", flush=True)
