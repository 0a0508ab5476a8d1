"""Push / pull sweep on graphs WITH locality: vertex j is linked to vertices within a window |i - j| <= W of the vertex
numbering (what a bandwidth-reducing reordering of a real graph gives), against the bench's planted partition, whose
numbering is random.  The scatter `nxt[rev[e]]` then stays inside windows of W x degree x 64 bytes.
    python scripts/locality_sweep.py [W ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graphbix_b200.engine import SbmEngine
from graphbix_b200.synth import planted_partition

n, deg, q = 1_000_000, 16, 8
st = torch.cuda.current_stream()


def window_graph(W, seed=1):
    rng = np.random.default_rng(seed)
    m = n * deg // 2
    u = rng.integers(0, n, size=int(m * 1.05))
    d = rng.integers(1, W + 1, size=u.size) * rng.choice([-1, 1], size=u.size)
    v = u + d
    ok = (v >= 0) & (v < n)
    u, v = u[ok], v[ok]
    lo, hi = np.minimum(u, v), np.maximum(u, v)
    key = np.unique(lo.astype(np.int64) * n + hi)[:m]
    lo, hi = key // n, key % n
    und = np.stack([lo + 1, hi + 1], axis=1)
    return np.vstack([und, und[:, ::-1]]).astype(np.int64)


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); fn(reps); e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(name, links):
    K = links.shape[0]
    out = []
    for kind in ("push", "pull"):
        os.environ["GBX_PULL"] = "1" if kind == "pull" else "0"
        eng = SbmEngine(n, links, q, stream=st)
        eng.set_options(degreecorrection=1)
        eng.init(np.ones(q) / q, 20 * np.eye(q) + 0.01, None, seed=3)
        eng.iterate(8, read_conv=False)
        sw = timed(lambda r: eng.lib.gbx_sbm_bp_sweeps(eng._h, r), 20)
        it = timed(lambda r: eng.iterate(r, read_conv=False), 20)
        b = float(eng.lib.gbx_sbm_sweep_bytes(eng._h))
        out.append(f"{kind}: sweep {sw * 1e3:.0f} us = {b / sw * 1e-6:.0f} GB/s ({b / sw * 1e-6 / 6544.3:.3f} of peak), EM iteration {it:.3f} ms = {K / it * 1e-6:.1f} G updates/s")
        eng.close()
    print(f"{name:34s} K={K} | " + " | ".join(out), flush=True)


links, _ = planted_partition(n, 8, float(deg), 0.1, seed=100)
run("planted partition, random numbering", links)
for W in [int(x) for x in sys.argv[1:]] or [1 << 18, 1 << 14, 1 << 10]:
    run(f"window graph W = {W}", window_graph(W))
