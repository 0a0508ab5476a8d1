"""Pull vs push sweep kernels across q, sbm.jl and mod.jl models, N = 1M, degree 16, device-resident:
ms per EM iteration and per BP sweep (CUDA events).   python scripts/qsweep_kinds.py [q ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graphbix_b200.engine import ModEngine, SbmEngine
from graphbix_b200.synth import planted_partition

n, deg = int(os.environ.get("N", 1_000_000)), float(os.environ.get("DEG", 16.0))
qs = [int(x) for x in sys.argv[1:]] or [2, 3, 4, 5, 6, 8, 16]
links, _ = planted_partition(n, 8, deg, 0.1, seed=100)
K = links.shape[0]
st = torch.cuda.current_stream()


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); fn(reps); e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for q in qs:
    for model in ("sbm", "mod"):
        out = []
        for kind in ("pull", "push"):
            os.environ["GBX_PULL"] = "1" if kind == "pull" else "0"
            if model == "sbm":
                eng = SbmEngine(n, links, q, stream=st)
                eng.set_options(degreecorrection=1)
                eng.init(np.ones(q) / q, 20 * np.eye(q) + 0.01, None, seed=3)
            else:
                eng = ModEngine(n, links, q, stream=st)
                eng.set_options(dc=0)
                eng.init(np.ones(q) / q, None, seed=3)
            kind_used = "pull" if eng.lib.gbx_sbm_sweep_kind(eng._h) else "push"
            eng.iterate(8, read_conv=False)
            it_ms = timed(lambda r: eng.iterate(r, read_conv=False), 20)
            sw_ms = timed(lambda r: eng.lib.gbx_sbm_bp_sweeps(eng._h, r), 20) if model == "sbm" else float("nan")
            b = float(eng.lib.gbx_sbm_sweep_bytes(eng._h))
            out.append(f"{kind_used}: iter {it_ms:.3f} ms ({K / it_ms * 1e-6:.1f} G/s) sweep {sw_ms * 1e3:.0f} us ({b / sw_ms * 1e-6:.0f} GB/s)")
            eng.close()
        print(f"q={q} {model} | " + " | ".join(out), flush=True)
