"""Phases of a batched labeled-SBM run (BASELINE configs[4] shape): python scripts/labeled_batch_timing.py [graphs]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphbix_b200.labeled import EM_labeled, EM_labeled_batch, LabeledBatch, initial_affinity
from graphbix_b200.synth import labeled_planted_partition
n, B, G = 10_000, 2, 2
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eps_list = [0.05, 0.15, 0.25, 0.35, 0.45, 0.55, 0.65, 0.75]
graphs = []
for k in range(ng):
    eps = eps_list[k % 8]
    links, lab = labeled_planted_partition(n, B, [5.0, 3.0], ["a", "d"], eps, seed=977 * (k // 8) + int(eps * 100))
    links = np.asfortranarray(links)   # column-major, as the Julia host holds it
    und = links[links[:, 0] < links[:, 1]]
    graphs.append((n, [int((und[:, 2] == a + 1).sum()) for a in range(G)], B, G, links))
for rep in range(3):
    torch.cuda.synchronize(); t = [time.perf_counter()]
    eng = LabeledBatch([(g[0], g[4]) for g in graphs], B, G)
    t.append(time.perf_counter())
    eng.set_options(dc=1, learning=1, priorlearning=1, learningrate=0.3, itrmax=512, bp_conv_threshold=1e-6)
    aff = [initial_affinity(g[0], g[4].shape[0], g[1], B, G, True, 0, ["a", "d"], 1 / B) for g in graphs]
    t.append(time.perf_counter())
    eng.init(np.tile(np.ones(B) / B, (ng, 1)), np.stack(aff), None, seeds=list(range(ng)))
    t.append(time.perf_counter())
    res = eng.em()
    t.append(time.perf_counter())
    sts = [eng.state(g) for g in range(ng)]
    t.append(time.perf_counter())
    eng.close()
    d = np.diff(t)
    print(f"batch of {ng}: create {d[0]:.3f} host-init {d[1]:.3f} init {d[2]:.3f} em {d[3]:.3f} state {d[4]:.3f} total {t[-1]-t[0]:.3f} s"
          f" = {ng / (t[-1]-t[0]):.1f} graphs/s (em only {ng / d[3]:.1f}); converged {np.mean([r.cnv for r in res]):.2f}, "
          f"mean itr {np.mean([r.itrnum if r.cnv else 512 for r in res]):.0f}", flush=True)
t0 = time.perf_counter()
for k in range(4):
    EM_labeled(*graphs[k][:4], graphs[k][4], 512, 1e-6, True, True, True, 0.3, 0, ["a", "d"], 1 / B, rng=np.random.default_rng(k))
print("one at a time:", 4 / (time.perf_counter() - t0), "graphs/s")
