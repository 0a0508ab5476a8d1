#!/usr/bin/env python
"""Benchmark of the graphBIX hot path on B200: EM + BP on a synthetic planted-partition SBM.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the single-GPU configuration the metric is quoted on):
sbm.jl full-affinity SBM, N = 1M vertices, average degree 16, q = 8, FP64 messages.

A "step" is one EM iteration of sbm.jl:342-347: one BP sweep over all K = 2E directed edges, the M-step
(update_Crs) and update_theta.  metric = directed-edge message updates / second = K * steps / time, summed
over ranks (weak scaling: every rank owns an independent graph of the same shape; no data-path collective).

  value      device-resident: graph and messages already in HBM, K steps timed with CUDA events
  e2e        the same metric through the public call a user of the reference makes -- EM(...) with HOST
             buffers: graph build + H2D of links and initial messages + K iterations + free energy / CV
             errors + D2H of PSI, gr, Crs
  roofline   BP-sweep kernel alone (the dominant kernel): algorithmic bytes / CUDA-event time
  cpu_baseline / --impl reference: the CPU restatement of the reference's EM() (oracle/) on the host cores,
             on a bounded sample of the same workload (smaller N, same degree and q).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bp_directed_edge_message_updates_per_sec"
UNIT = "updates/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--deg", type=float, default=16.0)
    ap.add_argument("--q", type=int, default=8)
    ap.add_argument("--eps", type=float, default=0.1)
    ap.add_argument("--dc", type=int, default=1)
    ap.add_argument("--cpu-n", type=int, default=0, help="vertices of the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="auto", choices=["auto", "independent", "partitioned"],
                    help="N > 1: ONE graph of N x n vertices vertex-partitioned over the GPUs is the headline (auto / "
                         "partitioned; the independent-replica number is reported beside it), or one independent graph "
                         "per GPU (independent)")
    ap.add_argument("--part-n", type=int, default=0, help="vertices of the partitioned graph (0 = gpus * n)")
    ap.add_argument("--part-deg", type=float, default=0.0, help="average degree of the partitioned graph (0 = deg)")
    ap.add_argument("--part-shuffle", type=int, default=1,
                    help="1: random vertex numbering (a contiguous partition cuts (P-1)/P of the links); "
                         "0: vertices numbered group by group (the partition follows the planted groups)")
    ap.add_argument("--part-exchange", type=int, default=1,
                    help="1: packed boundary exchange (copy engine pushes, unpack); 0: 64-B peer stores from the sweep")
    ap.add_argument("--no-partitioned-leg", action="store_true")
    ap.add_argument("--no-configs3-leg", action="store_true",
                    help="skip the BASELINE configs[3] leg (N = 50M, degree 20, q = 8 partitioned over the GPUs; N >= 2)")
    ap.add_argument("--configs3-n", type=int, default=50_000_000)
    ap.add_argument("--configs3-deg", type=float, default=20.0)
    ap.add_argument("--configs3-steps", type=int, default=10)
    ap.add_argument("--no-convergence-leg", action="store_true")
    ap.add_argument("--no-batch-leg", action="store_true")
    ap.add_argument("--no-labeled-leg", action="store_true")
    ap.add_argument("--labeled-graphs", type=int, default=512, help="graphs per GPU of the labeled-SBM leg (configs[4]: 4096 over 8 GPUs)")
    ap.add_argument("--batch-graphs", type=int, default=128, help="graphs per GPU of the batched-graphs leg")
    return ap.parse_args()


def workload_name(a):
    return f"sbm.jl full-affinity SBM, planted partition N={a.n}, avg degree {a.deg:g}, q={a.q}, FP64 messages"


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md), streamed every 20 ms."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.proc = index, [], None
        self.marks = []  # (label, sample index) so that only samples taken under load are summarised

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                f = [x.strip() for x in line.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
        except Exception:
            pass

    def mark(self, label):
        self.marks.append((label, len(self.samples)))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)

    def summary(self):
        lo = dict(self.marks).get("start", 0)
        hi = dict(self.marks).get("stop", len(self.samples))
        sel = self.samples[lo:hi] or self.samples
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(s[0]) for s in sel if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in sel)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(sel[0][1]), "reasons": reasons,
                "samples": len(sel)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(a):
    """DRAM bytes per launch of the sweep kernel from the committed ncu capture (profiles/sweep_traffic.json), if it
    was taken on this workload; None otherwise (ncu cannot run inside a timed bench)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "sweep_traffic.json")))
        w = t["workload"]
        if (w["n"], float(w["deg"]), w["q"], w["dc"]) == (a.n, float(a.deg), a.q, a.dc):
            return float(t["dram_bytes_per_launch"]), t["source"]
    except Exception:
        pass
    return None, None


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(a, steps, warmup):
    """The oracle (CPU restatement of the reference's EM loop) on a bounded sample of the workload."""
    from graphbix_b200.synth import planted_partition, random_messages
    try:
        from oracle import sbm_oracle_c as OC
        have_c = OC.available()
    except Exception:
        have_c = False
    n = a.cpu_n or (200_000 if have_c else 1500)
    links, _ = planted_partition(n, a.q, a.deg, a.eps, seed=7, connected_only=True)
    n = int(links.max())
    K = links.shape[0]
    psi0 = random_messages(K, a.q, seed=8)
    gr0 = np.ones(a.q) / a.q
    C0 = 20 * np.eye(a.q) + 0.01 * np.ones((a.q, a.q))
    total = steps + warmup
    if have_c:
        t_per_itr, cores = OC.time_em_iterations(n, a.q, links, psi0, gr0, C0, bool(a.dc), True, 0.3, warmup, steps)
        kind_note = "C port of oracle/sbm_oracle.py (gcc -O2)"
    else:
        from oracle import sbm_oracle as O
        t0 = time.perf_counter()
        O.EM(n, a.q, links, n, (gr0, C0), bool(a.dc), warmup, True, 0.3, PSIcav0=psi0, BPconvthreshold=0.0)
        t1 = time.perf_counter()
        O.EM(n, a.q, links, n, (gr0, C0), bool(a.dc), total, True, 0.3, PSIcav0=psi0, BPconvthreshold=0.0)
        t2 = time.perf_counter()
        t_per_itr = max(((t2 - t1) - (t1 - t0)) / steps, 1e-9)
        cores = 1
        kind_note = "numpy/Python restatement oracle/sbm_oracle.py"
    value = K / t_per_itr
    sample = (f"{steps} EM iterations (sweep + update_Crs + update_theta) on a planted partition N={n}, "
              f"avg degree {a.deg:g}, q={a.q} ({K} directed edges); {kind_note}")
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}, t_per_itr


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(a.steps, 5)
    warmup = min(a.warmup, 1)
    base, t_per_itr = cpu_reference_run(a, steps, warmup)
    line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": t_per_itr * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(a), "sample": base["sample"]},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def partition_parity(a, torch, dist, world, rank, local):
    """Driver-visible correctness of the partitioned path: a small graph run partitioned over ALL ranks against rank 0's
    single-GPU engine from the same initial messages -- marginals (max abs), Crs / gr (max relative), free energy and
    Bayes CV error (relative) after a few EM iterations.  Returned on every rank (the collectives need all of them)."""
    from graphbix_b200.engine import SbmEngine
    from graphbix_b200.engine_dist import PartEngine
    from graphbix_b200.synth import planted_partition, random_messages
    n, q, its = 20_000, a.q, 6
    links, _ = planted_partition(n, q, 12.0, a.eps, seed=777, connected_only=True)
    n = int(links.max())
    psi0 = random_messages(links.shape[0], q, seed=778)
    gr0 = np.ones(q) / q
    C0 = 20 * np.eye(q) + 0.01 * np.ones((q, q))
    part = PartEngine(n, links, q)
    part.set_options(degreecorrection=a.dc, priorlearning=1, learningrate=0.3)
    part.init(gr0, C0, psi0)
    part.iterate(its)
    PSI_p = part.marginals_all()
    sp = part.state()
    rp = part.finalize()
    kind = "pull" if part.pull else "push"
    part.close()
    out = {"n_vertices": n, "q": q, "iterations": its, "ranks": world, "sweep_kernels": kind}
    if rank == 0:
        ref = SbmEngine(n, links, q, device=local)
        ref.set_options(degreecorrection=a.dc, priorlearning=1, learningrate=0.3)
        ref.init(gr0, C0, psi0)
        ref.iterate(its, read_conv=False)
        sr = ref.state()
        rr = ref.finalize()
        ref.close()
        out.update({"parity_max_abs": float(np.abs(PSI_p - sr["PSI"]).max()),
                    "crs_max_rel": float(np.abs(sp["Crs"] / sr["Crs"] - 1).max()),
                    "gr_max_rel": float(np.abs(sp["gr"] / sr["gr"] - 1).max()),
                    "fe_rel": abs(rp.FE / rr.FE - 1), "cvbayes_rel": abs(rp.CVBayes / rr.CVBayes - 1),
                    "reference": "single-GPU engine on rank 0 (push sweep kernels unless GBX_PULL=1)"})
    return out


def run_partitioned(a, torch, dist, world, rank, local, n=None, deg=None, steps=None, warmup=None):
    """ONE graph, vertex-partitioned over the ranks (BASELINE.json configs[3]).  With the pull sweep kernels (default)
    every rank stores the per-vertex records of its vertices into every rank's copy over NVLink from inside the sweep
    kernel and the next sweep gathers locally; the per-iteration collectives are two small NCCL all-reduces (totals,
    S_theta).  GBX_PULL=0: push kernels, boundary MESSAGES packed / copied / unpacked, theta all-gathered."""
    from graphbix_b200.engine_dist import PartEngine
    from graphbix_b200.synth import planted_partition_device
    dev = torch.device("cuda", local)
    n = n or a.part_n or world * a.n
    deg = deg or a.part_deg or a.deg
    steps = steps or a.steps
    warmup = a.warmup if warmup is None else warmup
    t0 = time.perf_counter()
    if rank == 0:
        links, _ = planted_partition_device(n, a.q, deg, a.eps, seed=4242, device=dev, shuffle=bool(a.part_shuffle))
        kk = torch.tensor([links.shape[1]], dtype=torch.int64, device=dev)
    else:
        kk = torch.zeros(1, dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(kk, 0)
    K = int(kk.item())
    if rank != 0:
        links = torch.empty((2, K), dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(links, 0)
    torch.cuda.synchronize()
    chunk = (n + world - 1) // world
    ncut = 0
    for b in range(0, K, 1 << 26):   # in slices: at 10^9 links whole-array temporaries would cost 8 GB each
        sl = links[:, b:b + (1 << 26)]
        ncut += int((((sl[0] - 1) // chunk) != ((sl[1] - 1) // chunk)).sum().item())
    cut = ncut / max(K, 1)
    t_gen = time.perf_counter() - t0
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    eng = PartEngine(n, links, a.q)
    del links
    torch.cuda.empty_cache()
    t_build = time.perf_counter() - t0
    eng.set_options(degreecorrection=a.dc, priorlearning=1, learningrate=0.3, itrmax=steps)
    eng.set_exchange(a.part_exchange)
    gr0 = np.ones(a.q) / a.q
    C0 = 20 * np.eye(a.q) + 0.01 * np.ones((a.q, a.q))
    eng.init(gr0, C0, None, seed=99)
    pull = eng.pull
    eng.iterate(warmup)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    eng.iterate(steps)
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    resid = eng.residual()
    slots = eng.K
    prof = eng.profile_summary()
    if prof and rank == 0:
        print("partitioned step profile (ms per iteration, incl. warm-up iterations):",
              {k: round(v / (steps + warmup), 4) for k, v in prof.items()}, file=sys.stderr)
    eng.close()
    qm = (a.q + 3) // 4 * 4                      # message / record rows are padded to whole 32-B sectors
    if pull:
        # every vertex's record goes to every other rank once per sweep
        boundary = (world - 1) * (n / world) * qm * 8
        exchange = "per-vertex records stored into every rank's copy by the sweep kernel (peer stores over NVLink)"
        what = ("one graph vertex-partitioned over the GPUs, pull sweep kernels: one record per vertex and peer rank crosses "
                "NVLink every sweep; two small NCCL all-reduces (totals, S_theta) per iteration")
    else:
        boundary = cut * K / world * qm * 8      # one message per cut edge and direction
        exchange = ("packed runs pushed by the copy engine, unpacked by the owner" if a.part_exchange else
                    "64-B peer stores from the sweep kernel")
        what = ("one graph vertex-partitioned over the GPUs, push sweep kernels: boundary messages cross NVLink every "
                "sweep, NCCL all-reduce of the totals and all-gather of theta per iteration")
    return {"value": K * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "n_vertices": n,
            "sweep_kernels": "pull" if pull else "push",
            "nvlink_bytes_per_gpu_per_step": boundary,
            "nvlink_egress_gbs_lower_bound": boundary / (ms / steps * 1e-3) / 1e9,
            "nvlink_note": ("bytes that leave each GPU per sweep / whole iteration time: a lower bound of the link rate the "
                            "exchange sustains (peer copy 770 GB/s per direction, B200_PROFILING.md)"),
            "avg_degree": deg, "directed_edges": K, "directed_edges_rank0": slots, "cut_fraction": cut,
            "vertex_numbering": "random" if a.part_shuffle else "by planted group",
            "exchange": exchange, "generate_s": t_gen, "build_s": t_build, "residual_after": resid, "what": what}


def run_convergence(a, torch, dist, world, rank, local):
    """Second half of BASELINE.json's metric, on configs[1]: mod.jl community-restricted SBM on a planted partition
    N = 100k, average degree 10, q = 2..10, one q per GPU (dealt round-robin), each EM() run to BPconv < 1e-6 through
    the public call with host buffers.  Reports iterations and seconds per q."""
    from graphbix_b200.engine import EM_mod
    from graphbix_b200.parallel import my_share
    from graphbix_b200.synth import planted_partition
    n, deg, qtrue = 100_000, 10.0, 4
    links, _ = planted_partition(n, qtrue, deg, 0.1, seed=31, connected_only=True)
    n = int(links.max())
    qs = list(range(2, 11))
    mine = []
    torch.cuda.synchronize()
    t_all = time.perf_counter()
    for pos, q in my_share(qs, rank, world):
        # the first call of a q in a process also loads that q's kernels (CUDA lazy module loading): reported apart
        t0 = time.perf_counter()
        EM_mod(np.ones(q) / q, n, q, links, 1, 256, 0.0, True, True, False, 0.3, seed=5 + q, device=local)
        torch.cuda.synchronize()
        first = time.perf_counter() - t0
        t0 = time.perf_counter()
        out, eng, _ = EM_mod(np.ones(q) / q, n, q, links, 1, 256, 0.0, True, True, False, 0.3, seed=5 + q, device=local,
                             return_engine=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        mine.append({"q": q, "iterations": int(out[16]), "converged": bool(out[15]), "seconds": dt,
                     "seconds_first_call": first, "seconds_em_loop": eng.timings.get("em"),
                     "seconds_graph_build": eng.timings.get("create"), "CVBayes": float(out[7]), "FE": float(out[6])})
        eng.close()
    wall = time.perf_counter() - t_all
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        mine = sorted((x for part in gathered for x in part), key=lambda x: x["q"])
        t = torch.tensor([wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
    return {"workload": f"mod.jl EM() to BPconv < 1e-6, planted partition N={n} (giant component of 100000), avg degree "
                        f"{deg:g}, 4 planted groups, q = 2..10, one q per GPU round-robin",
            "seconds_total": wall, "seconds_q_sweep": sum(r["seconds"] for r in mine) if world == 1 else None,
            "directed_edges": int(links.shape[0]), "runs": mine}


def run_batch(a, torch, dist, world, rank, local):
    """The shape of BASELINE.json configs[4] with the sbm.jl model (the labeled model itself is the `labeled_graphs` leg): a
    detectability sweep over epsilon = c_out / c_in on independent planted-partition graphs of N = 10k, average degree 8,
    2 groups, dealt to the GPUs round-robin, one EM() per graph through the public call (host links in, PSI out).
    Reports graphs per second and the mean overlap with the planted groups per epsilon."""
    from graphbix_b200.engine import EM
    from graphbix_b200.parallel import my_share
    from graphbix_b200.synth import planted_partition
    n, deg, q = 10_000, 8.0, 2
    eps_list = [0.05 * k for k in range(1, 17)]                   # 0.05 .. 0.80; threshold at (sqrt(c)-1)/(sqrt(c)+1) ~ 0.48
    per_eps = max(1, a.batch_graphs // len(eps_list))
    tasks = [(e, s) for e in eps_list for s in range(per_eps * world)]
    mine = my_share(tasks, rank, world)
    graphs = [(pos, eps, planted_partition(n, q, deg, eps, seed=1000 * s + int(eps * 100))) for pos, (eps, s) in mine]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = []
    edges = 0
    for pos, eps, (links, lab) in graphs:
        out = EM(n, q, links, n, "uniformAssortative", False, 256, True, 0.3, seed=pos, device=local)
        block = np.argmax(out[2], axis=1)
        agree = float((block == lab).mean())
        ov = (max(agree, 1 - agree) - 0.5) / 0.5                  # 2-group overlap, 0 = chance
        res.append((eps, ov, int(out[14]), bool(out[13])))
        edges += links.shape[0]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gathered = [None] * world
        dist.all_gather_object(gathered, res)
        res = [x for part in gathered for x in part]
    dt = float(t.item())
    by_eps = {}
    for eps, ov, it, cnv in res:
        by_eps.setdefault(round(eps, 2), []).append((ov, it, cnv))
    return {"workload": f"{len(tasks)} independent planted-partition graphs N={n}, avg degree {deg:g}, q={q}, sbm.jl EM() "
                        f"(no degree correction, itrmax 256), epsilon = 0.05..0.80, dealt to {world} GPU(s)",
            "graphs": len(tasks), "seconds": dt, "graphs_per_sec": len(tasks) / dt,
            "overlap_by_epsilon": {str(k): round(float(np.mean([x[0] for x in v])), 3) for k, v in sorted(by_eps.items())},
            "converged_fraction": float(np.mean([x[3] for x in res])),
            "mean_iterations": float(np.mean([x[2] if x[3] else 256 for x in res]))}


def run_labeled(a, torch, dist, world, rank, local):
    """BASELINE.json configs[4]: the labeled SBM of the reference's notebook on independent N = 10k graphs (two labels:
    one assortative with mean degree 5, one disassortative with mean degree 3; 2 groups), a sweep over epsilon, dealt to
    the GPUs round-robin; the graphs of a GPU run in ONE handle (gbx_lsbm_create_batch: one launch set per EM iteration
    for all of them, graphs drop out as they converge).  Also times the sweep on one large labeled graph (N = 1M, two
    labels of mean degree 8 each, q = 8)."""
    from graphbix_b200.labeled import EM_labeled, EM_labeled_batch, LabeledEngine, initial_affinity
    from graphbix_b200.parallel import my_share
    from graphbix_b200.synth import labeled_planted_partition
    n, B, G = 10_000, 2, 2
    eps_list = [0.05, 0.15, 0.25, 0.35, 0.45, 0.55, 0.65, 0.75]
    per_eps = max(1, a.labeled_graphs // len(eps_list))
    tasks = [(e, s) for e in eps_list for s in range(per_eps * world)]
    graphs = []
    for pos, (eps, s) in my_share(tasks, rank, world):
        links, lab = labeled_planted_partition(n, B, [5.0, 3.0], ["a", "d"], eps, seed=977 * s + int(eps * 100))
        links = np.asfortranarray(links)   # column-major, as the Julia host holds its K x 3 matrix (pageable memory)
        und = links[links[:, 0] < links[:, 1]]
        graphs.append((pos, eps, links, lab, [int((und[:, 2] == k + 1).sum()) for k in range(G)]))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # one at a time (a handle per graph), on a few graphs, for the comparison
    t0 = time.perf_counter()
    sample = graphs[::max(1, len(graphs) // 8)]          # a spread over epsilon: easy and hard graphs alike
    for pos, eps, links, lab, Larray in sample:
        EM_labeled(n, Larray, B, G, links, 512, 1e-6, True, True, True, 0.3, 0, ["a", "d"], 1 / B, seed=pos, device=local)
    torch.cuda.synchronize()
    serial_rate = len(sample) / (time.perf_counter() - t0)
    if world > 1:
        dist.barrier()
    # all graphs of this GPU in one handle; timed: union + CSR build from host edge lists, initial messages (drawn on the
    # device), EM to convergence of every graph, free energy / CV errors, read-back of gr / affinity / PSI of every graph
    t0 = time.perf_counter()
    outs = EM_labeled_batch([(n, Larray, B, G, links) for _, _, links, _, Larray in graphs], 512, 1e-6, True, True, True,
                            0.3, 0, ["a", "d"], seeds=[pos for pos, *_ in graphs], device=local)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # the same batch again: what the next batch of a longer sweep costs (device blocks come from the library's cache)
    t0 = time.perf_counter()
    EM_labeled_batch([(n, Larray, B, G, links) for _, _, links, _, Larray in graphs], 512, 1e-6, True, True, True,
                     0.3, 0, ["a", "d"], seeds=[pos for pos, *_ in graphs], device=local)
    torch.cuda.synchronize()
    dt2 = time.perf_counter() - t0
    res = []
    for (pos, eps, links, lab, Larray), out in zip(graphs, outs):
        agree = float((np.argmax(out[2], axis=1) == lab).mean())
        res.append((eps, (max(agree, 1 - agree) - 0.5) / 0.5, int(out[13]), bool(out[12])))
    t = torch.tensor([dt, dt2], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gathered = [None] * world
        dist.all_gather_object(gathered, res)
        res = [x for part in gathered for x in part]
    dt, dt2 = (float(x) for x in t.tolist())
    by_eps = {}
    for eps, ov, it, cnv in res:
        by_eps.setdefault(round(eps, 2), []).append(ov)
    out = {"workload": f"{len(tasks)} independent labeled graphs N={n}, labels: assortative c=5 + disassortative c=3, "
                       f"q={B}, EM_labeled() with degree correction, itrmax 512, epsilon 0.05..0.75, {world} GPU(s)",
           "graphs": len(tasks), "seconds": dt, "graphs_per_sec": len(tasks) / dt,
           "seconds_second_batch": dt2, "graphs_per_sec_second_batch": len(tasks) / dt2,
           "how": "gbx_lsbm_create_batch: the graphs of a GPU in one handle (disjoint union, grid y = graph), one launch "
                  "set per EM iteration for all live graphs; time includes graph build from host edge lists and read-back",
           "graphs_per_sec_one_at_a_time_per_gpu": serial_rate,
           "overlap_by_epsilon": {str(k): round(float(np.mean(v)), 3) for k, v in sorted(by_eps.items())},
           "converged_fraction": float(np.mean([x[3] for x in res])),
           "mean_iterations": float(np.mean([x[2] if x[3] else 512 for x in res]))}
    if rank == 0:   # device-resident rate of the (untuned) labeled sweep on one large graph
        nb, q = 1_000_000, 8
        links, _ = labeled_planted_partition(nb, q, [8.0, 8.0], ["a", "d"], 0.1, seed=5)
        K = links.shape[0]
        und = links[links[:, 0] < links[:, 1]]
        Larray = [int((und[:, 2] == k + 1).sum()) for k in range(G)]
        eng = LabeledEngine(nb, links, q, G, device=local, stream=torch.cuda.current_stream())
        rng = np.random.default_rng(3)
        psi0 = rng.random((K, q))
        psi0 /= psi0.sum(axis=1, keepdims=True)
        eng.init(np.ones(q) / q, initial_affinity(nb, K, Larray, q, G, True, 0, ["a", "d"], 1 / q), psi0)
        eng.iterate(2, read_conv=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.iterate(5, read_conv=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        eng.close()
        out["large_graph"] = {"n_vertices": nb, "directed_edges": K, "q": q, "labels": G, "ms_per_em_iteration": ms,
                              "updates_per_sec": K / (ms * 1e-3)}
    return out


def run_ours(a):
    import torch
    import torch.distributed as dist
    from graphbix_b200.engine import EM, SbmEngine
    from graphbix_b200.synth import planted_partition, random_messages

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.mode == "auto":
        a.mode = "partitioned" if world > 1 else "independent"
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for --impl ours)")
    torch.cuda.set_device(local)
    if world > 1 or a.mode == "partitioned":
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29541")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), rank=rank, world_size=world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    links, _ = planted_partition(a.n, a.q, a.deg, a.eps, seed=100 + rank, connected_only=False)
    n = a.n
    K = int(links.shape[0])
    gr0 = np.ones(a.q) / a.q
    C0 = 20 * np.eye(a.q) + 0.01 * np.ones((a.q, a.q))       # "uniformAssortative", sbm.jl:290-292

    stream = torch.cuda.current_stream()
    eng = SbmEngine(n, links, a.q, device=local, stream=stream)
    eng.set_options(degreecorrection=a.dc, priorlearning=1, learningrate=0.3, itrmax=a.steps)
    eng.init(gr0, C0, None, seed=1234 + rank)

    # ---- device-resident EM iterations ----
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    eng.iterate(30, read_conv=False)          # bring the clocks up before the W warm-up steps the caller asked for
    eng.iterate(a.warmup, read_conv=False)
    barrier()
    sampler.mark("start")
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    eng.iterate(a.steps, read_conv=False)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - l0
    # ---- the BP sweep kernel alone (roofline) ----
    nsweep = max(10, min(a.steps, 50))
    eng.bp_sweeps(3)
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(stream)
    eng.bp_sweeps(nsweep)
    s1.record(stream)
    torch.cuda.synchronize()
    sweep_ms = s0.elapsed_time(s1) / nsweep
    sampler.mark("stop")
    sampler.stop()
    sweep_bytes = eng.sweep_bytes
    dev_bytes = eng.device_bytes
    sweep_kind = eng.sweep_kind
    eng.close()

    # ---- end to end through EM() with host buffers ----
    # The call a user of the reference makes: EM(Ntot, B, links, ...) -- the initial messages are drawn inside, as in
    # sbm.jl:319-323 (here on the device from a seed), so the host inputs are `links` (pinned, column-major as Julia
    # hands it over) and the outputs PSI, gr, Crs and the scalars.
    links_pin = torch.from_numpy(np.asfortranarray(links)).pin_memory()
    links_h = links_pin.numpy()
    # warm-up: one short call, so that the timed one does not pay CUDA's lazy loading of the free-energy kernels
    EM(n, a.q, links_h, n, (gr0, C0), bool(a.dc), 3, True, 0.3, seed=54 + rank, BPconvthreshold=0.0, device=local)
    barrier()
    t0 = time.perf_counter()
    out, eng2, res2 = EM(n, a.q, links_h, n, (gr0, C0), bool(a.dc), a.steps, True, 0.3, seed=55 + rank,
                         BPconvthreshold=0.0, device=local, return_engine=True)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    e2e_parts = dict(eng2.timings)
    eng2.close()
    # the same call with `links` in ordinary pageable memory, which is what a Julia `ccall` hands over
    links_page = np.asfortranarray(links).copy(order="F")
    # (a short call first, as for the pinned leg: the first staged upload of a process allocates its pinned chunk buffers)
    EM(n, a.q, links_page, n, (gr0, C0), bool(a.dc), 3, True, 0.3, seed=54 + rank, BPconvthreshold=0.0, device=local)
    barrier()
    t0 = time.perf_counter()
    EM(n, a.q, links_page, n, (gr0, C0), bool(a.dc), a.steps, True, 0.3, seed=55 + rank, BPconvthreshold=0.0, device=local)
    torch.cuda.synchronize()
    t_e2e_page = time.perf_counter() - t0
    h2d = links.nbytes
    d2h = out[2].nbytes + out[1].nbytes + out[0].nbytes

    t = torch.tensor([ms, sweep_ms, t_e2e, t_e2e_page], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, sweep_ms, t_e2e, t_e2e_page = (float(x) for x in t.tolist())
    value = world * K * a.steps / (ms * 1e-3)
    e2e_value = world * K * a.steps / t_e2e
    peak, peak_src = measured_peak_gbs()
    achieved = sweep_bytes / (sweep_ms * 1e-3) / 1e9
    traffic, traffic_src = recorded_traffic(a) if sweep_kind == "push" else (None, None)
    part = None
    parity = None
    configs3 = None
    if world > 1 and not a.no_partitioned_leg:
        try:
            parity = partition_parity(a, torch, dist, world, rank, local)
        except Exception as e:  # noqa: BLE001
            parity = {"error": f"{type(e).__name__}: {e}"[:400]}
        if not a.no_configs3_leg:
            try:   # BASELINE configs[3]: N = 50M, degree 20, q = 8 over the GPUs of the box (strong scaling in N_gpus)
                configs3 = run_partitioned(a, torch, dist, world, rank, local, n=a.configs3_n, deg=a.configs3_deg,
                                           steps=a.configs3_steps, warmup=2)
            except Exception as e:  # noqa: BLE001
                configs3 = {"error": f"{type(e).__name__}: {e}"[:400]}
            torch.cuda.empty_cache()
    if (world > 1 and not a.no_partitioned_leg) or a.mode == "partitioned":
        if a.mode == "partitioned":
            part = run_partitioned(a, torch, dist, world, rank, local)
        else:
            # an extra next to the headline: a failure here (every rank fails alike: same code, same sizes) must not
            # cost the independent-graphs line
            try:
                part = run_partitioned(a, torch, dist, world, rank, local)
            except Exception as e:  # noqa: BLE001
                part = {"error": f"{type(e).__name__}: {e}"[:400]}

    conv = None
    if not a.no_convergence_leg:
        try:
            conv = run_convergence(a, torch, dist, world, rank, local)
        except Exception as e:  # noqa: BLE001
            conv = {"error": f"{type(e).__name__}: {e}"[:400]}

    batch = None
    if not a.no_batch_leg:
        try:
            batch = run_batch(a, torch, dist, world, rank, local)
        except Exception as e:  # noqa: BLE001
            batch = {"error": f"{type(e).__name__}: {e}"[:400]}

    labeled = None
    if not a.no_labeled_leg:
        try:
            labeled = run_labeled(a, torch, dist, world, rank, local)
        except Exception as e:  # noqa: BLE001
            labeled = {"error": f"{type(e).__name__}: {e}"[:400]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(a), "directed_edges_per_gpu": K, "degreecorrection": bool(a.dc),
                           "priorlearning": True, "init": "uniformAssortative", "damping": 1.0,
                           "parallelism": f"{world} independent graph(s), one per GPU",
                           "l2_policy": f"inputs larger than L2: {dev_bytes / 1e6:.0f} MB resident per GPU vs 126 MB L2",
                           "step": "one EM iteration = BP sweep + update_Crs + update_theta"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / a.steps,
                        "d2h_bytes_per_step": d2h / a.steps, "seconds": t_e2e, "breakdown_s": e2e_parts,
                        "pageable_value": world * K * a.steps / t_e2e_page, "pageable_seconds": t_e2e_page,
                        "pageable_note": "the same EM() call with `links` in ordinary (unpinned) host memory",
                        "what": "EM() with host buffers: H2D of links, CSR build, random initial messages (device), steps iterations, free energy + CV, D2H of PSI / gr / Crs"},
                "gpu_launches": launches, "clocks": sampler.summary(),
                "roofline": {"bound": "hbm",
                             "kernel": (f"sbm_vertex_kernel<q={a.q},SWEEP> (push BP sweep)" if sweep_kind == "push" else
                                        f"sbm_pull_kernel<q={a.q},SWEEP> (pull BP sweep)"), "achieved": achieved,
                             "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                             "traffic_source": traffic_src, "peak_source": peak_src, "ms_per_launch": sweep_ms, "algorithmic_bytes": sweep_bytes,
                             "sweep_updates_per_sec": K / (sweep_ms * 1e-3)}}
        if conv is not None:
            line["em_convergence"] = conv
        if batch is not None:
            line["batched_graphs"] = batch
        if labeled is not None:
            line["labeled_graphs"] = labeled
        if parity is not None:
            line["partition_parity"] = parity
        if configs3 is not None:
            line["configs3_50M"] = configs3
        if part is not None:
            if parity is not None:
                part["parity_max_abs"] = parity.get("parity_max_abs")
            line["partitioned"] = part
            if a.mode == "partitioned":
                line["independent"] = {"value": value, "ms_per_step": ms / a.steps}
                line["value"], line["ms_per_step"] = part["value"], part["ms_per_step"]
                line["scaling"] = "weak" if part["n_vertices"] == world * a.n else "strong"
                line["config"]["parallelism"] = f"one graph of {part['n_vertices']} vertices partitioned over {world} GPU(s)"
        if not a.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"], _ = cpu_reference_run(a, 2, 1)
            except Exception as e:  # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
