"""Host program of `sbm.jl` (src/sbm.jl:694-1067) on top of the CUDA `EM()`: same command line, same output files.

    python -m graphbix_b200.sbm edgelist.txt [--dataset=NAME] [--dc=true] [--q=2:3] [--init=normalizedLaplacian ...]
                                [--initnum=3] [--itrmax=256] [--prior=true] [--learning_rate=0.3] [--alluvial=false]

writes output_<dataset>/summary.txt, assessment.txt, assignment.txt, partition<q>_<dataset>.smap and
ProbDistribution_partition<q>_<dataset>.txt like the reference (the PyPlot figures are produced only when matplotlib
is importable).  Everything here is the reference's host side; the inner loop is `graphbix_b200.engine.EM`.
"""
from __future__ import annotations

import argparse
import math
import os
import sys

import numpy as np

from . import graphio as gio
from .engine import EM

VERSION = "0.2.6-12-julia1.1 (graphbix_b200)"


def _bool(s: str) -> bool:
    return s == "true"  # sbm.jl:758: anything but "true" is false


def parse_args(argv):
    ap = argparse.ArgumentParser(prog="sbm.py", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("filename")
    ap.add_argument("--dataset", default=None)
    ap.add_argument("--dc", default="true")
    ap.add_argument("--q", default="2:3")
    ap.add_argument("--init", action="append", default=None)
    ap.add_argument("--initnum", type=int, default=3)
    ap.add_argument("--itrmax", type=int, default=256)
    ap.add_argument("--prior", default="true")
    ap.add_argument("--learning_rate", type=float, default=0.3)
    ap.add_argument("--alluvial", default="false")
    ap.add_argument("--version", action="version", version=VERSION)
    # not in the reference: where to run and how the synchronous sweep is damped
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--damping", type=float, default=1.0)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--max-retries", type=int, default=20, help="cap on the `cnv == false -> continue` loop of sbm.jl:919")
    a = ap.parse_args(argv)
    inits = []
    for item in (a.init or ["normalizedLaplacian"]):  # --init={a,b} after shell brace expansion, or repeated
        inits += [x for x in item.replace("{", "").replace("}", "").split(",") if x]
    a.init = inits
    return a


def alluvial_diagram(B, block, maxPSIval, dataset, outdir):
    """sbm.jl:548-599."""
    significant = 0.7
    Ntot = len(maxPSIval)
    order = np.argsort(block, kind="stable")
    labels = block[order]
    trueB, labelprev, moduleindices = 0, 0, []
    mod = np.zeros(Ntot, dtype=np.int64)
    sig = np.zeros(Ntot, dtype=np.int64)
    for k in range(Ntot):
        if labels[k] != labelprev:
            trueB += 1
            labelprev = labels[k]
            moduleindices.append(f"{order[k] + 1},...")
        mod[k] = trueB
        if maxPSIval[order[k]] > significant:
            sig[k] = 1
    f16 = gio.julia_float16
    with open(os.path.join(outdir, f"partition{B}_{dataset}.smap"), "w") as fp:
        fp.write("*Undirected\n")
        fp.write(f"*Modules {trueB}\n")
        for md in range(1, trueB + 1):
            fp.write(f'{md} "{moduleindices[md - 1]}" {f16(0.1 / trueB)} {f16(0.01 / trueB)}\n')
        fp.write("*Insignificants 0\n")
        fp.write(f"*Nodes {Ntot}\n")
        sublabel, labelprev = 0, 0
        for k in range(Ntot):
            sublabel = sublabel + 1 if mod[k] == labelprev else 1
            labelprev = mod[k]
            sep = ":" if sig[k] == 1 else ";"
            fp.write(f'{mod[k]}{sep}{sublabel} "Node {order[k] + 1}" {f16(0.1 / Ntot)}\n')
        fp.write(f"*Links {trueB ** 2}\n")
        for s in range(1, trueB + 1):
            for t in range(1, trueB + 1):
                fp.write(f"{s} {t} {f16(0.1 / trueB ** 2)}\n")


def assignment_probs(B, PSIopt, dataset, outdir):
    """sbm.jl:494-508."""
    jf = gio.julia_float
    with open(os.path.join(outdir, f"ProbDistribution_partition{B}_{dataset}.txt"), "w") as fp:
        for k in range(PSIopt.shape[0]):
            fp.write(f"{k + 1}" + "".join(" " + jf(x) for x in PSIopt[k]) + "\n")


def _julia_matrix(M) -> str:
    """How `$(matrix)` prints in Julia: rows separated by ';', entries by spaces."""
    M = np.atleast_2d(M)
    return "[" + "; ".join(" ".join(gio.julia_float(x) for x in row) for row in M) + "]"


def run(a, log=print):
    strdataset = a.filename
    dataset = a.dataset if a.dataset is not None else os.path.splitext(os.path.basename(strdataset))[0]
    Barray = gio.parse_q_list(a.q)
    initialconditions, samples = a.init, a.initnum
    degreecorrection, priorlearning = _bool(a.dc), _bool(a.prior)
    itrmax, learningrate = a.itrmax, float(np.float32(a.learning_rate))  # parse(Float32, ...), sbm.jl:760
    rng = np.random.default_rng(a.seed)
    jf = gio.julia_float

    outdir = f"output_{dataset}"
    os.makedirs(outdir, exist_ok=True)
    fpmeta = open(os.path.join(outdir, "summary.txt"), "w")
    fpmeta.write(f"dataset: {strdataset}\n")
    links, Ntotinput = gio.read_edgelist(strdataset)
    links = gio.simplegraph(links)
    fpmeta.write(f"number of vertices (input): {Ntotinput}\n")
    fpmeta.write(f"number of edges (input, converted to simple graph): {int(0.5 * links.shape[0])}\n")
    Nthreshold = round(Ntotinput * 2 / 3)
    Bsize = len(Barray)
    alluvial = True  # sbm.jl:816: overridden to true in the reference
    fpmeta.write(f"initial conditions: {', '.join(initialconditions)}\n")
    fpmeta.write(f"numbmer of samples for each initial condition: {samples}\n")
    fpmeta.write(f"degree correction: {'true' if degreecorrection else 'false'}\n")
    fpmeta.write(f"maximum number of iteration: {itrmax}\n")
    fpmeta.write(f"cluster size learning: {'true' if priorlearning else 'false'}\n")
    fpmeta.write(f"learning rate: {np.float32(learningrate)}\n")  # a Float32 in the reference (sbm.jl:760)

    Ntot, links = gio.giant_component(links, Ntotinput, Nthreshold, log)
    Ltot = int(0.5 * links.shape[0])
    blockprev = np.zeros(Ntot, dtype=np.int64)
    assignment = np.zeros((Ntot, Bsize + 1), dtype=np.int64)
    assignment[:, 0] = np.arange(1, Ntot + 1)
    fpmeta.write(f"numbmer of vertices (giant component): {Ntot}\n")
    fpmeta.write(f"numbmer of edges (giant component): {Ltot}\n")
    nullerror = gio.uniform_graph_error(links, Ntot, degreecorrection)
    kind = "degree-corrected" if degreecorrection else "degree-uncorrected"
    fpmeta.write(f"Error for q=1 ({kind}): {jf(nullerror)}\n")
    log("BwithNBT = 0")  # spectral = false in the reference (sbm.jl:815)
    fpmeta.write("spectral radius of the non-backtracking matrix = -\n")
    fpmeta.write("non-backtracking matrix: q* = -\n")

    FEvec = np.zeros(Bsize)
    CV = {k: np.zeros((Bsize, 2)) for k in ("Bayes", "GP", "GT", "MAP")}
    grDict, CrsDict = {}, {}
    for bb, B in enumerate(Barray):
        log(f"B = {B}")
        shape = (samples, len(initialconditions))
        FEs = 100 * np.ones(shape)
        CVs = {k: 100 * np.ones(shape) for k in CV}
        varCVs = {k: 100 * np.ones(shape) for k in CV}
        FEmin, itrnumopt = 0.0, 0
        PSIopt, gropt, Crsopt = np.zeros((Ntot, B)), np.zeros((B, 1)), np.zeros((B, B))
        maxCrs = Ntot
        for init, name in enumerate(initialconditions):
            log(f"Initial partition = {name}")
            sm = giveup = overflow = retries = 0
            while sm < samples:
                (gr, Crs, PSI, FE, CVBayes, CVGP, CVGT, CVMAP, vB, vGP, vGT, vMAP, fail, cnv, itrnum) = EM(
                    Ntot, B, links, maxCrs, name, degreecorrection, itrmax, priorlearning, learningrate, rng=rng,
                    damping=a.damping, device=a.device)
                if not cnv:  # sbm.jl:919-922 (the reference retries without bound)
                    retries += 1
                    if retries > a.max_retries:
                        log("BP does not converge: giving up on this initial condition (raise --itrmax or lower --damping)")
                        break
                    continue
                if np.isnan(np.max(gr)) or np.isnan(np.max(Crs)) or np.max(PSI) == np.inf:  # sbm.jl:923-932
                    overflow += 1
                    if overflow > 2:
                        log("overflow occurs too often : Use different initial partition or modifiy maxCrs.")
                        break
                    log("overflow... trying again.")
                    continue
                if fail:  # sbm.jl:933-940
                    giveup += 1
                    if giveup > 2:
                        log("too many fails: give up!")
                        break
                    continue
                sm += 1
                FEs[sm - 1, init] = FE
                for k, v, vv in (("Bayes", CVBayes, vB), ("GP", CVGP, vGP), ("GT", CVGT, vGT), ("MAP", CVMAP, vMAP)):
                    CVs[k][sm - 1, init] = v
                    varCVs[k][sm - 1, init] = vv
                log(f". itr = {itrnum}")
                if FE < FEmin:
                    FEmin, itrnumopt, PSIopt, Crsopt, gropt = FE, itrnum, PSI, Crs, gr
        FEvec[bb] = FEs.min()
        for k in CV:  # standard errors, sbm.jl:966-979 (column-major findmin like Julia)
            idx = np.unravel_index(np.argmin(CVs[k].T), CVs[k].T.shape)
            CV[k][bb, 0] = CVs[k].min()
            CV[k][bb, 1] = math.sqrt(varCVs[k].T[idx] / Ltot)
        grDict[bb], CrsDict[bb] = gropt, Crsopt
        maxPSIval = PSIopt.max(axis=1)
        block = np.argmax(PSIopt, axis=1) + 1
        assignment[:, bb + 1] = block
        fpmeta.write(f"q = {B}: actual q = {len(np.unique(block))}, number of iteration = {itrnumopt}\n")
        block = gio.sortblock(block, blockprev)
        blockprev = block.copy()
        if alluvial:
            alluvial_diagram(B, block, maxPSIval, dataset, outdir)
            assignment_probs(B, PSIopt, dataset, outdir)

    with open(os.path.join(outdir, "assignment.txt"), "w") as fp:
        for i in range(Ntot):
            fp.write(" ".join(str(x) for x in assignment[i]) + "\n")
    with open(os.path.join(outdir, "assessment.txt"), "w") as fp:
        fp.write("Bethe free energy:\n(q, assessed value)\n")
        for bb, B in enumerate(Barray):
            fp.write(f"{B} {jf(FEvec[bb])}\n")
        for title, k in (("Bayes prediction error", "Bayes"), ("Gibbs prediction error", "GP"),
                         ("Gibbs training error", "GT"), ("Gibbs prediction error (MAP estimate)", "MAP")):
            fp.write(f"\n{title}:\n(q, assessed value, standard error)\n")
            for bb, B in enumerate(Barray):
                fp.write(f"{B} {jf(CV[k][bb, 0])} {jf(CV[k][bb, 1])}\n")
        fp.write("\nCluster sizes:\n")
        for bb, B in enumerate(Barray):
            fp.write(f"q = {B}\n{_julia_matrix(np.asarray(grDict[bb]).reshape(1, -1))}\n")
        fp.write("\nRescaled affinity matrix C:\n")
        for bb, B in enumerate(Barray):
            fp.write(f"q = {B}\n{_julia_matrix(CrsDict[bb])}\n")
    fpmeta.close()
    _plots(Barray, FEvec, CV, grDict, CrsDict, dataset, outdir, log)
    return dict(q=Barray, FE=FEvec, CV=CV, assignment=assignment, outdir=outdir)


def _plots(Barray, FEvec, CV, grDict, CrsDict, dataset, outdir, log):
    """sbm.jl:602-686 (PlotAssessments): only when matplotlib is available; the reference uses PyPlot."""
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:
        log("matplotlib not available: assessment plots skipped")
        return
    x = np.array(Barray)
    fig = plt.figure(figsize=(9, 3))
    fig.subplots_adjust(hspace=0.5, wspace=0.3)
    ax = fig.add_subplot(121)
    ax.plot(x, FEvec, color="0.4", marker="8", markersize=8)
    ax.set_ylabel("Bethe free energy")
    ax.set_xlabel("Number of clusters q")
    ax = fig.add_subplot(122)
    for k, c, m in (("GP", "#27AE60", "^"), ("Bayes", "#C0392B", "o"), ("GT", "#2980B9", "D"), ("MAP", "#F39C12", "s")):
        ax.plot(x, CV[k][:, 0], color=c, marker=m, markerfacecolor="white", linewidth=2)
        ax.fill_between(x, CV[k][:, 0] - CV[k][:, 1], CV[k][:, 0] + CV[k][:, 1], color=c, alpha=0.4)
    ax.set_ylabel("Prediction/training error")
    ax.set_xlabel("Number of clusters q")
    fig.suptitle(dataset)
    fig.savefig(os.path.join(outdir, f"assessment_{dataset}.pdf"), bbox_inches="tight", pad_inches=0.1)


def main(argv=None):
    a = parse_args(sys.argv[1:] if argv is None else argv)
    return run(a)


if __name__ == "__main__":
    main()
