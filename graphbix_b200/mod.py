"""Host program of `mod.jl` (src/mod.jl:609-1018) on top of the CUDA `EM()` of the community-restricted SBM.

    python -m graphbix_b200.mod edgelist.txt [--dataset=NAME] [--q=2:6] [--dc=true] [--learning=true] [--prior=false]
                                [--initnum=10] [--itrmax=256] [--learningrate=0.3] [--alluvial=false]

writes summary.txt, assessment.txt, assignment.txt (and partition<q>.smap with --alluvial=true) into the current
directory like the reference.  The inner loop is `graphbix_b200.engine.EM_mod`; everything here is host side.
"""
from __future__ import annotations

import argparse
import math
import os
import sys

import numpy as np

from . import graphio as gio
from .engine import EM_mod

VERSION = "0.1.6.1 (graphbix_b200)"


def parse_args(argv):
    ap = argparse.ArgumentParser(prog="mod.py", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("filename")
    ap.add_argument("--dataset", default="")
    ap.add_argument("--q", default="2:6")
    ap.add_argument("--dc", default="true")
    ap.add_argument("--learning", default="true")
    ap.add_argument("--prior", default="false")
    ap.add_argument("--initnum", type=int, default=10)
    ap.add_argument("--itrmax", type=int, default=256)
    ap.add_argument("--learningrate", type=float, default=0.3)
    ap.add_argument("--alluvial", default="false")
    ap.add_argument("--version", action="version", version=VERSION)
    ap.add_argument("--device", type=int, default=0)       # not in the reference
    ap.add_argument("--damping", type=float, default=1.0)  # not in the reference
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--spectral", default="true", help="non-backtracking estimate of q* (mod.jl:720)")
    return ap.parse_args(argv)


def retrieval_q(links, degrees, block):
    """mod.jl:348-367 (modularity of the partition)."""
    Ktot = links.shape[0]
    sel = links[:, 0] >= links[:, 1]
    same = block[links[sel, 0] - 1] == block[links[sel, 1] - 1]
    bd = np.bincount(block - 1, weights=degrees, minlength=int(block.max()))
    return float(same.sum()) * 2 / Ktot - float(bd @ bd) / Ktot ** 2


def minimum_description_length(links, degrees, block):
    """mod.jl:370-407 (MDL of the map equation)."""
    Ktot = links.shape[0]
    B = int(block.max())
    sel = links[:, 0] >= links[:, 1]
    bi, bj = block[links[sel, 0] - 1], block[links[sel, 1] - 1]
    same = bi == bj
    lin = np.bincount(bi[same] - 1, minlength=B)
    lout = np.bincount(bi[~same] - 1, minlength=B) + np.bincount(bj[~same] - 1, minlength=B)
    cut = int((~same).sum())
    MDL = 0.0
    for k in range(B):
        if lin[k] + lout[k] == 0:
            continue
        MDL += (lin[k] + lout[k]) * math.log2(lin[k] + lout[k])
        if lout[k] == 0:
            continue
        MDL -= lout[k] * math.log2(lout[k])
    MDL = 2 * MDL + Ktot
    if cut > 0:
        MDL += 2 * cut * (1 + math.log2(2 * cut))
    d = degrees[degrees > 0]
    MDL -= float(np.sum(d * np.log2(d)))
    return MDL / Ktot


def nonbacktracking(links, degrees, Ntot, Bmax):
    """mod.jl:410-418: leading eigenvalues of the 2N x 2N reduced non-backtracking matrix."""
    from scipy.sparse import bmat, coo_matrix, diags, identity
    from scipy.sparse.linalg import eigs
    K = links.shape[0]
    A = coo_matrix((np.ones(K), (links[:, 0] - 1, links[:, 1] - 1)), shape=(Ntot, Ntot)).tocsr()
    NBT = bmat([[None, diags(degrees - 1.0)], [-identity(Ntot), A]], format="csr")
    vals, _ = eigs(NBT, k=Bmax, which="LR")
    return vals[np.argsort(-vals.real)]


def alluvial_diagram(B, block, maxPSIval):
    """mod.jl:454-504 (significance 0.9, file partition<B>.smap in the working directory)."""
    significant = 0.9
    Ntot = len(maxPSIval)
    order = np.argsort(block, kind="stable")
    labels = block[order]
    trueB, labelprev, heads = 0, 0, []
    mod, sig = np.zeros(Ntot, dtype=np.int64), np.zeros(Ntot, dtype=np.int64)
    for k in range(Ntot):
        if labels[k] != labelprev:
            trueB += 1
            labelprev = labels[k]
            heads.append(f"{order[k] + 1},...")
        mod[k] = trueB
        sig[k] = 1 if maxPSIval[order[k]] > significant else 0
    f16 = gio.julia_float16
    with open(f"partition{B}.smap", "w") as fp:
        fp.write("*Undirected\n")
        fp.write(f"*Modules {trueB}\n")
        for md in range(1, trueB + 1):
            fp.write(f'{md} "{heads[md - 1]}" {f16(0.1 / trueB)} {f16(0.01 / trueB)}\n')
        fp.write("*Insignificants 0\n")
        fp.write(f"*Nodes {Ntot}\n")
        sub, labelprev = 0, 0
        for k in range(Ntot):
            sub = sub + 1 if mod[k] == labelprev else 1
            labelprev = mod[k]
            fp.write(f'{mod[k]}{":" if sig[k] else ";"}{sub} "Node {order[k] + 1}" {f16(0.1 / Ntot)}\n')
        fp.write(f"*Links {trueB ** 2}\n")
        for s in range(1, trueB + 1):
            for t in range(1, trueB + 1):
                fp.write(f"{s} {t} {f16(0.1 / trueB ** 2)}\n")


def run(a, log=print):
    t = lambda s: s == "true"
    strdataset = a.filename
    Barray = gio.parse_q_list(a.q)
    dc, learning, priorlearning, alluvial = t(a.dc), t(a.learning), t(a.prior), t(a.alluvial)
    samples, itrmax = a.initnum, a.itrmax
    learningrate = float(np.float32(a.learningrate))  # parse(Float32, ...), mod.jl:663
    jf = gio.julia_float
    tf = lambda b: "true" if b else "false"

    fpmeta = open("summary.txt", "w")
    fpmeta.write(f"dataset: {strdataset}\n\n")
    links, Ntotinput = gio.read_edgelist(strdataset)
    links = gio.simplegraph(links)
    fpmeta.write(f"number of vertices (input): {Ntotinput}\n")
    fpmeta.write(f"number of edges (input, converted to simple graph): {int(0.5 * links.shape[0])}\n\n")
    Nthreshold = round(Ntotinput / 2)
    Bsize, Bmax = len(Barray), max(Barray)
    cc = gio.connected_component_of(links, Ntotinput, 1)  # mod.jl:735: vertex 1 is assumed to be in the giant component
    log("connected component identified...")
    Ntot, links = gio.links_connected(links, Ntotinput, cc)
    log(f"vertices & edges updated... : N = {Ntot} 2L = {links.shape[0]}")
    if Ntot < Nthreshold:
        log("This is not a giant component... try again.")
    Ltot = int(0.5 * links.shape[0])
    degrees = gio.degree_sequence(links, Ntot)
    blockprev = np.zeros(Ntot, dtype=np.int64)
    assignment = np.zeros((Ntot, Bsize + 1), dtype=np.int64)
    assignment[:, 0] = np.arange(1, Ntot + 1)
    BwithretQ = BwithMDL = BwithCVBP = BwithCVGP = Bmax
    retQopt, MDLopt, CVBPopt, CVGPopt = -1000000.0, 1000000.0, 1000000.0, 1000000.0
    fpmeta.write(f"numbmer of vertices (giant component): {Ntot}\n")
    fpmeta.write(f"numbmer of edges (giant component): {Ltot}\n\n")
    fpmeta.write(f"Degree correction = {tf(dc)}\n")
    fpmeta.write(f"learning rate = {np.float32(learningrate)}\n")
    fpmeta.write(f"prior learning = {tf(priorlearning)}\n")
    fpmeta.write(f"model parameter learning = {tf(learning)}\n")
    fpmeta.write(f"Max number of iteration = {itrmax}\n")
    fpmeta.write(f"Number of initial states = {samples}\n\n")
    nullerror = gio.uniform_graph_error(links, Ntot, dc)
    fpmeta.write(f"Error for q=1 ({'degree-corrected' if dc else 'degree-uncorrected'}): {jf(nullerror)}\n\n")

    retQvec, MDLvec, FEvec = np.zeros(Bsize), np.zeros(Bsize), np.zeros(Bsize)
    CV = {k: np.zeros((Bsize, 2)) for k in ("Bayes", "GP", "GT", "MAP")}
    Alphavec, Betavec, Omegavec = np.zeros(Bsize), np.zeros(Bsize), np.zeros((Bsize, 2))
    betafrac, excm = 0.0, 0.0
    seed = a.seed if a.seed is not None else int(np.random.default_rng().integers(0, 2 ** 31))
    for bb, B in enumerate(Barray):
        log(f"B = {B}")
        FEs = np.zeros(samples)
        CVs = {k: np.zeros(samples) for k in CV}
        varCVs = {k: np.zeros(samples) for k in CV}
        FEmin, itrnumopt = 100000.0, 0
        PSIopt, omegaopt, alphabetaopt = np.zeros((Ntot, B)), np.zeros(2), np.zeros(2)
        gr = np.ones(B) / B
        maxomega = 1
        sm = overflow = 0
        while sm < samples:
            seed += 1
            (excm, alpha, beta, omegain, omegaout, PSI, FE, CVBayes, CVGP, CVGT, CVMAP, vB, vGP, vGT, vMAP, cnv,
             itrnum) = EM_mod(gr, Ntot, B, links, maxomega, itrmax, betafrac, learning, priorlearning, dc, learningrate,
                              seed=seed, damping=a.damping, device=a.device)
            if not cnv:  # mod.jl:819-826
                betafrac += 0.1
                if betafrac <= 1:
                    continue
                log("BP does not converge.... : Raise itrnum or/and BPconv.")
            if (np.isnan(np.max(PSI)) or np.max(PSI) == np.inf or abs(FE) == np.inf or abs(CVGP) == np.inf
                    or abs(CVMAP) == np.inf):  # mod.jl:827-836
                overflow += 1
                if overflow > 10:
                    log("overflow occurs too often...")
                    break
                log("overflow...")
                continue
            sm += 1
            FEs[sm - 1] = FE
            for k, v, vv in (("Bayes", CVBayes, vB), ("GP", CVGP, vGP), ("GT", CVGT, vGT), ("MAP", CVMAP, vMAP)):
                CVs[k][sm - 1] = v
                varCVs[k][sm - 1] = vv
            if FE < FEmin:
                FEmin, itrnumopt, PSIopt = FE, itrnum, PSI
                omegaopt, alphabetaopt = np.array([omegain, omegaout]), np.array([alpha, beta])
        FEvec[bb] = FEs.min()
        for k in CV:
            CV[k][bb, 0] = CVs[k].min()
            CV[k][bb, 1] = math.sqrt(varCVs[k][int(np.argmin(CVs[k]))] / Ltot)
        Alphavec[bb], Betavec[bb] = 2 * Ltot * alphabetaopt[0], alphabetaopt[1]
        Omegavec[bb] = omegaopt
        if CVs["Bayes"].min() < CVBPopt:
            CVBPopt, BwithCVBP = CVs["Bayes"].min(), B
        if CVs["GP"].min() < CVGPopt:
            CVGPopt, BwithCVGP = CVs["GP"].min(), B
        maxPSIval = PSIopt.max(axis=1)
        block = np.argmax(PSIopt, axis=1) + 1
        assignment[:, bb + 1] = block
        fpmeta.write(f"q = {B}: actual q = {len(np.unique(block))}, number of iteration = {itrnumopt} \n")
        retQvec[bb] = retrieval_q(links, degrees, block)
        if retQvec[bb] > retQopt:
            retQopt, BwithretQ = retQvec[bb], B
        MDLvec[bb] = minimum_description_length(links, degrees, block)
        if MDLvec[bb] < MDLopt:
            MDLopt, BwithMDL = MDLvec[bb], B
        block = gio.sortblock(block, blockprev)
        blockprev = block.copy()
        if alluvial:
            alluvial_diagram(B, block, maxPSIval)

    spectralradius, BwithNBT = "-", 0
    if t(a.spectral):  # mod.jl:910-920
        try:
            lam = nonbacktracking(links, degrees, Ntot, Bmax)
            radius = math.sqrt(lam[0].real)
            spectralradius = jf(radius)
            BwithNBT = int(np.sum(lam.real > radius))
        except Exception as e:  # ARPACK may fail on tiny graphs
            log(f"non-backtracking spectrum not available: {e}")
    if BwithNBT == Bmax:
        BwithNBT = 0
    fpmeta.write("\n")
    fpmeta.write(f"mean excess degree = {np.float32(excm)}\n\n")
    fpmeta.write(f"spectral radius of the non-backtracking matrix = {spectralradius}\n")
    fpmeta.write("non-backtracking matrix: q* = -\n" if BwithNBT == 0 else f"non-backtracking matrix: q* = {BwithNBT}\n")
    fpmeta.write(f"modularity: q* = {BwithretQ} (This may not be parsimonius.)\n")
    fpmeta.write(f"map equation: q* = {BwithMDL} (This may not be parsimonius.)\n")
    fpmeta.write(f"CV(Bayes Prediction): q* = {BwithCVBP} (This may not be parsimonius.)\n")
    fpmeta.write(f"CV(Gibbs Prediction): q* = {BwithCVGP} (This may not be parsimonius.)\n")
    fpmeta.close()
    with open("assignment.txt", "w") as fp:
        for i in range(Ntot):
            fp.write(" ".join(str(x) for x in assignment[i]) + "\n")
    with open("assessment.txt", "w") as fp:
        fp.write("Bethe free energy:\n(q, assessed value)\n")
        for bb, B in enumerate(Barray):
            fp.write(f"{B} {jf(FEvec[bb])}\n")
        for title, k in (("Bayes prediction error", "Bayes"), ("Gibbs prediction error", "GP"),
                         ("Gibbs training error", "GT"), ("Gibbs prediction error (MAP estimate)", "MAP")):
            fp.write(f"\n{title}:\n(q, assessed value, standard error)\n")
            for bb, B in enumerate(Barray):
                fp.write(f"{B} {jf(CV[k][bb, 0])} {jf(CV[k][bb, 1])}\n")
        for title, vec in (("Modularity", retQvec), ("MDL (map equation)", MDLvec)):
            fp.write(f"\n{title}:\n(q, assessed value)\n")
            for bb, B in enumerate(Barray):
                fp.write(f"{B} {jf(vec[bb])}\n")
        for title, vec in (("alpha", Alphavec), ("beta", Betavec)):
            fp.write(f"\n{title}:\n(q, learned value)\n")
            for bb, B in enumerate(Barray):
                fp.write(f"{B} {jf(vec[bb])}\n")
        fp.write("\nomega (affinity matrix elements):\n(q, omega_in, omega_out)\n")
        for bb, B in enumerate(Barray):
            fp.write(f"{B} {jf(Omegavec[bb, 0])} {jf(Omegavec[bb, 1])}\n")
    return dict(q=Barray, FE=FEvec, CV=CV, assignment=assignment, modularity=retQvec, mdl=MDLvec,
                alpha=Alphavec, beta=Betavec, omega=Omegavec, q_star=dict(nbt=BwithNBT, modularity=BwithretQ,
                                                                         mdl=BwithMDL, cvbp=BwithCVBP, cvgp=BwithCVGP))


def main(argv=None):
    return run(parse_args(sys.argv[1:] if argv is None else argv))


if __name__ == "__main__":
    main()
