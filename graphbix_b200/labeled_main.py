"""Host mirror of the MAIN cell of the reference's labeled-SBM notebook
(src/labeled_sbm/labeled-sbmBIX_0.2.3.8_Julia0.6.2.ipynb, cell 0, "MAIN"): reads a labeled edge list ("i j label" per
line), keeps the connected component of node 1, optionally discards a fraction of one label's edges to the neutral
label (RED), runs EM for every B of `Barray` from `samples` random initial states and writes

    summary_<dataset>.txt   assessment_<dataset>.txt   assignment_<dataset>.txt   [REDedgelist_<dataset>.txt]

in the notebook's formats.  Plots and the alluvial .smap files are not produced.  The EM + BP loop runs in
libgraphbix_cuda (gbx_lsbm_*) through graphbix_b200.labeled.EM_labeled.

    python -m graphbix_b200.labeled_main edgelist.txt [--dataset=NAME] [--q=2:3] [--samples=5] [--dc=true]
           [--redalpha=1] [--redfrac=0.5] [--assortativity=d,a] [--itrmax=512] [--learningrate=0.3] [--seed=0]
"""
from __future__ import annotations

import argparse
import math

import numpy as np

from .graphio import connected_component_of, julia_float, links_connected, parse_q_list
from .labeled import EM_labeled, labeledsimplegraph


def read_labeled_edgelist(path):
    """Notebook MAIN: node ids are the distinct ids in order of first appearance, renumbered 1..Ntotinput."""
    rows = []
    with open(path) as fp:
        for line in fp:
            u = line.rstrip("\n").split(" ")
            if len(u) < 3 or not u[0]:
                continue
            rows.append((int(u[0]), int(u[1]), float(u[2])))
    ids = []
    seen = {}
    for a, b, _ in rows:
        for x in (a, b):
            if x not in seen:
                seen[x] = len(ids) + 1
                ids.append(x)
    links = np.array([(seen[a], seen[b], int(c)) for a, b, c in rows], dtype=np.int64).reshape(-1, 3)
    return links, np.array(ids, dtype=np.int64)


def edges_per_label(links):
    """Notebook MAIN: Larray[alpha] = number of undirected links of each label (labels sorted ascending)."""
    und = links[links[:, 0] < links[:, 1]]
    G = int(links[:, 2].max())
    return [int((und[:, 2] == a + 1).sum()) for a in range(G)]


def random_edge_discarding(links, G, Larray, REDalpha, REDfrac, rng):
    """Notebook `RandomEdgeDiscarding`: a fraction REDfrac of the links of label REDalpha gets the new label G + 1."""
    und = links[links[:, 0] > links[:, 1]]
    red = und[und[:, 2] == REDalpha].copy()
    rest = und[und[:, 2] != REDalpha]
    k = math.ceil(REDfrac * len(red))
    red[rng.permutation(len(red))[:k], 2] = G + 1
    und = np.vstack([red, rest])
    both = np.vstack([und, und[:, [1, 0, 2]]])
    Larray = list(Larray)
    Larray[REDalpha - 1] -= k
    Larray.append(k)
    return both, Larray


def uniform_graph_error(links, Ntot, Larray, dc, G, Ltot, REDfrac):
    """Notebook `UniformGraphError`: prediction error of the one-group model."""
    Ls = np.array(Larray[:G - 1] if REDfrac != 0 else Larray, dtype=np.float64)
    Pas = Ls / Ls.sum()
    entropy = float(-(Pas[Pas > 0] * np.log(Pas[Pas > 0])).sum())
    if dc:
        d = np.zeros((Ntot, G))
        np.add.at(d, (links[:, 1] - 1, links[:, 2] - 1), 1.0)
        Gmax = G - 1 if REDfrac != 0 else G
        dd = d[:, :Gmax]
        dd = dd[dd > 0]
        return 1 - float((dd * np.log(dd)).sum()) / Ltot - entropy + math.log(2 * Ltot)
    return 1 + entropy - math.log(2 * Ltot / (Ntot ** 2))


def relabel_psi(PSI, gr, affinity):
    """Notebook `RelabelPSI`: groups in descending order of their number of MAP members."""
    B = len(gr)
    hist = np.bincount(np.argmax(PSI, axis=1), minlength=B)
    ind = np.argsort(-hist, kind="stable")
    return PSI[:, ind], np.asarray(gr)[ind], affinity[:, ind][:, :, ind]


def _julia_vec(v):
    return "[" + " ".join(julia_float(float(x)) for x in np.ravel(v)) + "]"


def run(path, dataset="labeledSBM", Barray=(2, 3), samples=5, dc=True, priorlearning=True, learning=True, itrmax=512,
        learningrate=0.3, BPconvthreshold=0.000001, REDalpha=1, REDfrac=0.5, assortativity=("d", "a"), seed=0, device=0,
        outdir=".", max_retries=20, log=print):
    import os
    rng = np.random.default_rng(seed)
    links, nodeids = read_labeled_edgelist(path)
    Ntotinput = len(nodeids)
    links = labeledsimplegraph(links)
    fpmeta = open(os.path.join(outdir, f"summary_{dataset}.txt"), "w")
    fpmeta.write(f"dataset: {os.path.basename(path)}\n\n")
    fpmeta.write(f"number of vertices (input): {Ntotinput}\n")
    fpmeta.write(f"number of edges (input, converted to simple graph): {links.shape[0] // 2}\n\n")
    cc = connected_component_of(links[:, :2], Ntotinput, 1)          # "Assume node 1 belongs to the connected component"
    log("connected component identified...")
    Ntot, l2 = links_connected(links[:, :2], Ntotinput, cc)
    keep = np.isin(links[:, 0], cc) | np.isin(links[:, 1], cc)
    links = np.column_stack([l2, links[keep, 2]])
    log(f"vertices & edges updated... : N = {Ntot} 2L = {links.shape[0]}")
    if Ntot != Ntotinput:
        log("Input graph is NOT a connected graph!")
    Larray = edges_per_label(links)
    Ltot = int(sum(Larray))
    G = len(Larray)
    if REDfrac != 0:
        links, Larray = random_edge_discarding(links, G, Larray, REDalpha, REDfrac, rng)
        G += 1
        with open(os.path.join(outdir, f"REDedgelist_{dataset}.txt"), "w") as f:
            for i, j, a in links:
                if i <= j:
                    f.write(f"{nodeids[i - 1]} {nodeids[j - 1]} {a}\n")
    Barray = list(Barray)
    Bsize = len(Barray)
    assignment = np.zeros((Ntot, Bsize + 1), dtype=np.int64)
    assignment[:, 0] = nodeids[np.sort(cc) - 1] if Ntot != Ntotinput else nodeids
    fpmeta.write(f"Total numbmer of vertices (giant component): {Ntot}\n")
    fpmeta.write(f"Total numbmer of edges (giant component): {Ltot}\n")
    fpmeta.write(f"Numbmer of edges of each type (giant component): {Larray}\n\n")
    fpmeta.write(f"Max number of iteration = {itrmax}\n")
    fpmeta.write(f"BP convergence threshold = {julia_float(BPconvthreshold)}\n")
    fpmeta.write(f"degree correction: {str(bool(dc)).lower()}\n")
    fpmeta.write(f"learning rate: {julia_float(learningrate)}\n")
    fpmeta.write(f"Label for Random Edge Discarding: {REDalpha}\n")
    fpmeta.write(f"Fraction of edges that are discarded (neutralized): {julia_float(REDfrac)}\n")
    fpmeta.write(f"Number of initial states = {samples}\n\n")
    nullerror = uniform_graph_error(links, Ntot, Larray, dc, G, Ltot, REDfrac)
    fpmeta.write(f"Prediction error for q = 1 = {julia_float(nullerror)}\n\n")

    FEvec = np.zeros(Bsize)
    CV = {k: np.zeros((Bsize, 2)) for k in ("Bayes", "GP", "GT", "MAP")}
    grD, affD = {}, {}
    BwithCVBP = BwithCVGP = max(Barray)
    CVBPopt = CVGPopt = 1000000.0
    for bb, B in enumerate(Barray):
        log(f"B = {B}")
        vals = {k: np.zeros(samples) for k in ("FE", "Bayes", "GP", "GT", "MAP", "vBayes", "vGP", "vGT", "vMAP")}
        FEmin, itrnumopt = 100000.0, 0
        PSIopt = gropt = affopt = None
        sm = tries = 0
        while sm < samples:
            out = EM_labeled(Ntot, Larray, B, G, links, itrmax, BPconvthreshold, learning, priorlearning, dc, learningrate,
                             REDfrac, list(assortativity), 1 / B, rng=rng, device=device)
            (gr, aff, PSI, FE, cvb, cvgp, cvgt, cvmap, vb, vgp, vgt, vmap, cnv, itrnum) = out
            if not cnv:
                tries += 1
                log("try again...")
                if tries > max_retries:      # the notebook retries for ever; a mirror used in tests must stop
                    raise RuntimeError(f"B = {B}: BP did not converge in {max_retries} attempts")
                continue
            for k, v in zip(vals, (FE, cvb, cvgp, cvgt, cvmap, vb, vgp, vgt, vmap)):
                vals[k][sm] = v
            sm += 1
            if FE < FEmin:
                FEmin, itrnumopt, PSIopt, gropt, affopt = FE, itrnum, PSI, gr, aff
        PSIopt, gropt, affopt = relabel_psi(PSIopt, gropt, affopt)
        FEvec[bb] = vals["FE"].min()
        for k in ("Bayes", "GP", "GT", "MAP"):
            j = int(np.argmin(vals[k]))
            CV[k][bb] = (vals[k][j], math.sqrt(vals["v" + k][j] / Ltot))
        grD[bb], affD[bb] = gropt, affopt
        if CV["Bayes"][bb, 0] < CVBPopt:
            CVBPopt, BwithCVBP = CV["Bayes"][bb, 0], B
        if CV["GP"][bb, 0] < CVGPopt:
            CVGPopt, BwithCVGP = CV["GP"][bb, 0], B
        block = np.argmax(PSIopt, axis=1) + 1
        assignment[:, bb + 1] = block
        fpmeta.write(f"q = {B}: actual q = {len(np.unique(block))}, number of iteration = {itrnumopt} \n")
    fpmeta.write("\n")
    fpmeta.write(f"CV(Bayes Prediction): q* = {BwithCVBP} (This may not be parsimonius.)\n")
    fpmeta.write(f"CV(Gibbs Prediction): q* = {BwithCVGP} (This may not be parsimonius.)\n")
    fpmeta.close()
    with open(os.path.join(outdir, f"assignment_{dataset}.txt"), "w") as f:
        for row in assignment:
            f.write(" ".join(str(int(x)) for x in row) + "\n")
    with open(os.path.join(outdir, f"assessment_{dataset}.txt"), "w") as f:
        f.write("Bethe free energy:\n(q, assessed value)\n")
        for bb, B in enumerate(Barray):
            f.write(f"{B} {julia_float(FEvec[bb])}\n")
        for title, k in (("Bayes prediction error", "Bayes"), ("Gibbs prediction error", "GP"),
                         ("Gibbs training error", "GT"), ("Gibbs prediction error (MAP estimate)", "MAP")):
            f.write(f"\n{title}:\n(q, assessed value, standard error)\n")
            for bb, B in enumerate(Barray):
                f.write(f"{B} {julia_float(CV[k][bb, 0])} {julia_float(CV[k][bb, 1])}\n")
        f.write("\nModule sizes:\n")
        for bb, B in enumerate(Barray):
            f.write(f"q = {B}\n{_julia_vec(grD[bb])}\n")
        f.write("\nAffinity matrix of each alpha \n")
        for bb, B in enumerate(Barray):
            f.write(f"q = {B}\n")
            for a in range(G):
                f.write(f"{a + 1} => " + "; ".join(" ".join(julia_float(float(x)) for x in r) for r in affD[bb][a]) + "\n")
    return dict(Barray=Barray, FE=FEvec, CV=CV, q_star_bayes=BwithCVBP, q_star_gibbs=BwithCVGP, assignment=assignment,
                gr=grD, affinity=affD, Larray=Larray, Ntot=Ntot)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("edgelist")
    ap.add_argument("--dataset", default="labeledSBM")
    ap.add_argument("--q", default="2:3")
    ap.add_argument("--samples", type=int, default=5)
    ap.add_argument("--dc", default="true")
    ap.add_argument("--redalpha", type=int, default=1)
    ap.add_argument("--redfrac", type=float, default=0.5)
    ap.add_argument("--assortativity", default="d,a")
    ap.add_argument("--itrmax", type=int, default=512)
    ap.add_argument("--learningrate", type=float, default=0.3)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    run(a.edgelist, a.dataset, parse_q_list(a.q), a.samples, a.dc.lower() == "true", itrmax=a.itrmax,
        learningrate=a.learningrate, REDalpha=a.redalpha, REDfrac=a.redfrac, assortativity=a.assortativity.split(","),
        seed=a.seed, device=a.device)


if __name__ == "__main__":
    main()
