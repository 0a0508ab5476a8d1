"""CPU ORACLE (test infrastructure, NOT product code) -- restatement of `EM()` of the labeled stochastic block model in
/root/reference/src/labeled_sbm/labeled-sbmBIX_0.2.3.8_Julia0.6.2.ipynb (cell 0; the notebook has no line numbers, so
each function below cites the notebook function it follows).

Edge labels alpha = 1..G; one affinity matrix per label; `degrees[i, alpha]` = number of alpha-labelled edges at i
(or all ones without degree correction); external field h[alpha, :] per label.

PARITY UNPINNED: the reference ships no tests or golden vectors for this model and Julia is not available here.  What
pins this transcription: with one label, no degree correction and learning switched off its BP fixed point and
cross-validation errors equal those of oracle/sbm_oracle.py for Crs = Ntot * affinity (tests/test_oracle_labeled.py),
which is itself cross-checked against a C transcription, exact enumeration on trees and the q = 1 closed form.

This is the oracle for DESIGN.md row (f') (CUDA path: graphbix_b200/csrc/gbx_lsbm.cu).  Only tests/, smoke() and the CPU-baseline leg of bench.py may import it.

Conventions kept from the notebook
  * `links` is K x 3 (i, j, label), 1-based vertices, labels 1..G, BOTH directions of every undirected edge with the
    same label (what `labeledsimplegraph` returns);
  * message k <-> links[k] = (i, j) is PSIcav[sub2ind((N,N), i, j)][1:B], the cavity message from i to j; the label
    the notebook appends as a last column is kept in `lab[k]`;
  * nb[i] = in-neighbours of i in ascending order (column i of the sparse matrix).
The asynchronous schedule (`for i in randperm(Ntot)`) is kept; permutations come from a numpy Generator (or `perms`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .sbm_oracle import _Graph, logsumexp, normalize_logprob  # same `logsumexp` / `normalize_logprob` as sbm.jl


@dataclass
class LabeledResult:
    gr: np.ndarray
    affinity: np.ndarray       # (G, B, B)
    PSI: np.ndarray
    FE: float
    CVBayes: float
    CVGP: float
    CVGT: float
    CVMAP: float
    varCVBayes: float
    varCVGP: float
    varCVGT: float
    varCVMAP: float
    cnv: bool
    itrnum: int
    PSIcav: np.ndarray = None
    h: np.ndarray = None
    BPconv: float = float("nan")
    itrs_done: int = 0


def label_degrees(Ntot, links, G, dc=True):
    """`degrees[j, alpha] += 1` over the columns of sparse(links[:,1], links[:,2], links[:,3]) (notebook EM, "initial
    state"); all ones when dc == false."""
    links = np.asarray(links, dtype=np.int64)
    d = np.zeros((Ntot, G))
    np.add.at(d, (links[:, 1] - 1, links[:, 2] - 1), 1.0)
    return d if dc else np.ones((Ntot, G))


def initial_affinity(Ntot, Ktot, Larray, B, G, dc, REDfrac, assortativity, Omega):
    """Notebook EM, "initial state": one matrix per label, assortative ('a'), disassortative ('d') or flat."""
    cm = 2 * np.asarray(Larray, dtype=np.float64) / Ntot
    Norm = Ntot if not dc else Ktot
    aff = np.zeros((G, B, B))
    last = G if REDfrac == 0 else G - 1
    for k in range(last):
        a = assortativity[k] if k < len(assortativity) else ""
        if a == "a":
            delta = 0.99 * cm[k] / Omega
        elif a == "d":
            delta = -0.99 * cm[k] / (1 - Omega)
        else:
            delta = 0.0
        aff[k] = ((cm[k] - delta * Omega) * np.ones((B, B)) + delta * np.eye(B)) / Norm
    if REDfrac != 0:
        aff[G - 1] = cm[G - 1] * np.ones((B, B)) / Norm
    return aff


def EM(Ntot, Larray, B, G, links, itrmax, BPconvthreshold, learning, priorlearning, dc, learningrate, REDfrac,
       assortativity, Omega, *, rng=None, PSIcav0=None, affinity0=None, gr0=None, perms=None) -> LabeledResult:
    """Notebook `function EM(Ntot,Larray,B,G,links,itrmax,BPconvthreshold,learning,priorlearning,dc,learningrate,
    REDfrac,assortativity,Omega)`.  `PSIcav0` (K x B, rows in `links` order) replaces the `rand(1,B)` draws,
    `affinity0` / `gr0` the initial parameters, `perms` the vertex order of each sweep (0-based)."""
    rng = rng if rng is not None else np.random.default_rng(0)
    links = np.asarray(links, dtype=np.int64)
    g = _Graph(Ntot, links[:, :2])
    Ktot = g.Ktot
    Ltot = int(0.5 * Ktot)
    lab = links[:, 2] - 1
    ptr, in_link, in_src, rev = g.ptr, g.in_link, g.in_src, g.rev
    if np.any(lab[rev] != lab):
        raise ValueError("both directions of an edge must carry the same label")
    degrees = label_degrees(Ntot, links, G, dc)
    aff = (np.array(affinity0, dtype=np.float64).copy() if affinity0 is not None else
           initial_affinity(Ntot, Ktot, Larray, B, G, dc, REDfrac, assortativity, Omega))
    gr = np.ones(B) / B if gr0 is None else np.array(gr0, dtype=np.float64).ravel().copy()
    h = np.zeros((G, B))
    if PSIcav0 is None:
        PSIcav = rng.random((Ktot, B))
        PSIcav = PSIcav / PSIcav.sum(axis=1, keepdims=True)
    else:
        PSIcav = np.array(PSIcav0, dtype=np.float64).reshape(Ktot, B).copy()
    PSI = np.zeros((Ntot, B))
    st = dict(gr=gr, aff=aff, h=h)

    def update_h():                                      # notebook update_h
        hh = np.zeros((G, B))
        for a in range(G):
            hh[a] = (degrees[:, a][:, None] * (PSI @ st["aff"][a])).sum(axis=0)
        return hh

    def logunnormalizedMessage(i):                       # notebook logunnormalizedMessage
        logmessages = np.zeros(B)
        for p in range(ptr[i], ptr[i + 1]):
            s = in_src[p]
            k = in_link[p]                               # message s -> i
            a = lab[k]
            logmessages = logmessages + np.log(degrees[i, a] * degrees[s, a] * (PSIcav[k] @ st["aff"][a]) * Ntot)
        ext = degrees[i] @ st["h"]                       # sum_alpha degrees[i,alpha] h[alpha,:]
        return logmessages + np.log(st["gr"]) - ext

    def BP(perm):                                        # notebook BP
        conv = 0.0
        for i in perm:
            logPSIi = logunnormalizedMessage(i)
            for p in range(ptr[i], ptr[i + 1]):
                kji = in_link[p]
                kij = rev[kji]
                j = in_src[p]
                a = lab[kji]
                cav = logPSIi - np.log(degrees[i, a] * degrees[j, a] * (PSIcav[kji] @ st["aff"][a]) * Ntot)
                PSIcav[kij] = normalize_logprob(cav)
            prev = PSI[i] / Ntot
            logPSIi = logunnormalizedMessage(i)          # "new PSI with new PSIcav"
            PSI[i] = normalize_logprob(logPSIi)
            conv += np.sum(np.abs(PSI[i] / Ntot - prev))
        return conv

    def update_affinity():                               # notebook update_affinity
        new = np.zeros((G, B, B))
        for k in range(Ktot):
            i, j = links[k, 0] - 1, links[k, 1] - 1
            if i > j:
                continue
            a = lab[k]
            PSIij = degrees[i, a] * degrees[j, a] * np.outer(PSIcav[k], PSIcav[rev[k]]) * (Ntot * st["aff"][a])
            new[a] += PSIij / PSIij.sum()
        denom = degrees.T @ PSI                          # denom[alpha,:] = sum_i degrees[i,alpha] PSI[i,:]
        out = np.zeros((G, B, B))
        for a in range(G):
            out[a] = (new[a] + new[a].T) / np.outer(denom[a], denom[a])
            out[a][out[a] < 10 ** (-8.0)] = 10 ** (-8.0)
        if REDfrac != 0:
            out[G - 1] = (1 / Ktot) * np.ones((B, B))
        return out

    def freeenergy():                                    # notebook freeenergy
        sumlogZi = 0.0
        for i in range(Ntot):
            sumlogZi += logsumexp(logunnormalizedMessage(i))
        logZijs, CVGPs, CVGTs, CVMAPs = np.zeros(Ltot), np.zeros(Ltot), np.zeros(Ltot), np.zeros(Ltot)
        cnt = 0
        for k in range(Ktot):
            i, j = links[k, 0] - 1, links[k, 1] - 1
            if i < j:
                continue
            a = lab[k]
            A = st["aff"][a]
            dd = degrees[i, a] * degrees[j, a]
            w = np.outer(PSIcav[k], PSIcav[rev[k]])
            Zij = float(np.sum(dd * A * w))
            logZijs[cnt] = math.log(Zij)
            with np.errstate(divide="ignore"):
                lterm = np.log(dd * A)
            CVGPs[cnt] = np.sum(w * lterm)
            CVGTs[cnt] = np.sum(w * dd * A * lterm) / Zij
            s, t = int(np.argmax(PSIcav[k])), int(np.argmax(PSIcav[rev[k]]))
            CVMAPs[cnt] = lterm[s, t]
            cnt += 1
        FE = -(sumlogZi - logZijs.sum()) / Ntot - Ltot / Ntot
        var = lambda x: float(np.var(x, ddof=1)) if len(x) > 1 else float("nan")
        return (FE, 1 - logZijs.sum() / Ltot, 1 - CVGPs.sum() / Ltot, 1 - CVGTs.sum() / Ltot,
                1 - CVMAPs.sum() / Ltot, var(logZijs), var(CVGPs), var(CVGTs), var(CVMAPs))

    # ---- main body ---------------------------------------------------------------------------------------------
    for i in range(Ntot):                                # updatePSI
        PSI[i] = normalize_logprob(logunnormalizedMessage(i))
    st["h"] = update_h()
    cnv, itrnum, BPconv, itrs_done = False, 0, float("nan"), 0
    for itr in range(1, itrmax + 1):
        st["h"] = update_h()
        perm = perms[itr - 1] if perms is not None else rng.permutation(Ntot)
        BPconv = BP(perm)
        itrs_done = itr
        # Closure write-through of the notebook: `update_gr` rebinds EM()'s `gr` before the right operand of
        # `gr = learningrate*update_gr(PSI) + (1 - learningrate)*gr` is read (Julia evaluates left to right), and
        # `update_affinity` mutates and returns the SAME Dict, so both blends mix the new value with itself:
        # the learning rate has no effect on the notebook's trajectory.
        if priorlearning:
            st["gr"] = PSI.mean(axis=0)                                          # update_gr: gr = mean(PSI,1)
            st["gr"] = learningrate * st["gr"] + (1 - learningrate) * st["gr"]
        if learning:
            st["aff"] = update_affinity()                                        # affinity[alpha] = ... in place
            st["aff"] = learningrate * st["aff"] + (1 - learningrate) * st["aff"]
        if BPconv < BPconvthreshold:
            cnv, itrnum = True, itr
            break
    fe = freeenergy()
    return LabeledResult(st["gr"], st["aff"], PSI, *fe, cnv, itrnum, PSIcav=PSIcav, h=st["h"], BPconv=BPconv,
                         itrs_done=itrs_done)


def labeled_planted_partition(n, B, c_by_label, kinds, eps=0.1, seed=0):
    """Synthetic labeled graph in the spirit of the notebook's `labeledSBMedgelist_Nblock=[300,300,300]_c=[5,3]...`:
    B equal planted groups; label alpha has mean degree c_by_label[alpha] and is assortative ('a': c_out = eps c_in) or
    disassortative ('d': c_in = eps c_out).  Returns (links K x 3 with both directions, planted labels 0..B-1)."""
    rng = np.random.default_rng(seed)
    lab = (np.arange(n) * B // n).astype(np.int64)
    rows = []
    seen = set()
    for a, (c, kind) in enumerate(zip(c_by_label, kinds)):
        if kind == "a":
            cin = c * B / (1 + (B - 1) * eps)
            cout = eps * cin
        else:
            cout = c * B / ((B - 1) + eps)
            cin = eps * cout
        m = rng.poisson(c * n / 2 * 1.3)
        u = rng.integers(0, n, size=m)
        v = rng.integers(0, n, size=m)
        p = np.where(lab[u] == lab[v], cin, cout) / max(cin, cout)
        keep = (rng.random(m) < p) & (u != v)
        for x, y in zip(u[keep], v[keep]):
            e = (min(x, y), max(x, y))
            if e in seen:
                continue
            seen.add(e)
            rows.append((e[0] + 1, e[1] + 1, a + 1))
    und = np.array(rows, dtype=np.int64)
    links = np.vstack([und, und[:, [1, 0, 2]]])
    return links, lab
