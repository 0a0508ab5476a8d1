"""CPU ORACLE (test infrastructure, NOT product code) -- restatement of `EM()` in
/root/reference/src/mod.jl:3-244 (graphBIX, SBM restricted to community structure: omega_in / omega_out,
alpha, beta; EM + belief propagation + leave-one-out cross-validation errors).

PARITY UNPINNED: as for oracle/sbm_oracle.py -- no Julia toolchain, no golden vectors upstream; this is a
statement-by-statement transcription (each function cites the mod.jl lines it follows) INCLUDING the reference's
closure semantics (nested functions that assign a variable of the enclosing EM() rebind it: `h` in update_h,
`gr` in update_gr, `omegain`/`omegaout` in update_omega, `alpha`/`beta` in update_alphabeta), checked against the
independent synchronous formulation (tests/sync_model_mod.py) and, because belief propagation is exact on trees,
against brute-force enumeration of the model's labelings on small trees (tests/test_oracle_exact_tree.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .sbm_oracle import _Graph, logsumexp, normalize_logprob  # mod.jl:5-20 are the same two helpers


@dataclass
class ModResult:
    """Tuple returned by EM(), mod.jl:243, in the same order (plus the final messages)."""
    excm: float
    alpha: float
    beta: float
    omegain: float
    omegaout: float
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
    PSIcav: np.ndarray = field(repr=False, default=None)
    gr: np.ndarray = field(repr=False, default=None)
    BPconv: float = float("nan")
    itrs_done: int = 0


def initial_omegas(degrees_true: np.ndarray, Ntot: int, Ktot: int, B: int, betafrac: float):
    """mod.jl:184-196."""
    d = degrees_true
    excm = float(d @ d) / (Ntot * d.mean()) - 1          # average excess degree
    beta0 = math.log(1 + B / (excm - 1))
    betaast = math.log(1 + B / (math.sqrt(excm) - 1))
    alpha = 1 / Ktot
    beta = betaast - betafrac * (betaast - beta0)
    omegain = math.exp(beta) * alpha * beta / (math.exp(beta) - 1)
    omegaout = alpha * beta / (math.exp(beta) - 1)
    return excm, alpha, beta, omegain, omegaout


def EM(gr, Ntot, B, links, maxomega, itrmax, betafrac, learning, priorlearning, dc, learningrate, *,
       rng=None, PSIcav0=None, BPconvthreshold=0.000001, perms=None) -> ModResult:
    """mod.jl:3-244.  `PSIcav0` (K x B, rows in `links` order) replaces the `rand(1,B)` draws of mod.jl:198-203;
    `perms` optionally fixes the vertex order of each sweep."""
    rng = rng if rng is not None else np.random.default_rng(0)
    g = _Graph(Ntot, links)
    Ktot = g.Ktot
    links = g.links
    ptr, in_link, in_src, rev = g.ptr, g.in_link, g.in_src, g.rev
    nbdeg = np.diff(ptr).astype(np.float64)               # length(nb[i])

    cnv = False
    itrnum = 0
    degrees = g.degrees.copy()                            # sum(A,2), mod.jl:184
    excm, alpha, beta, omegain, omegaout = initial_omegas(degrees, Ntot, Ktot, B, betafrac)
    if not dc:
        degrees = np.ones(Ntot)                           # mod.jl:189-191
    if PSIcav0 is None:                                   # mod.jl:198-203
        PSIcav = rng.random((Ktot, B))
        PSIcav = PSIcav / PSIcav.sum(axis=1, keepdims=True)
    else:
        PSIcav = np.array(PSIcav0, dtype=np.float64).reshape(Ktot, B).copy()
    st = dict(gr=np.array(gr, dtype=np.float64).ravel().copy(), h=np.zeros(B), alpha=alpha, beta=beta,
              omegain=omegain, omegaout=omegaout)
    PSI = np.zeros((Ntot, B))

    def update_h():                                       # mod.jl:22-24
        return degrees @ PSI

    def logunnormalizedMessage(i):                        # mod.jl:38-52
        eb1 = math.exp(st["beta"]) - 1
        logmessages = np.zeros(B)
        for p in range(ptr[i], ptr[i + 1]):
            logmessages = logmessages + np.log(np.ones(B) + PSIcav[in_link[p]] * eb1)
        if not dc:
            return logmessages - st["alpha"] * st["beta"] * st["h"] + np.log(st["gr"])
        return logmessages - st["alpha"] * st["beta"] * nbdeg[i] * st["h"] + np.log(st["gr"])

    def BP(perm):                                         # mod.jl:54-75
        conv = 0.0
        st["h"] = update_h()
        for i in perm:
            eb1 = math.exp(st["beta"]) - 1
            logPSIi = logunnormalizedMessage(i)
            for p in range(ptr[i], ptr[i + 1]):
                kji = in_link[p]
                kij = rev[kji]
                PSIcav[kij] = normalize_logprob(logPSIi - np.log(np.ones(B) + PSIcav[kji] * eb1))
            hprev = degrees[i] * PSI[i]
            prev = PSI[i] / Ntot
            logPSIi = logunnormalizedMessage(i)
            PSI[i] = normalize_logprob(logPSIi)
            st["h"] = st["h"] + (degrees[i] * PSI[i] - hprev)
            conv += np.sum(np.abs(PSI[i] / Ntot - prev))
        return conv

    def update_omega():                                   # mod.jl:77-106
        oin, oout = st["omegain"], st["omegaout"]
        omegainnew = 0.0
        for k in range(Ktot):
            i, j = links[k, 0], links[k, 1]
            if i > j:
                continue
            s = float(PSIcav[k] @ PSIcav[rev[k]])
            omegainnew += oin * s / ((oin - oout) * s + oout)
        dsigmanorm = float(np.linalg.norm(update_h()) ** 2)
        oin = 2 * omegainnew / dsigmanorm
        if not dc:
            oout = 2 * (0.5 * Ktot - omegainnew) / (Ntot ** 2 - dsigmanorm)
        else:
            oout = 2 * (0.5 * Ktot - omegainnew) / (Ktot ** 2 - dsigmanorm)
        if oin > maxomega:
            oin = maxomega
        elif oin < 0:
            oin = 0
        elif oout < 0:
            oout = 0
        return oin, oout

    def freeenergy():                                     # mod.jl:114-168
        al, be, oin, oout = st["alpha"], st["beta"], st["omegain"], st["omegaout"]
        sumlogZi = 0.0
        for i in range(Ntot):
            sumlogZi += logsumexp(logunnormalizedMessage(i))
        Ltot = int(0.5 * Ktot)
        logZijs, CVBPs, CVGPs, CVGTs, CVMAPs = (np.zeros(Ltot) for _ in range(5))
        cnt = 0
        for k in range(Ktot):
            i, j = links[k, 0] - 1, links[k, 1] - 1
            if i < j:
                continue
            a, b = PSIcav[k], PSIcav[rev[k]]
            s = float(a @ b)
            logZijs[cnt] = math.log(1 + (math.exp(be) - 1) * s)
            CVBPs[cnt] = logZijs[cnt] + math.log(degrees[i]) + math.log(oout) + math.log(degrees[j])
            CVGPs[cnt] = be * s
            CVGTs[cnt] = degrees[i] * degrees[j] * (oout * math.log(oout) + (oin * math.log(oin) - oout * math.log(oout)) * s) / math.exp(CVBPs[cnt])
            CVMAPs[cnt] = math.log(oin) if int(np.argmax(a)) == int(np.argmax(b)) else math.log(oout)
            cnt += 1
        constshift = float(np.sum(degrees * np.log(degrees))) / Ltot
        hn2 = float(np.linalg.norm(update_h()) ** 2)
        FE = (-(sumlogZi - logZijs.sum() + 0.5 * al * be * hn2) / (be * Ntot)
              - (Ltot / Ntot) * (2 * Ltot * oout + math.log(oout) + constshift) / be)
        var = lambda x: float(np.var(x, ddof=1))
        return (FE, -CVBPs.sum() / Ltot + 1, -CVGPs.sum() / Ltot - math.log(oout) + 1 - constshift,
                -CVGTs.sum() / Ltot + 1 - constshift, -CVMAPs.sum() / Ltot + 1 - constshift,
                var(logZijs), var(CVGPs), var(CVGTs), var(CVMAPs))

    # ---- main body, mod.jl:204-241 ------------------------------------------------------
    for i in range(Ntot):                                 # PSI = updatePSI() with h = 0, mod.jl:206
        PSI[i] = normalize_logprob(logunnormalizedMessage(i))
    BPconv = float("nan")
    itrs_done = 0
    for itr in range(1, itrmax + 1):
        perm = perms[itr - 1] if perms is not None else rng.permutation(Ntot)
        BPconv = BP(perm)
        if priorlearning:
            st["gr"] = PSI.mean(axis=0)                   # mod.jl:221-223
        if learning:                                      # mod.jl:224-229
            # Closure write-through of the reference: update_omega() assigns the ENCLOSING EM() locals
            # `omegain` / `omegaout` (mod.jl:90-94 -- it reads them at mod.jl:87, so they are not new locals), hence
            # when the caller blends at mod.jl:226-227 the "old" value already is the new one:
            # learningrate*new + (1-learningrate)*new.  --learningrate has no effect on mod.jl's trajectory.
            oin_new, oout_new = update_omega()
            st["omegain"], st["omegaout"] = oin_new, oout_new                  # the write-through, mod.jl:90-103
            st["omegain"] = learningrate * oin_new + (1 - learningrate) * st["omegain"]       # mod.jl:226
            st["omegaout"] = learningrate * oout_new + (1 - learningrate) * st["omegaout"]    # mod.jl:227
            st["beta"] = math.log(st["omegain"] / st["omegaout"])          # update_alphabeta, mod.jl:108-112
            st["alpha"] = (st["omegain"] - st["omegaout"]) / st["beta"]
        itrs_done = itr
        if BPconv < BPconvthreshold:
            cnv = True
            itrnum = itr
            break
    st["h"] = update_h()                                  # mod.jl:240
    out = freeenergy()
    return ModResult(excm, st["alpha"], st["beta"], st["omegain"], st["omegaout"], PSI, *out, cnv, itrnum,
                     PSIcav=PSIcav, gr=st["gr"], BPconv=BPconv, itrs_done=itrs_done)
