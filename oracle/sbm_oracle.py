"""CPU ORACLE (test infrastructure, NOT product code) -- restatement of `EM()` in
/root/reference/src/sbm.jl:28-370 (graphBIX, full-affinity SBM: EM + belief
propagation + leave-one-out cross-validation errors).

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, and
there is no Julia toolchain in this image, so this restatement could not be
checked against outputs of the reference itself.  It is a statement-by-statement
transcription (each function cites the sbm.jl lines it follows), cross-checked by
(a) an independent C transcription (oracle/sbm_oracle.c), (b) exact enumeration of
the pairwise model on small trees (tests/test_oracle_exact_tree.py) and (c) the
reference's own closed form for q=1 (sbm.jl:477-492, `UniformGraphError`).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (graphbix_b200/) never does.

Conventions kept from the reference
  * vertices are 1-based in `links` (K x 2 int array holding BOTH directions of every
    undirected edge, as produced by `simplegraph`, sbm.jl:414-425);
  * message k  <->  links[k] = (i, j)  is  PSIcav[sub2ind((N,N), i, j)], the cavity
    message from i to j (sbm.jl:106-109);
  * nb[i] = in-neighbours of i in ascending order (column i of the sparse adjacency,
    sbm.jl:267-273).
The asynchronous schedule (`for i in randperm(Ntot)`, sbm.jl:103) is kept; the random
permutations come from a numpy Generator instead of Julia's MersenneTwister, so single
trajectories differ from a Julia run while fixed points do not.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

LOG_1EM8 = math.log(10 ** (-8.0))  # sbm.jl:42  log.(10^(-8.0))


# --------------------------------------------------------------------------------------
# sbm.jl:30-37
def logsumexp(array) -> float:
    a = np.sort(np.asarray(array).ravel())  # vec(sortslices(array, dims=2)) on a 1 x n matrix
    maxval = a.max()
    if a.dtype == np.float32:  # freeenergy() feeds a Float32 vector (sbm.jl:170-176)
        a = a.copy()
        lim = np.float32(maxval - np.float32(500))
        a[(maxval - a) > np.float32(500)] = lim
        a = a.astype(np.float64)
    else:
        a = np.where(maxval - a > 500, maxval - 500, a)
    return float(a[0] + np.log(np.sum(np.exp(a - a[0]))))


# sbm.jl:39-45
def normalize_logprob(array) -> np.ndarray:
    a = np.array(array, dtype=np.float64).ravel()
    a[a < LOG_1EM8] = LOG_1EM8
    return np.exp(a - logsumexp(a))


# --------------------------------------------------------------------------------------
@dataclass
class EMResult:
    """Tuple returned by EM(), sbm.jl:369, in the same order (plus the final messages)."""
    gr: np.ndarray
    Crs: np.ndarray
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
    fail: bool
    cnv: bool
    itrnum: int
    PSIcav: np.ndarray = field(repr=False, default=None)   # K x B, row k <-> links[k]
    theta: np.ndarray = field(repr=False, default=None)
    h: np.ndarray = field(repr=False, default=None)
    BPconv: float = float("nan")
    itrs_done: int = 0

    def astuple(self):
        return (self.gr, self.Crs, self.PSI, self.FE, self.CVBayes, self.CVGP, self.CVGT, self.CVMAP,
                self.varCVBayes, self.varCVGP, self.varCVGT, self.varCVMAP, self.fail, self.cnv, self.itrnum)


class _Graph:
    """Index structures of sbm.jl:265-275."""

    def __init__(self, Ntot: int, links: np.ndarray):
        links = np.asarray(links, dtype=np.int64)
        self.Ntot = int(Ntot)
        self.links = links
        self.Ktot = links.shape[0]
        src = links[:, 0] - 1
        dst = links[:, 1] - 1
        key = src * self.Ntot + dst
        order = np.argsort(key, kind="stable")
        skey = key[order]
        if self.Ktot and np.any(skey[1:] == skey[:-1]):
            raise ValueError("links must not contain duplicates (run simplegraph first)")
        rkey = dst * self.Ntot + src
        pos = np.searchsorted(skey, rkey)
        if self.Ktot and (np.any(pos >= self.Ktot) or np.any(skey[np.minimum(pos, self.Ktot - 1)] != rkey)):
            raise ValueError("links must contain both directions of every edge")
        self.rev = order[pos]                       # rev[k] = index of the link (j, i)
        # in_links[i]: link ids of messages s -> i, s ascending (nb[i], sbm.jl:269-273)
        ordd = np.lexsort((src, dst))
        counts = np.bincount(dst, minlength=self.Ntot)
        self.ptr = np.concatenate(([0], np.cumsum(counts)))
        self.in_link = ordd                          # link ids grouped by destination
        self.in_src = src[ordd]
        self.degrees = counts.astype(np.float64)     # sum(A, dims=2); A is symmetric


def initial_parameters(initial: str, B: int, Ntot: int, rng: np.random.Generator):
    """sbm.jl:279-302 (all but normalizedLaplacian, whose spectral part stays on the host)."""
    if initial == "random":
        RandMat = 0.001 * rng.random((B, B)) + 0.0001
        polarize = rng.integers(0, B)
        polarize2 = rng.integers(0, B)
        RandMat[polarize, polarize2] = 10
        Crs = RandMat + RandMat.T
        gr = np.ones(B) / B
    elif initial == "uniformAssortative":
        Crs = 20 * np.eye(B) + 0.01 * np.ones((B, B))
        gr = np.ones(B) / B
    elif initial == "uniformDisassortative":
        Crs = -9.9 * np.eye(B) + 10 * np.ones((B, B))
        gr = np.ones(B) / B
    elif initial == "Bipartite":
        halfB = int(round(B / 2))
        layer1 = np.array([[0.0, 1.0], [1.0, 0.0]])
        layer2 = 20 * np.eye(halfB) + 0.01 * np.ones((halfB, halfB))
        Crs = np.kron(layer1, layer2)
        gr = np.ones(B) / B
    else:
        raise ValueError(f"unknown initial condition {initial!r}")
    return gr, Crs


def EM(Ntot, B, links, maxCrs, initial, degreecorrection, itrmax, priorlearning, learningrate,
       *, rng=None, PSIcav0=None, BPconvthreshold=0.000001, float32_logbiprob=True,
       perms=None) -> EMResult:
    """sbm.jl:28-370.

    `initial` is one of the reference's strings or an explicit `(gr, Crs)` pair (what the
    normalizedLaplacian branch, sbm.jl:280-282, hands over).  `PSIcav0` (K x B, rows in `links`
    order) replaces the `rand(1,B)` draws of sbm.jl:319-323.  `perms` optionally fixes the vertex
    order of each sweep (list of 0-based permutations)."""
    rng = rng if rng is not None else np.random.default_rng(0)
    g = _Graph(Ntot, links)
    Ktot = g.Ktot
    links = g.links
    degrees = g.degrees
    ptr, in_link, in_src, rev = g.ptr, g.in_link, g.in_src, g.rev

    cnv = False
    fail = False
    itrnum = 0
    theta = np.ones(Ntot)                                            # sbm.jl:277
    if isinstance(initial, str):
        gr, Crs = initial_parameters(initial, B, Ntot, rng)
    else:
        gr, Crs = (np.array(initial[0], dtype=np.float64).ravel(), np.array(initial[1], dtype=np.float64))
    if not priorlearning:                                            # sbm.jl:304-306
        gr = np.ones(B) / B
    Crs = Crs.copy()
    for r in range(B):                                               # sbm.jl:308-317
        for s in range(B):
            if math.isnan(Crs[r, s]) or abs(Crs[r, s]) == math.inf:
                fail = True
            elif Crs[r, s] < 10 ** (-6.0):
                Crs[r, s] = 10 ** (-6.0)

    if PSIcav0 is None:                                              # sbm.jl:319-323
        PSIcav = rng.random((Ktot, B))
        PSIcav = PSIcav / PSIcav.sum(axis=1, keepdims=True)
    else:
        PSIcav = np.array(PSIcav0, dtype=np.float64).reshape(Ktot, B).copy()
    h = np.zeros(B)
    PSI = np.zeros((Ntot, B))
    st = dict(gr=gr.astype(np.float64).copy(), Crs=Crs, theta=theta, h=h)

    # ---- closures of EM() --------------------------------------------------------------
    def update_h():                                                  # sbm.jl:70-73
        return (st["theta"][:, None] * PSI).sum(axis=0) @ st["Crs"] / Ntot

    def logunnormalizedMessage(i):                                   # sbm.jl:83-98
        th = st["theta"]
        logmessages = np.zeros(B)
        for p in range(ptr[i], ptr[i + 1]):
            s = in_src[p]
            logmessages = logmessages + np.log(th[i] * th[s] * (PSIcav[in_link[p]] @ st["Crs"]))
        grl = st["gr"]
        for r in range(B):
            if math.isnan(grl[r]) or grl[r] < 10 ** (-9.0):
                grl[r] = 1 / Ntot
        grl = grl / grl.sum()
        st["gr"] = grl
        return np.log(grl) - th[i] * st["h"] + logmessages

    def BP(perm):                                                    # sbm.jl:100-126
        conv = 0.0
        st["h"] = update_h()
        th = st["theta"]
        for i in perm:
            logPSIi = logunnormalizedMessage(i)
            for p in range(ptr[i], ptr[i + 1]):
                kji = in_link[p]              # message j -> i
                kij = rev[kji]                # message i -> j
                j = in_src[p]
                cav = logPSIi - np.log(th[j] * th[i] * (PSIcav[kji] @ st["Crs"]))
                PSIcav[kij] = normalize_logprob(cav)
            hprev = th[i] * PSI[i] @ st["Crs"] / Ntot
            grprev = PSI[i] / Ntot
            prev = PSI[i] / Ntot
            logPSIi = logunnormalizedMessage(i)
            PSI[i] = normalize_logprob(logPSIi)
            st["h"] = st["h"] + (th[i] * PSI[i] @ st["Crs"] / Ntot - hprev)
            if priorlearning:
                st["gr"] = st["gr"] + (PSI[i] / Ntot - grprev)
            conv += np.sum(np.abs(PSI[i] / Ntot - prev))
        return conv

    def update_Crs():                                                # sbm.jl:128-148
        th = st["theta"]
        C = st["Crs"]
        Crsnew = np.zeros((B, B))
        for k in range(Ktot):
            i, j = links[k, 0] - 1, links[k, 1] - 1
            ccrs = th[i] * th[j] * np.outer(PSIcav[k], PSIcav[rev[k]]) * C
            Crsnew += ccrs / ccrs.sum()
        grl = st["gr"]
        C = learningrate * Crsnew / (Ntot * np.outer(grl, grl)) + (1 - learningrate) * C
        C = np.where(C > maxCrs, maxCrs, C)
        C = np.where(C < 0, 0.0, C)
        return C

    def update_theta():                                              # sbm.jl:47-64
        block = np.argmax(PSI, axis=1)        # findmax: first maximum
        drmean = np.bincount(block, weights=degrees, minlength=B)
        nr = np.bincount(block, minlength=B).astype(np.float64)
        with np.errstate(invalid="ignore", divide="ignore"):
            drmean = drmean / nr
        return degrees / drmean[block]

    def freeenergy():                                                # sbm.jl:150-213
        th = st["theta"]
        C = st["Crs"]
        sumlogZi = 0.0
        for i in range(Ntot):
            sumlogZi += logsumexp(logunnormalizedMessage(i))
        Ltot = int(0.5 * Ktot)
        logZijs = np.zeros(Ltot)
        CVGPs = np.zeros(Ltot)
        CVGTs = np.zeros(Ltot)
        CVMAPs = np.zeros(Ltot)
        cnt = 0
        with np.errstate(divide="ignore"):
            logC = np.log(C)
        for k in range(Ktot):
            i, j = links[k, 0] - 1, links[k, 1] - 1
            if i < j:
                continue
            a = PSIcav[k]          # PSIcav[indij]
            b = PSIcav[rev[k]]     # PSIcav[indji]
            lti, ltj = math.log(th[i]), math.log(th[j])
            lb = np.log(a)[:, None] + lti + logC + ltj + np.log(b)[None, :]
            if float32_logbiprob:                                    # `logbiprob = Float32[]`, sbm.jl:170
                lbv = lb.astype(np.float32).ravel()
            else:
                lbv = lb.ravel()
            logZijs[cnt] = logsumexp(lbv) - math.log(Ntot)
            w = np.outer(a, b)
            lterm = lti + logC + ltj - math.log(Ntot)
            CVGPs[cnt] = np.sum(w * lterm)                           # sbm.jl:178-184
            CVGTs[cnt] = np.sum(w * th[i] * C * th[j] * lterm) / (Ntot * math.exp(logZijs[cnt]))  # :186-192
            s = int(np.argmax(a))                                    # sbm.jl:194-198
            t = int(np.argmax(b))
            CVMAPs[cnt] = lti + logC[s, t] + ltj - math.log(Ntot)
            cnt += 1
        grl = st["gr"]
        avedegree = grl @ C @ grl
        FE = -((sumlogZi - logZijs.sum()) / Ntot + 0.5 * avedegree)
        var = lambda x: float(np.var(x, ddof=1)) if len(x) > 1 else float("nan")
        return (FE, 1 - logZijs.sum() / Ltot, 1 - CVGPs.sum() / Ltot, 1 - CVGTs.sum() / Ltot,
                1 - CVMAPs.sum() / Ltot, var(logZijs), var(CVGPs), var(CVGTs), var(CVMAPs))

    # ---- main body, sbm.jl:337-367 -----------------------------------------------------
    FE = CVBayes = CVGP = CVGT = CVMAP = 0.0
    varCVBayes = varCVGP = varCVGT = varCVMAP = 0.0
    BPconv = float("nan")
    itrs_done = 0
    if not fail:
        for i in range(Ntot):                                        # updatePSI, sbm.jl:75-81
            PSI[i] = normalize_logprob(logunnormalizedMessage(i))
        st["gr"] = PSI.mean(axis=0)                                  # sbm.jl:341
        for itr in range(1, itrmax + 1):
            perm = perms[itr - 1] if perms is not None else rng.permutation(Ntot)
            BPconv = BP(perm)
            st["Crs"] = update_Crs()
            if degreecorrection:
                st["theta"] = update_theta()
            itrs_done = itr
            if BPconv < BPconvthreshold:
                cnv = True
                itrnum = itr
                break
        st["h"] = update_h()                                         # sbm.jl:364
        st["gr"] = PSI.mean(axis=0)                                  # sbm.jl:365
        (FE, CVBayes, CVGP, CVGT, CVMAP, varCVBayes, varCVGP, varCVGT, varCVMAP) = freeenergy()

    return EMResult(st["gr"], st["Crs"], PSI, FE, CVBayes, CVGP, CVGT, CVMAP, varCVBayes, varCVGP,
                    varCVGT, varCVMAP, fail, cnv, itrnum, PSIcav=PSIcav, theta=st["theta"], h=st["h"],
                    BPconv=BPconv, itrs_done=itrs_done)
