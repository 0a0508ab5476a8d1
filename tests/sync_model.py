"""Executable specification (numpy, log domain) of the SYNCHRONOUS schedule the CUDA path runs.

Test helper only.  It states, sweep by sweep, what `libgraphbix_cuda.so` computes, so the
kernels can be checked per sweep to ~1e-12 (tests/test_gpu_sweep.py) -- a much tighter
check than the fixed-point comparison with the asynchronous oracle (oracle/sbm_oracle.py).

One EM iteration (mirrors sbm.jl:342-347 with a Jacobi sweep in place of `BP()`):
    gr   <- clamp/renormalise(gr)                      (sbm.jl:89-96)
    h    <- sum_i theta_i PSI_i * Crs / N              (sbm.jl:70-73)
    for all i at once, from the OLD messages:
        logPSI_i = log gr - theta_i h + sum_s log(theta_i theta_s psi_{s->i} Crs)   (sbm.jl:83-98)
        psi'_{i->j} = normalize_logprob(logPSI_i - log(theta_i theta_j psi_{j->i} Crs)) (sbm.jl:108-109)
        psi_{i->j} <- (1-damping) psi_{i->j} + damping psi'_{i->j}
        PSI'_i = normalize_logprob(logPSI_i)                                        (sbm.jl:117-118)
    conv = sum |PSI' - PSI| / N                         (sbm.jl:123)
    gr  += mean(PSI') - mean(PSI)   if priorlearning    (sbm.jl:120-122)
    Crs <- update_Crs()                                 (sbm.jl:128-148; new outgoing x previous incoming message)
    theta <- update_theta() if degreecorrection         (sbm.jl:47-64)
"""
from __future__ import annotations

import math

import numpy as np

LOG_1EM8 = math.log(10 ** (-8.0))


def build_csr(n, links):
    """Destination-major CSR: slot e of row i holds the message col[e] -> i."""
    links = np.asarray(links, dtype=np.int64)
    src = links[:, 0] - 1
    dst = links[:, 1] - 1
    order = np.lexsort((src, dst))          # slot -> link id
    rowptr = np.concatenate(([0], np.cumsum(np.bincount(dst, minlength=n))))
    col = src[order]
    row = dst[order]
    slot_of_link = np.empty_like(order)
    slot_of_link[order] = np.arange(order.size)
    # rev[e]: slot of the opposite message (row = col[e], col = row[e])
    key = row * n + col
    rkey = col * n + row
    rev = np.searchsorted(key, rkey)        # key is sorted because of the lexsort
    assert np.all(key[rev] == rkey)
    return dict(n=n, rowptr=rowptr, col=col, row=row, rev=rev, order=order, slot_of_link=slot_of_link)


def _normalize_logprob_rows(x):
    x = np.maximum(x, LOG_1EM8)
    m = x.max(axis=1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=1, keepdims=True)


def clamp_gr(gr, n):
    gr = np.array(gr, dtype=np.float64)
    bad = np.isnan(gr) | (gr < 10 ** (-9.0))
    gr[bad] = 1.0 / n
    return gr / gr.sum()


def vertex_log(g, psi, theta, C, gr, h):
    """logT (per slot) and logPSI (per vertex) from messages in slot order."""
    n = g["n"]
    T = (psi @ C) * (theta[g["row"]] * theta[g["col"]])[:, None]
    with np.errstate(divide="ignore"):
        logT = np.log(T)
    agg = np.zeros((n, psi.shape[1]))
    np.add.at(agg, g["row"], logT)
    logPSI = np.log(gr)[None, :] - theta[:, None] * h[None, :] + agg
    return logT, logPSI


def sweep(g, psi, theta, C, gr, h, damping=1.0):
    logT, logPSI = vertex_log(g, psi, theta, C, gr, h)
    cav = logPSI[g["row"]] - logT                     # message row[e] -> col[e], lives at slot rev[e]
    out = _normalize_logprob_rows(cav)
    new = np.empty_like(psi)
    new[g["rev"]] = out
    if damping != 1.0:
        new = (1.0 - damping) * psi + damping * new
    return new, _normalize_logprob_rows(logPSI), logPSI


def update_Crs(g, psi_new, psi_old, theta, C, gr, n, maxCrs, lr):
    """sbm.jl:128-148 on the pairs the sweep has in hand: for the link (i, j) handled at slot e (row i, col j) the
    NEW message i -> j (just emitted, stored at rev[e]) and the message j -> i it was computed from.  Summed over
    all directed links and symmetrised (with all-new messages the sum is symmetric by itself); same fixed point."""
    a = psi_new[g["rev"]]      # message row -> col   (PSIcav[indij] with i = row)
    b = psi_old                # message col -> row
    w = (theta[g["row"]] * theta[g["col"]])[:, None, None] * a[:, :, None] * b[:, None, :] * C[None]
    w = w / w.sum(axis=(1, 2), keepdims=True)
    Cn = w.sum(axis=0)
    Cn = 0.5 * (Cn + Cn.T)
    C2 = lr * Cn / (n * np.outer(gr, gr)) + (1 - lr) * C
    return np.clip(C2, 0.0, maxCrs)


def update_theta(PSI, degrees):
    B = PSI.shape[1]
    block = np.argmax(PSI, axis=1)
    drmean = np.bincount(block, weights=degrees, minlength=B)
    nr = np.bincount(block, minlength=B)
    with np.errstate(invalid="ignore", divide="ignore"):
        drmean = drmean / nr
    return degrees / drmean[block]


class SyncSBM:
    def __init__(self, n, links, q, gr0, Crs0, psi0_links, *, dc=True, prior=True, lr=0.3,
                 maxCrs=None, damping=1.0, mstep_exact=False):
        self.g = build_csr(n, links)
        self.n, self.q = n, q
        self.dc, self.prior, self.lr, self.damping = dc, prior, lr, damping
        self.mstep_exact = mstep_exact       # gbx_sbm_options::mstep_exact: pair the NEW messages of both directions
        self.maxCrs = float(n if maxCrs is None else maxCrs)
        self.degrees = np.diff(self.g["rowptr"]).astype(np.float64)
        self.theta = np.ones(n)
        self.C = np.maximum(np.array(Crs0, dtype=np.float64), 10 ** (-6.0))
        self.gr = np.ones(q) / q if not prior else np.array(gr0, dtype=np.float64).ravel().copy()
        self.psi = np.array(psi0_links, dtype=np.float64)[self.g["order"]].copy()
        self.h = np.zeros(q)
        self.gr = clamp_gr(self.gr, n)
        _, self.PSI, _ = sweep(self.g, self.psi, self.theta, self.C, self.gr, self.h)
        self.gr = self.PSI.mean(axis=0)
        self.itr = 0
        self.conv = float("nan")

    def update_h(self):
        return (self.theta[:, None] * self.PSI).sum(axis=0) @ self.C / self.n

    def iterate(self):
        self.gr = clamp_gr(self.gr, self.n)
        self.h = self.update_h()
        psi, PSI, _ = sweep(self.g, self.psi, self.theta, self.C, self.gr, self.h, self.damping)
        self.conv = float(np.abs(PSI - self.PSI).sum() / self.n)
        if self.prior:
            self.gr = self.gr + (PSI.mean(axis=0) - self.PSI.mean(axis=0))
        old = self.psi
        self.psi, self.PSI = psi, PSI
        self.C = update_Crs(self.g, self.psi, self.psi if self.mstep_exact else old, self.theta, self.C, self.gr, self.n,
                            self.maxCrs, self.lr)
        if self.dc:
            self.theta = update_theta(self.PSI, self.degrees)
        self.itr += 1
        return self.conv

    def run(self, itrmax=256, tol=1e-6):
        for _ in range(itrmax):
            if self.iterate() < tol:
                return True
        return False

    def messages_in_link_order(self):
        return self.psi[self.g["slot_of_link"]]

    def finalize(self):
        """sbm.jl:364-366 and freeenergy(), sbm.jl:150-213 (float64 throughout)."""
        g, n = self.g, self.n
        self.h = self.update_h()
        self.gr = self.PSI.mean(axis=0)
        gr = clamp_gr(self.gr, n)
        self.gr = gr
        _, logPSI = vertex_log(g, self.psi, self.theta, self.C, gr, self.h)
        m = logPSI.max(axis=1)
        sumlogZi = float(np.sum(m + np.log(np.exp(np.maximum(logPSI - m[:, None], -500.0)).sum(axis=1))))
        sel = g["row"] > g["col"]
        a = self.psi[g["rev"]][sel]
        b = self.psi[sel]
        ti = self.theta[g["row"][sel]]
        tj = self.theta[g["col"][sel]]
        with np.errstate(divide="ignore"):
            logC = np.log(self.C)
        w = a[:, :, None] * b[:, None, :]
        Z = (w * self.C[None]).sum(axis=(1, 2)) * ti * tj / n
        logZ = np.log(Z)
        lterm = (np.log(ti) + np.log(tj) - math.log(n))[:, None, None] + logC[None]
        cvgp = (w * lterm).sum(axis=(1, 2))
        cvgt = (w * (ti * tj)[:, None, None] * self.C[None] * lterm).sum(axis=(1, 2)) / (n * Z)
        s = np.argmax(a, axis=1)
        t = np.argmax(b, axis=1)
        cvmap = np.log(ti) + np.log(tj) - math.log(n) + logC[s, t]
        L = int(sel.sum())
        FE = -((sumlogZi - logZ.sum()) / n + 0.5 * float(gr @ self.C @ gr))
        v = lambda x: float(np.var(x, ddof=1))
        return dict(FE=FE, CVBayes=1 - logZ.sum() / L, CVGP=1 - cvgp.sum() / L, CVGT=1 - cvgt.sum() / L,
                    CVMAP=1 - cvmap.sum() / L, varCVBayes=v(logZ), varCVGP=v(cvgp), varCVGT=v(cvgt),
                    varCVMAP=v(cvmap))
