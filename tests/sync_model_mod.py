"""Executable specification (numpy) of the SYNCHRONOUS schedule the CUDA path runs for mod.jl's EM()
(src/mod.jl:3-244).  Test helper only; see tests/sync_model.py for the sbm.jl counterpart.

One EM iteration:
    h <- degrees' PSI                                                              mod.jl:22-24
    for all i at once, from the OLD messages:
        logPSI_i = sum_s log(1 + psi_{s->i} (e^beta - 1)) - alpha beta d_i h + log gr      mod.jl:38-52
        psi'_{i->j} = normalize_logprob(logPSI_i - log(1 + psi_{j->i}(e^beta - 1)))        mod.jl:62-63
        PSI'_i = normalize_logprob(logPSI_i)                                              mod.jl:68-69
    conv = sum |PSI' - PSI| / N ; gr <- mean(PSI') if priorlearning                  mod.jl:72, 221-223
    (omega_in, omega_out) <- update_omega() (the blend of mod.jl:226-227 mixes new with new) ; (alpha, beta)  mod.jl:224-229
"""
from __future__ import annotations

import math

import numpy as np

from sync_model import _normalize_logprob_rows, build_csr


class SyncMOD:
    def __init__(self, n, links, q, gr, psi0_links, *, dc=True, learning=True, prior=False, lr=0.3, betafrac=0.0,
                 maxomega=1.0, damping=1.0):
        self.g = build_csr(n, links)
        self.n, self.q, self.K = n, q, len(self.g["col"])
        self.dc, self.learning, self.prior, self.lr, self.maxomega, self.damping = dc, learning, prior, lr, maxomega, damping
        self.nbdeg = np.diff(self.g["rowptr"]).astype(np.float64)
        d = self.nbdeg
        self.excm = float(d @ d) / (n * d.mean()) - 1
        beta0 = math.log(1 + q / (self.excm - 1))
        betaast = math.log(1 + q / (math.sqrt(self.excm) - 1))
        self.alpha = 1 / self.K
        self.beta = betaast - betafrac * (betaast - beta0)
        self.omegain = math.exp(self.beta) * self.alpha * self.beta / (math.exp(self.beta) - 1)
        self.omegaout = self.alpha * self.beta / (math.exp(self.beta) - 1)
        self.degrees = d.copy() if dc else np.ones(n)
        self.gr = np.array(gr, dtype=np.float64).ravel().copy()
        self.psi = np.array(psi0_links, dtype=np.float64)[self.g["order"]].copy()
        self.h = np.zeros(q)
        _, self.PSI = self._sweep(self.psi, 1.0)
        self.itr = 0
        self.conv = float("nan")

    def _logs(self, psi):
        g = self.g
        logT = np.log1p(psi * (math.exp(self.beta) - 1)) if False else np.log(1.0 + psi * (math.exp(self.beta) - 1))
        agg = np.zeros((self.n, self.q))
        np.add.at(agg, g["row"], logT)
        field = self.alpha * self.beta * self.h[None, :]
        if self.dc:
            field = field * self.nbdeg[:, None]
        return logT, agg - field + np.log(self.gr)[None, :]

    def _sweep(self, psi, damping):
        g = self.g
        logT, logPSI = self._logs(psi)
        out = _normalize_logprob_rows(logPSI[g["row"]] - logT)
        new = np.empty_like(psi)
        new[g["rev"]] = out
        if damping != 1.0:
            new = (1.0 - damping) * psi + damping * new
        return new, _normalize_logprob_rows(logPSI)

    def update_h(self):
        return self.degrees @ self.PSI

    def iterate(self):
        g = self.g
        self.h = self.update_h()
        psi, PSI = self._sweep(self.psi, self.damping)
        self.conv = float(np.abs(PSI - self.PSI).sum() / self.n)
        old = self.psi
        self.psi, self.PSI = psi, PSI
        if self.prior:
            self.gr = self.PSI.mean(axis=0)
        if self.learning:
            # mod.jl:79-88 on the pairs the sweep has in hand: new message row -> col with the message col -> row it
            # was computed from, over all directed slots, i.e. twice per undirected edge (hence the 0.5)
            s = (self.psi[g["rev"]] * old).sum(axis=1)
            oin, oout = self.omegain, self.omegaout
            omegainnew = 0.5 * float(np.sum(oin * s / ((oin - oout) * s + oout)))
            dn = float(np.linalg.norm(self.update_h()) ** 2)
            oin_n = 2 * omegainnew / dn
            oout_n = 2 * (0.5 * self.K - omegainnew) / ((self.K ** 2 if self.dc else self.n ** 2) - dn)
            if oin_n > self.maxomega:
                oin_n = self.maxomega
            elif oin_n < 0:
                oin_n = 0
            elif oout_n < 0:
                oout_n = 0
            # mod.jl:226-227 after the closure write-through of update_omega (mod.jl:90-94): new blended with itself
            self.omegain = self.lr * oin_n + (1 - self.lr) * oin_n
            self.omegaout = self.lr * oout_n + (1 - self.lr) * oout_n
            self.beta = math.log(self.omegain / self.omegaout)
            self.alpha = (self.omegain - self.omegaout) / self.beta
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
        g, n = self.g, self.n
        self.h = self.update_h()
        _, logPSI = self._logs(self.psi)
        m = logPSI.max(axis=1)
        sumlogZi = float(np.sum(m + np.log(np.exp(np.maximum(logPSI - m[:, None], -500.0)).sum(axis=1))))
        sel = g["row"] > g["col"]
        a, b = self.psi[g["rev"]][sel], self.psi[sel]
        di, dj = self.degrees[g["row"][sel]], self.degrees[g["col"][sel]]
        s = (a * b).sum(axis=1)
        al, be, oin, oout = self.alpha, self.beta, self.omegain, self.omegaout
        logZ = np.log(1 + (math.exp(be) - 1) * s)
        cvbp = logZ + np.log(di) + math.log(oout) + np.log(dj)
        cvgp = be * s
        cvgt = di * dj * (oout * math.log(oout) + (oin * math.log(oin) - oout * math.log(oout)) * s) / np.exp(cvbp)
        cvmap = np.where(np.argmax(a, axis=1) == np.argmax(b, axis=1), math.log(oin), math.log(oout))
        L = int(sel.sum())
        constshift = float(np.sum(self.degrees * np.log(self.degrees))) / L
        hn2 = float(np.linalg.norm(self.h) ** 2)
        FE = (-(sumlogZi - logZ.sum() + 0.5 * al * be * hn2) / (be * n)
              - (L / n) * (2 * L * oout + math.log(oout) + constshift) / be)
        v = lambda x: float(np.var(x, ddof=1))
        return dict(FE=FE, CVBayes=-cvbp.sum() / L + 1, CVGP=-cvgp.sum() / L - math.log(oout) + 1 - constshift,
                    CVGT=-cvgt.sum() / L + 1 - constshift, CVMAP=-cvmap.sum() / L + 1 - constshift,
                    varCVBayes=v(logZ), varCVGP=v(cvgp), varCVGT=v(cvgt), varCVMAP=v(cvmap))
