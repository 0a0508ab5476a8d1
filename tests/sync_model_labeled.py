"""Executable specification (numpy) of the SYNCHRONOUS schedule the CUDA path of the labeled SBM runs (gbx_lsbm.cu):
the notebook's EM() with a Jacobi sweep in place of the asynchronous BP().  Test helper only.

One iteration:  h = update_h();  for all i at once from the old messages: cavity messages and marginals;
conv = sum |PSI' - PSI| / N;  gr = mean(PSI');  affinity = update_affinity()  (the notebook's learning-rate blends mix the new value with itself)."""
import math

import numpy as np

LOG_1EM8 = math.log(10 ** (-8.0))


def _normalize_rows(x):
    x = np.maximum(x, LOG_1EM8)
    m = x.max(axis=1, keepdims=True)
    e = np.exp(np.maximum(x, m - 500.0) - m)
    return e / e.sum(axis=1, keepdims=True)


class SyncLabeled:
    def __init__(self, n, links3, B, G, gr0, aff0, psi0, *, dc=True, learning=True, prior=True, lr=0.3, red=False):
        links3 = np.asarray(links3, dtype=np.int64)
        self.n, self.B, self.G = n, B, G
        self.src, self.dst, self.lab = links3[:, 0] - 1, links3[:, 1] - 1, links3[:, 2] - 1
        self.K = len(links3)
        key = self.src * n + self.dst
        order = np.argsort(key)
        self.rev = order[np.searchsorted(key[order], self.dst * n + self.src)]   # link (j, i) of link (i, j)
        d = np.zeros((n, G))
        np.add.at(d, (self.dst, self.lab), 1.0)
        self.deg = d if dc else np.ones((n, G))
        self.dc, self.learning, self.prior, self.lr, self.red = dc, learning, prior, lr, red
        self.gr = np.array(gr0, dtype=np.float64).ravel().copy()
        self.aff = np.array(aff0, dtype=np.float64).copy()
        self.psi = np.array(psi0, dtype=np.float64).copy()      # row k: message links[k,0] -> links[k,1]
        self.h = np.zeros((G, B))
        self.PSI = self._marginals()[0]
        self.h = self.update_h()
        self.conv = float("nan")

    def update_h(self):
        return np.stack([(self.deg[:, a][:, None] * (self.PSI @ self.aff[a])).sum(axis=0) for a in range(self.G)])

    def _terms(self):
        """log(d d N psi A) of every message k (as received by its target dst[k])."""
        dd = self.deg[self.dst, self.lab] * self.deg[self.src, self.lab]
        T = np.einsum("kr,krt->kt", self.psi, self.aff[self.lab]) * (dd * self.n)[:, None]
        with np.errstate(divide="ignore"):
            return np.log(T)

    def _marginals(self):
        lt = self._terms()
        agg = np.zeros((self.n, self.B))
        np.add.at(agg, self.dst, lt)
        x = agg + np.log(self.gr)[None, :] - self.deg @ self.h
        return _normalize_rows(x), x, lt

    def iterate(self):
        self.h = self.update_h()
        PSI, x, lt = self._marginals()
        # message i -> j (link k = (i, j)) from the vertex aggregate of i minus the term of the message j -> i
        cav = x[self.src] - lt[self.rev]
        self.psi = _normalize_rows(cav)
        self.conv = float(np.abs(PSI - self.PSI).sum() / self.n)
        self.PSI = PSI
        if self.prior:
            g = self.PSI.mean(axis=0)                       # update_gr rebinds gr: the blend mixes new with new
            self.gr = self.lr * g + (1 - self.lr) * g
        if self.learning:
            a = self.update_affinity()                      # mutates and returns the same Dict in the notebook
            self.aff = self.lr * a + (1 - self.lr) * a
        return self.conv

    def update_affinity(self):
        sel = self.src < self.dst
        a, b = self.psi[sel], self.psi[self.rev[sel]]
        lab = self.lab[sel]
        w = a[:, :, None] * b[:, None, :] * self.aff[lab]
        w = w / w.sum(axis=(1, 2), keepdims=True)
        new = np.zeros((self.G, self.B, self.B))
        np.add.at(new, lab, w)
        denom = self.deg.T @ self.PSI
        out = np.zeros_like(new)
        for g in range(self.G):
            out[g] = (new[g] + new[g].T) / np.outer(denom[g], denom[g])
            out[g][out[g] < 10 ** (-8.0)] = 10 ** (-8.0)
        if self.red:
            out[self.G - 1] = 1.0 / self.K
        return out

    def finalize(self):
        _, x, _ = self._marginals()
        m = x.max(axis=1)
        sumlogZi = float(np.sum(m + np.log(np.exp(np.maximum(x - m[:, None], -500.0)).sum(axis=1))))
        sel = self.src > self.dst
        a, b, lab = self.psi[sel], self.psi[self.rev[sel]], self.lab[sel]
        dd = self.deg[self.src[sel], lab] * self.deg[self.dst[sel], lab]
        X = dd[:, None, None] * self.aff[lab]
        w = a[:, :, None] * b[:, None, :]
        Z = (X * w).sum(axis=(1, 2))
        logZ = np.log(Z)
        gp = (w * np.log(X)).sum(axis=(1, 2))
        gt = (w * X * np.log(X)).sum(axis=(1, 2)) / Z
        s, t = np.argmax(a, axis=1), np.argmax(b, axis=1)
        mp = np.log(X[np.arange(len(s)), s, t])
        L = int(sel.sum())
        v = lambda y: float(np.var(y, ddof=1))
        return dict(FE=-(sumlogZi - logZ.sum()) / self.n - L / self.n, CVBayes=1 - logZ.sum() / L, CVGP=1 - gp.sum() / L,
                    CVGT=1 - gt.sum() / L, CVMAP=1 - mp.sum() / L, varCVBayes=v(logZ), varCVGP=v(gp), varCVGT=v(gt),
                    varCVMAP=v(mp))
