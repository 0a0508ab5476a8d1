"""Host mirror of the labeled SBM of the reference's notebook (src/labeled_sbm/labeled-sbmBIX_0.2.3.8_Julia0.6.2.ipynb,
cell 0): `EM_labeled(...)` has the argument order and the return tuple of the notebook's `EM(...)`; the parts of that
function that stay on the host (initial affinity matrices from Larray / assortativity / Omega, random initial messages,
`labeledsimplegraph`) are restated here, the EM + BP loop and the free energy run in libgraphbix_cuda (gbx_lsbm_*)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import LsbmOptions, LsbmResult, check


def labeledsimplegraph(links: np.ndarray) -> np.ndarray:
    """Notebook `labeledsimplegraph`: labels renumbered 1..G in ascending order of their values, both directions of every
    edge, duplicates and self loops dropped (first occurrence kept, order preserved)."""
    links = np.asarray(links).copy()
    labels = np.unique(links[:, 2])
    out = np.column_stack([links[:, 0], links[:, 1], np.searchsorted(labels, links[:, 2]) + 1]).astype(np.int64)
    both = np.vstack([out, out[:, [1, 0, 2]]])
    _, first = np.unique(both, axis=0, return_index=True)
    both = both[np.sort(first)]
    return both[both[:, 0] != both[:, 1]]


def initial_affinity(Ntot, Ktot, Larray, B, G, dc, REDfrac, assortativity, Omega) -> np.ndarray:
    """Notebook EM, "initial state": (cm - delta * Omega) * ones + delta * eye, over Norm = Ntot (no degree correction) or
    Ktot; delta = 0.99 cm / Omega for an assortative label, -0.99 cm / (1 - Omega) for a disassortative one."""
    cm = 2 * np.asarray(Larray, dtype=np.float64) / Ntot
    Norm = Ntot if not dc else Ktot
    aff = np.zeros((G, B, B))
    last = G if REDfrac == 0 else G - 1
    for k in range(last):
        a = assortativity[k] if k < len(assortativity) else ""
        delta = 0.99 * cm[k] / Omega if a == "a" else (-0.99 * cm[k] / (1 - Omega) if a == "d" else 0.0)
        aff[k] = ((cm[k] - delta * Omega) * np.ones((B, B)) + delta * np.eye(B)) / Norm
    if REDfrac != 0:
        aff[G - 1] = cm[G - 1] * np.ones((B, B)) / Norm
    return aff


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


class LabeledEngine:
    """One labeled graph + one EM state resident on one GPU (C ABI gbx_lsbm_*)."""

    def __init__(self, Ntot: int, links3, B: int, G: int, device: int = 0, stream=None):
        self.lib = _lib.load()
        links3 = np.asarray(links3)
        if links3.ndim != 2 or links3.shape[1] != 3:
            raise ValueError("links must be K x 3 (i, j, label)")
        self.n, self.q, self.G, self.K = int(Ntot), int(B), int(G), int(links3.shape[0])
        colmajor = np.asfortranarray(links3, dtype=np.int64)
        self._h = C.c_void_p()
        check(self.lib.gbx_lsbm_create(C.byref(self._h), int(device), self.n, self.K,
                                       colmajor.ctypes.data_as(C.POINTER(C.c_int64)), self.q, self.G))
        self.opt = LsbmOptions()
        self.lib.gbx_lsbm_default_options(C.byref(self.opt))
        if stream is not None:
            ptr = getattr(stream, "cuda_stream", stream)
            check(self.lib.gbx_lsbm_set_stream(self._h, C.c_void_p(int(ptr) if ptr else 0)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.gbx_lsbm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_options(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.opt, k):
                raise TypeError(f"unknown option {k!r}")
            setattr(self.opt, k, v)
        check(self.lib.gbx_lsbm_set_options(self._h, C.byref(self.opt)))

    def init(self, gr, affinity, psicav=None, seed=None):
        """psicav: K x q initial messages (rows in the order of links3), or None: drawn on the device from `seed`."""
        gr = np.ascontiguousarray(gr, dtype=np.float64).ravel()
        affinity = np.ascontiguousarray(affinity, dtype=np.float64).reshape(self.G, self.q, self.q)
        if psicav is None:
            sd = (C.c_uint64 * 1)(int(seed or 0))
            check(self.lib.gbx_lsbm_init_random(self._h, _dp(gr), _dp(affinity), sd))
            return
        psicav = np.ascontiguousarray(psicav, dtype=np.float64).reshape(self.K, self.q)
        check(self.lib.gbx_lsbm_init(self._h, _dp(gr), _dp(affinity), _dp(psicav)))

    def iterate(self, n: int = 1, read_conv: bool = True):
        conv = C.c_double(float("nan"))
        check(self.lib.gbx_lsbm_iterate(self._h, int(n), C.byref(conv) if read_conv else None))
        return conv.value

    def em(self) -> LsbmResult:
        res = LsbmResult()
        check(self.lib.gbx_lsbm_em(self._h, C.byref(res)))
        return res

    def finalize(self) -> LsbmResult:
        res = LsbmResult()
        check(self.lib.gbx_lsbm_finalize(self._h, C.byref(res)))
        return res

    def state(self):
        q, G = self.q, self.G
        gr, aff, P, h = np.empty(q), np.empty((G, q, q)), np.empty((self.n, q)), np.empty((G, q))
        check(self.lib.gbx_lsbm_get_state(self._h, _dp(gr), _dp(aff), _dp(P), _dp(h)))
        return dict(gr=gr, affinity=aff, PSI=P, h=h)

    def messages(self):
        m = np.empty((self.K, self.q))
        check(self.lib.gbx_lsbm_get_messages(self._h, _dp(m)))
        return m

    @property
    def launch_count(self):
        return int(self.lib.gbx_lsbm_launch_count(self._h))


def EM_labeled(Ntot, Larray, B, G, links, itrmax, BPconvthreshold, learning, priorlearning, dc, learningrate, REDfrac,
               assortativity, Omega, *, rng=None, PSIcav0=None, affinity0=None, gr0=None, damping=1.0, device=0,
               return_engine=False, seed=None):
    """Drop-in for the notebook's EM(); returns its 14-tuple
    (gr, affinity, PSI, FE, CVBayes, CVGP, CVGT, CVMAP, varCVBayes, varCVGP, varCVGT, varCVMAP, cnv, itrnum) with
    `affinity` as a (G, B, B) array."""
    rng = rng if rng is not None else np.random.default_rng()
    links = np.asarray(links, dtype=np.int64)
    K = links.shape[0]
    eng = LabeledEngine(Ntot, links, B, G, device=device)
    eng.set_options(dc=int(bool(dc)), learning=int(bool(learning)), priorlearning=int(bool(priorlearning)),
                    learningrate=float(learningrate), itrmax=int(itrmax), bp_conv_threshold=float(BPconvthreshold),
                    REDfrac=float(REDfrac), damping=float(damping))
    aff = affinity0 if affinity0 is not None else initial_affinity(Ntot, K, Larray, B, G, dc, REDfrac, assortativity, Omega)
    gr = np.ones(B) / B if gr0 is None else gr0
    if PSIcav0 is None and seed is None:     # rand(1,B) normalised, one row per link
        PSIcav0 = rng.random((K, B))
        PSIcav0 /= PSIcav0.sum(axis=1, keepdims=True)
    eng.init(gr, aff, PSIcav0, seed=seed)    # seed given: the same draw on the device
    r = eng.em()
    st = eng.state()
    out = (st["gr"], st["affinity"], st["PSI"], r.FE, r.CVBayes, r.CVGP, r.CVGT, r.CVMAP, r.varCVBayes, r.varCVGP,
           r.varCVGT, r.varCVMAP, bool(r.cnv), int(r.itrnum))
    if return_engine:
        return out, eng, r
    eng.close()
    return out


class LabeledBatch:
    """Many labeled graphs in ONE handle on one GPU (gbx_lsbm_create_batch): every kernel launch serves all of them, each
    graph keeps its own parameters and converges on its own (BASELINE configs[4])."""

    def __init__(self, graphs, B: int, G: int, device: int = 0, stream=None):
        """graphs: list of (Ntot, links3) with links3 K x 3 (i, j, label), vertices 1..Ntot of that graph."""
        self.lib = _lib.load()
        self.q, self.G, self.count = int(B), int(G), len(graphs)
        self.n = [int(n) for n, _ in graphs]
        self._links = [np.asfortranarray(np.asarray(l), dtype=np.int64) for _, l in graphs]   # kept alive for the call
        self.K = [int(l.shape[0]) for l in self._links]
        ns = (C.c_int64 * self.count)(*self.n)
        Ks = (C.c_int64 * self.count)(*self.K)
        ptrs = (C.POINTER(C.c_int64) * self.count)(*[l.ctypes.data_as(C.POINTER(C.c_int64)) for l in self._links])
        self._h = C.c_void_p()
        check(self.lib.gbx_lsbm_create_batch(C.byref(self._h), int(device), self.count, ns, Ks, ptrs, self.q, self.G))
        self._links = None
        self.opt = LsbmOptions()
        self.lib.gbx_lsbm_default_options(C.byref(self.opt))
        if stream is not None:
            ptr = getattr(stream, "cuda_stream", stream)
            check(self.lib.gbx_lsbm_set_stream(self._h, C.c_void_p(int(ptr) if ptr else 0)))

    close = LabeledEngine.close
    __del__ = LabeledEngine.__del__
    set_options = LabeledEngine.set_options

    def init(self, gr, affinity, psicav=None, seeds=None):
        """gr: count x q, affinity: count x G x q x q, psicav: list of K_g x q arrays (rows in the order of links3), or
        None: the messages of graph g are drawn on the device from seeds[g]."""
        gr = np.ascontiguousarray(gr, dtype=np.float64).reshape(self.count, self.q)
        affinity = np.ascontiguousarray(affinity, dtype=np.float64).reshape(self.count, self.G, self.q, self.q)
        if psicav is None:
            sd = (C.c_uint64 * self.count)(*[int(x) for x in (seeds if seeds is not None else range(self.count))])
            check(self.lib.gbx_lsbm_init_random(self._h, _dp(gr), _dp(affinity), sd))
            return
        rows = [np.ascontiguousarray(m, dtype=np.float64).reshape(k, self.q) for m, k in zip(psicav, self.K)]
        ptrs = (C.POINTER(C.c_double) * self.count)(*[_dp(m) for m in rows])
        check(self.lib.gbx_lsbm_init_batch(self._h, _dp(gr), _dp(affinity), ptrs))

    def iterate(self, n: int = 1):
        check(self.lib.gbx_lsbm_iterate(self._h, int(n), None))

    def em(self):
        res = (LsbmResult * self.count)()
        check(self.lib.gbx_lsbm_em(self._h, res))
        return list(res)

    def finalize(self):
        res = (LsbmResult * self.count)()
        check(self.lib.gbx_lsbm_finalize(self._h, res))
        return list(res)

    def state(self, g: int):
        q, G = self.q, self.G
        gr, aff, P, h = np.empty(q), np.empty((G, q, q)), np.empty((self.n[g], q)), np.empty((G, q))
        check(self.lib.gbx_lsbm_get_state_of(self._h, int(g), _dp(gr), _dp(aff), _dp(P), _dp(h)))
        return dict(gr=gr, affinity=aff, PSI=P, h=h)

    def messages(self, g: int):
        m = np.empty((self.K[g], self.q))
        check(self.lib.gbx_lsbm_get_messages_of(self._h, int(g), _dp(m)))
        return m

    @property
    def launch_count(self):
        return int(self.lib.gbx_lsbm_launch_count(self._h))


def EM_labeled_batch(graphs, itrmax, BPconvthreshold, learning, priorlearning, dc, learningrate, REDfrac, assortativity, *,
                     rngs=None, device=0, max_concurrent=0, PSIcav0=None, seeds=None, return_engine=False):
    """The notebook's EM() on many graphs at once on one GPU (one handle for all of them, gbx_lsbm_create_batch; BASELINE
    configs[4]: a detectability sweep over thousands of N = 10k graphs).  `graphs` is a list of (Ntot, Larray, B, G, links)
    with equal B and G; Omega is 1 / B as in the notebook's MAIN.  Returns the list of EM()'s 14-tuples, in order; each
    equals what EM_labeled returns for that graph from the same initial messages (to the rounding of the sums).
    Initial messages: PSIcav0[k] if given, else drawn on the device from seeds[k] if `seeds` is given, else from rngs[k]."""
    if not graphs:
        return []
    B, G = int(graphs[0][2]), int(graphs[0][3])
    if any(int(g[2]) != B or int(g[3]) != G for g in graphs):
        raise ValueError("the graphs of a batch must share B and G")
    links = [np.asarray(g[4], dtype=np.int64) for g in graphs]
    eng = LabeledBatch([(g[0], l) for g, l in zip(graphs, links)], B, G, device=device)
    outs = []
    try:
        eng.set_options(dc=int(bool(dc)), learning=int(bool(learning)), priorlearning=int(bool(priorlearning)),
                        learningrate=float(learningrate), itrmax=int(itrmax), bp_conv_threshold=float(BPconvthreshold),
                        REDfrac=float(REDfrac))
        psi, aff = [], []
        for k, (Ntot, Larray, _, _, _) in enumerate(graphs):
            K = links[k].shape[0]
            aff.append(initial_affinity(Ntot, K, Larray, B, G, dc, REDfrac, assortativity, 1 / B))
            if PSIcav0 is None and seeds is not None:
                continue
            if PSIcav0 is not None:
                m = np.asarray(PSIcav0[k], dtype=np.float64)
            else:
                rng = rngs[k] if rngs is not None else np.random.default_rng()
                m = rng.random((K, B))
                m /= m.sum(axis=1, keepdims=True)
            psi.append(m)
        eng.init(np.tile(np.ones(B) / B, (len(graphs), 1)), np.stack(aff), psi if psi else None, seeds=seeds)
        res = eng.em()
        for g, r in enumerate(res):
            st = eng.state(g)
            outs.append((st["gr"], st["affinity"], st["PSI"], r.FE, r.CVBayes, r.CVGP, r.CVGT, r.CVMAP, r.varCVBayes,
                         r.varCVGP, r.varCVGT, r.varCVMAP, bool(r.cnv), int(r.itrnum)))
        if return_engine:
            return outs, eng, res
    finally:
        if not return_engine:
            eng.close()
    return outs
