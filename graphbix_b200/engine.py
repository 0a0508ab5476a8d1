"""Host-side mirror of the reference's `EM()` (src/sbm.jl:28-370) on top of the C ABI.

`EM(Ntot, B, links, maxCrs, initial, degreecorrection, itrmax, priorlearning, learningrate)` keeps the
reference's argument order and returns the same 15-tuple
`(gr, Crs, PSI, FE, CVBayes, CVGP, CVGT, CVMAP, varCVBayes, varCVGP, varCVGT, varCVMAP, fail, cnv, itrnum)`
(sbm.jl:369).  `SbmEngine` is the thin object around one `gbx_sbm` handle for callers that want the pieces
(benchmarks, tests, the q-sweep driver).
"""
from __future__ import annotations

import ctypes as C
import sys
import time

import numpy as np

from . import _lib
from ._lib import ModOptions, ModResult, SbmOptions, SbmResult, check


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _host_array(shape, dtype=np.float64):
    """Result buffer for a device -> host copy.  Large arrays come from PyTorch's caching pinned-memory allocator when the
    program has torch loaded and a CUDA device (a 64 MB PSI then arrives at PCIe speed, ~1.2 ms instead of ~4 ms into
    pageable memory; the allocator reuses the blocks of arrays that were dropped); otherwise plain numpy."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    torch = sys.modules.get("torch")
    if torch is not None and nbytes >= (1 << 20):
        try:
            tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32}[np.dtype(dtype)]
            return torch.empty(tuple(shape) if not np.isscalar(shape) else (int(shape),), dtype=tdt, pin_memory=True).numpy()
        except Exception:
            pass
    return np.empty(shape, dtype=dtype)


class SbmEngine:
    """One graph + one EM state resident on one GPU."""

    def __init__(self, Ntot: int, links, B: int, device: int = 0, stream=None):
        self.lib = _lib.load()
        links = np.asarray(links)
        if links.ndim != 2 or links.shape[1] != 2:
            raise ValueError("links must be K x 2")
        self.n, self.q, self.K = int(Ntot), int(B), int(links.shape[0])
        self.timings = {}
        t0 = time.perf_counter()
        colmajor = np.asfortranarray(links, dtype=np.int64)  # Julia's memory layout of the K x 2 matrix
        self._h = C.c_void_p()
        check(self.lib.gbx_sbm_create(C.byref(self._h), int(device), self.n, self.K,
                                      colmajor.ctypes.data_as(C.POINTER(C.c_int64)), self.q))
        self.timings["create"] = time.perf_counter() - t0
        self.opt = SbmOptions()
        self.lib.gbx_sbm_default_options(C.byref(self.opt))
        if stream is not None:
            self.set_stream(stream)

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.gbx_sbm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- configuration ----------------------------------------------------------------------
    def set_stream(self, stream):
        ptr = getattr(stream, "cuda_stream", stream)
        check(self.lib.gbx_sbm_set_stream(self._h, C.c_void_p(int(ptr) if ptr else 0)))

    def set_options(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.opt, k):
                raise TypeError(f"unknown option {k!r}")
            setattr(self.opt, k, v)
        check(self.lib.gbx_sbm_set_options(self._h, C.byref(self.opt)))

    def init(self, gr, Crs, psicav=None, seed: int = 0):
        gr = _f64(gr, (self.q,)) if gr is not None else None
        Crs = _f64(Crs, (self.q, self.q))
        psicav = _f64(psicav, (self.K, self.q))
        t0 = time.perf_counter()
        check(self.lib.gbx_sbm_init(self._h, _dp(gr), _dp(Crs), _dp(psicav), C.c_uint64(seed)))
        self.timings["init"] = time.perf_counter() - t0

    # -- running ----------------------------------------------------------------------------
    def em(self) -> SbmResult:
        res = SbmResult()
        t0 = time.perf_counter()
        check(self.lib.gbx_sbm_em(self._h, C.byref(res)))
        self.timings["em"] = time.perf_counter() - t0
        return res

    def iterate(self, n: int = 1, read_conv: bool = True):
        conv = C.c_double(float("nan"))
        check(self.lib.gbx_sbm_iterate(self._h, int(n), C.byref(conv) if read_conv else None))
        return conv.value

    def bp_sweeps(self, n: int = 1):
        check(self.lib.gbx_sbm_bp_sweeps(self._h, int(n)))

    def finalize(self) -> SbmResult:
        res = SbmResult()
        check(self.lib.gbx_sbm_finalize(self._h, C.byref(res)))
        return res

    # -- read back --------------------------------------------------------------------------
    def state(self, PSI=True, theta=True):
        q, n = self.q, self.n
        gr = np.empty(q)
        Crs = np.empty((q, q))
        h = np.empty(q)
        P = _host_array((n, q)) if PSI else None
        th = _host_array((n,)) if theta else None
        t0 = time.perf_counter()
        check(self.lib.gbx_sbm_get_state(self._h, _dp(gr), _dp(Crs), _dp(P), _dp(th), _dp(h)))
        self.timings["get_state"] = time.perf_counter() - t0
        return dict(gr=gr, Crs=Crs, PSI=P, theta=th, h=h)

    def messages(self):
        m = np.empty((self.K, self.q))
        check(self.lib.gbx_sbm_get_messages(self._h, _dp(m)))
        return m

    def assignment(self):
        b = np.empty(self.n, dtype=np.int32)
        check(self.lib.gbx_sbm_get_assignment(self._h, b.ctypes.data_as(C.POINTER(C.c_int32))))
        return b

    @property
    def launch_count(self):
        return int(self.lib.gbx_sbm_launch_count(self._h))

    @property
    def device_bytes(self):
        return int(self.lib.gbx_sbm_device_bytes(self._h))

    @property
    def sweep_bytes(self):
        return float(self.lib.gbx_sbm_sweep_bytes(self._h))

    @property
    def sweep_kind(self):
        """'pull' or 'push': which sweep kernels the current EM run uses (decided by init)."""
        return "pull" if self.lib.gbx_sbm_sweep_kind(self._h) else "push"


def initial_parameters(initial: str, B: int, rng: np.random.Generator):
    """Hyper-parameter initial conditions of sbm.jl:283-302 (the spectral one lives in graphbix_b200.spectral)."""
    if initial == "random":
        RandMat = 0.001 * rng.random((B, B)) + 0.0001
        RandMat[rng.integers(0, B), rng.integers(0, B)] = 10
        return np.ones(B) / B, RandMat + RandMat.T
    if initial == "uniformAssortative":
        return np.ones(B) / B, 20 * np.eye(B) + 0.01 * np.ones((B, B))
    if initial == "uniformDisassortative":
        return np.ones(B) / B, -9.9 * np.eye(B) + 10 * np.ones((B, B))
    if initial == "Bipartite":
        halfB = int(round(B / 2))
        layer2 = 20 * np.eye(halfB) + 0.01 * np.ones((halfB, halfB))
        return np.ones(B) / B, np.kron(np.array([[0.0, 1.0], [1.0, 0.0]]), layer2)
    raise ValueError(f"unknown initial condition {initial!r}")


def EM(Ntot, B, links, maxCrs, initial, degreecorrection, itrmax, priorlearning, learningrate, *,
       rng=None, PSIcav0=None, seed=0, BPconvthreshold=0.000001, damping=1.0, device=0, engine=None,
       return_engine=False, stream=None):
    """Drop-in for the reference's EM(), sbm.jl:28.  `initial` is a name or a `(gr, Crs)` pair.  `stream`: CUDA stream of
    the run (runs of DIFFERENT q on streams of their own overlap on one GPU, see parallel.threaded_sweep)."""
    rng = rng if rng is not None else np.random.default_rng(seed)
    if isinstance(initial, str):
        if initial == "normalizedLaplacian":
            from .spectral import laplacian_initial
            gr0, Crs0 = laplacian_initial(links, Ntot, B, rng)
        else:
            gr0, Crs0 = initial_parameters(initial, B, rng)
    else:
        gr0, Crs0 = initial
    eng = engine if engine is not None else SbmEngine(Ntot, links, B, device=device, stream=stream)
    eng.set_options(degreecorrection=int(bool(degreecorrection)), priorlearning=int(bool(priorlearning)),
                    learningrate=float(learningrate), maxCrs=float(maxCrs), itrmax=int(itrmax),
                    bp_conv_threshold=float(BPconvthreshold), damping=float(damping))
    eng.init(gr0, Crs0, PSIcav0, seed=int(rng.integers(0, 2 ** 63 - 1)))
    res = eng.em()
    st = eng.state(theta=False)
    out = (st["gr"].reshape(1, B), st["Crs"], st["PSI"], res.FE, res.CVBayes, res.CVGP, res.CVGT, res.CVMAP,
           res.varCVBayes, res.varCVGP, res.varCVGT, res.varCVMAP, bool(res.fail), bool(res.cnv), int(res.itrnum))
    if return_engine:
        return out, eng, res
    if engine is None:
        eng.close()
    return out


# =================================================================================================
# mod.jl: SBM restricted to community structure
# =================================================================================================
class ModEngine:
    """One graph + one EM state of mod.jl's model resident on one GPU (C ABI gbx_mod_*)."""

    def __init__(self, Ntot: int, links, B: int, device: int = 0, stream=None):
        self.lib = _lib.load()
        links = np.asarray(links)
        if links.ndim != 2 or links.shape[1] != 2:
            raise ValueError("links must be K x 2")
        self.n, self.q, self.K = int(Ntot), int(B), int(links.shape[0])
        colmajor = np.asfortranarray(links, dtype=np.int64)
        self._h = C.c_void_p()
        self.timings = {}
        t0 = time.perf_counter()
        check(self.lib.gbx_mod_create(C.byref(self._h), int(device), self.n, self.K,
                                      colmajor.ctypes.data_as(C.POINTER(C.c_int64)), self.q))
        self.timings["create"] = time.perf_counter() - t0
        self.opt = ModOptions()
        self.lib.gbx_mod_default_options(C.byref(self.opt))
        if stream is not None:
            ptr = getattr(stream, "cuda_stream", stream)
            check(self.lib.gbx_mod_set_stream(self._h, C.c_void_p(int(ptr) if ptr else 0)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.gbx_mod_destroy(self._h)
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
        check(self.lib.gbx_mod_set_options(self._h, C.byref(self.opt)))

    def init(self, gr, psicav=None, seed: int = 0):
        gr = _f64(gr, (self.q,))
        psicav = _f64(psicav, (self.K, self.q))
        t0 = time.perf_counter()
        check(self.lib.gbx_mod_init(self._h, _dp(gr), _dp(psicav), C.c_uint64(seed)))
        self.timings["init"] = time.perf_counter() - t0

    def em(self) -> ModResult:
        res = ModResult()
        t0 = time.perf_counter()
        check(self.lib.gbx_mod_em(self._h, C.byref(res)))
        self.timings["em"] = time.perf_counter() - t0
        return res

    def iterate(self, n: int = 1, read_conv: bool = True):
        conv = C.c_double(float("nan"))
        check(self.lib.gbx_mod_iterate(self._h, int(n), C.byref(conv) if read_conv else None))
        return conv.value

    def finalize(self) -> ModResult:
        res = ModResult()
        check(self.lib.gbx_mod_finalize(self._h, C.byref(res)))
        return res

    def state(self):
        q, n = self.q, self.n
        gr, h, P, par = np.empty(q), np.empty(q), _host_array((n, q)), np.empty(4)
        t0 = time.perf_counter()
        check(self.lib.gbx_mod_get_state(self._h, _dp(gr), _dp(P), _dp(h), _dp(par)))
        self.timings["get_state"] = time.perf_counter() - t0
        return dict(gr=gr, PSI=P, h=h, alpha=par[0], beta=par[1], omegain=par[2], omegaout=par[3])

    def messages(self):
        m = np.empty((self.K, self.q))
        check(self.lib.gbx_mod_get_messages(self._h, _dp(m)))
        return m

    def assignment(self):
        b = np.empty(self.n, dtype=np.int32)
        check(self.lib.gbx_mod_get_assignment(self._h, b.ctypes.data_as(C.POINTER(C.c_int32))))
        return b

    @property
    def launch_count(self):
        return int(self.lib.gbx_mod_launch_count(self._h))


def EM_mod(gr, Ntot, B, links, maxomega, itrmax, betafrac, learning, priorlearning, dc, learningrate, *,
           PSIcav0=None, seed=0, BPconvthreshold=0.000001, damping=1.0, device=0, return_engine=False, stream=None):
    """Drop-in for mod.jl's EM(), mod.jl:3.  Returns the reference's 17-tuple
    (excm,alpha,beta,omegain,omegaout,PSI,FE,CVBayes,CVGP,CVGT,CVMAP,varCVBayes,varCVGP,varCVGT,varCVMAP,cnv,itrnum)."""
    eng = ModEngine(Ntot, links, B, device=device, stream=stream)
    eng.set_options(dc=int(bool(dc)), learning=int(bool(learning)), priorlearning=int(bool(priorlearning)),
                    learningrate=float(learningrate), maxomega=float(maxomega), itrmax=int(itrmax),
                    betafrac=float(betafrac), bp_conv_threshold=float(BPconvthreshold), damping=float(damping))
    eng.init(gr, PSIcav0, seed=seed)
    r = eng.em()
    st = eng.state()
    out = (r.excm, r.alpha, r.beta, r.omegain, r.omegaout, st["PSI"], r.FE, r.CVBayes, r.CVGP, r.CVGT, r.CVMAP,
           r.varCVBayes, r.varCVGP, r.varCVGT, r.varCVMAP, bool(r.cnv), int(r.itrnum))
    if return_engine:
        return out, eng, r
    eng.close()
    return out
