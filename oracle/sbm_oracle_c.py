"""ctypes front end of oracle/sbm_oracle.c (CPU ORACLE, test infrastructure -- see the header of the C file).

Same signature and result object as oracle/sbm_oracle.py:EM, plus `time_em_iterations` for bench.py's CPU
baseline.  PARITY UNPINNED against the Julia reference (see sbm_oracle.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .sbm_oracle import EMResult

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsbm_oracle.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "sbm_oracle.c")):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(_SO)
        _lib.sbm_oracle_em.restype = C.c_int
    return _lib


def available() -> bool:
    try:
        _load()
        return True
    except Exception:
        return False


def EM(Ntot, B, links, maxCrs, initial, degreecorrection, itrmax, priorlearning, learningrate, *, PSIcav0,
       BPconvthreshold=0.000001, float32_logbiprob=True, seed=1, perms=None) -> EMResult:
    lib = _load()
    links = np.ascontiguousarray(links, dtype=np.int64)
    K = links.shape[0]
    gr0 = np.ascontiguousarray(initial[0], dtype=np.float64).ravel()
    C0 = np.ascontiguousarray(initial[1], dtype=np.float64)
    psi0 = np.ascontiguousarray(PSIcav0, dtype=np.float64).reshape(K, B)
    gr, Crs, PSI = np.empty(B), np.empty((B, B)), np.empty((Ntot, B))
    PSIcav, theta = np.empty((K, B)), np.empty(Ntot)
    scal, iout, dout = np.zeros(9), np.zeros(4, dtype=np.int32), np.zeros(2)
    pp = None
    if perms is not None:
        pp = np.ascontiguousarray(np.asarray(perms, dtype=np.int64)[:itrmax])
        assert pp.shape == (min(itrmax, len(perms)), Ntot) and pp.shape[0] == itrmax
    p = lambda a, t=C.c_double: a.ctypes.data_as(C.POINTER(t))
    rc = lib.sbm_oracle_em(C.c_int64(Ntot), C.c_int64(K), p(links, C.c_int64), C.c_int(B), C.c_double(maxCrs), p(gr0),
                           p(C0), p(psi0), C.c_int(int(degreecorrection)), C.c_int(int(itrmax)),
                           C.c_int(int(priorlearning)), C.c_double(learningrate), C.c_double(BPconvthreshold),
                           C.c_int(int(float32_logbiprob)), C.c_uint64(seed),
                           p(pp, C.c_int64) if pp is not None else None, p(gr), p(Crs), p(PSI), p(PSIcav), p(theta),
                           p(scal), p(iout, C.c_int32), p(dout))
    if rc != 0:
        raise RuntimeError(f"sbm_oracle_em failed with code {rc}")
    r = EMResult(gr, Crs, PSI, *[float(x) for x in scal], bool(iout[0]), bool(iout[1]), int(iout[2]), PSIcav=PSIcav,
                 theta=theta, h=None, BPconv=float(dout[0]), itrs_done=int(iout[3]))
    r.loop_seconds = float(dout[1])
    return r


def time_em_iterations(Ntot, B, links, psi0, gr0, C0, dc, prior, lr, warmup, steps):
    """Seconds per EM iteration (sweep + update_Crs + update_theta) of the C oracle, single thread."""
    r = EM(Ntot, B, links, Ntot, (gr0, C0), dc, warmup + steps, prior, lr, PSIcav0=psi0, BPconvthreshold=0.0)
    return r.loop_seconds / max(r.itrs_done, 1), 1
