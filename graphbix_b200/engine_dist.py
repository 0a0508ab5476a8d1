"""Vertex-partitioned EM + BP across the GPUs of one NVLink domain (BASELINE.json configs[3]): one process per GPU.

Rank r owns a contiguous range of vertices and the rows of their incoming messages.  During a sweep each rank
packs the new messages for rows of other ranks into per-destination runs that the copy engine pushes into the
owner's receive buffer over NVLink while the sweep continues (peer pointers, exchanged once here through CUDA IPC
handles), and the owner scatters them into its rows after the all-reduce (`set_exchange(0)`: 64-B peer stores straight
from the sweep kernel instead); the only per-iteration collectives are the all-reduce of the totals
vector (residual, sum PSI, block degree sums, M-step statistics: 4 + 4q + q^2 doubles) and, with degree correction,
the all-reduce of S_theta (q doubles) and the all-gather of theta (N doubles).  torch.distributed (NCCL) is the
plumbing for those; the kernels are the same as on one GPU (C ABI: gbx_sbm_create_part, gbx_sbm_step, ...).

Launch with torchrun; every rank passes the same `links`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import (MODE_LOGZ, MODE_MARGINAL, MODE_SWEEP, STEP_APPLY_TOTALS, STEP_COUNT_ITERATION, STEP_ESCALE,
                   STEP_FE_APPLY, STEP_FE_FINAL, STEP_FE_LOCAL, STEP_LOCAL_TOTALS, STEP_PREPARE, STEP_THETA_APPLY,
                   STEP_THETA_LOCAL, STEP_UNPACK, STEP_VERTEX_PASS, ModOptions, ModResult, SbmOptions, SbmResult, check)
from .engine import _dp, _f64


class PartEngine:
    """One rank's share of a partitioned sbm.jl (`model="sbm"`) or mod.jl (`model="mod"`) run."""

    def __init__(self, Ntot: int, links, B: int, model: str = "sbm"):
        import torch
        import torch.distributed as dist
        assert dist.is_initialized(), "initialise torch.distributed (backend nccl) first"
        self.torch, self.dist = torch, dist
        self.lib = _lib.load()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.local = int(os.environ.get("LOCAL_RANK", self.rank))
        torch.cuda.set_device(self.local)
        self.model = model
        self._debug = bool(int(os.environ.get("GBX_PART_SYNC", "0")))   # synchronise after every step (debugging)
        self._profile = [] if int(os.environ.get("GBX_PART_PROFILE", "0")) else None
        if torch.is_tensor(links):
            # edge list already on this GPU: a (2, K) int64 tensor = the K x 2 `links` matrix in column-major order
            assert links.is_cuda and links.dtype == torch.int64 and links.dim() == 2 and links.shape[0] == 2
            links = links.contiguous()
            self.K_global = int(links.shape[1])
            ptr = C.cast(C.c_void_p(links.data_ptr()), C.POINTER(C.c_int64))
            torch.cuda.synchronize()
        else:
            links = np.asarray(links)
            self.K_global = int(links.shape[0])
            colmajor = np.asfortranarray(links, dtype=np.int64)
            ptr = colmajor.ctypes.data_as(C.POINTER(C.c_int64))
        self.n_global, self.q = int(Ntot), int(B)
        create = self.lib.gbx_sbm_create_part if model == "sbm" else self.lib.gbx_mod_create_part
        self._h = C.c_void_p()
        check(create(C.byref(self._h), self.local, self.n_global, self.K_global, ptr, self.q, self.rank, self.world))
        info = (C.c_int64 * 8)()
        check(self.lib.gbx_sbm_part_info(self._h, info))
        self.vbeg, self.n, self.sbeg, self.K, self.tot_len, self.chunk = (int(info[i]) for i in range(6))
        dev = torch.device("cuda", self.local)
        self.totals = torch.zeros(self.tot_len, dtype=torch.float64, device=dev)
        self.theta = torch.ones(self.world * self.chunk, dtype=torch.float64, device=dev)
        check(self.lib.gbx_sbm_part_set_buffers(self._h, C.c_void_p(self.totals.data_ptr()),
                                                C.c_void_p(self.theta.data_ptr())))
        check(self.lib.gbx_sbm_set_stream(self._h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        # peer pointers: both message buffers (push sweeps store into the owner's rows), both receive buffers (packed
        # exchange of the push sweep, when it was planned) and both per-vertex record buffers (pull sweeps store a
        # vertex's record into every rank's copy) of every rank
        for which in ((0, 1, 2, 3, 4, 5) if self.world > 1 else ()):
            buf = (C.c_ubyte * 64)()
            rc = self.lib.gbx_sbm_ipc_export(self._h, which, buf)
            have = [None] * self.world
            dist.all_gather_object(have, int(rc))
            if any(r != 0 for r in have):       # no such buffer on this kind of handle (the same on every rank)
                continue
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(buf))
            for r, hb in enumerate(handles):
                if r != self.rank:
                    arr = (C.c_ubyte * 64).from_buffer_copy(hb)
                    check(self.lib.gbx_sbm_ipc_import(self._h, which, r, arr))
        if model == "sbm":
            self.opt = SbmOptions()
            self.lib.gbx_sbm_default_options(C.byref(self.opt))
        else:
            self.opt = ModOptions()
            self.lib.gbx_mod_default_options(C.byref(self.opt))
        dist.barrier()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.torch.cuda.synchronize()
            self.dist.barrier()          # nobody may still be storing into this rank's buffers
            self.lib.gbx_sbm_destroy(self._h)
            self._h = None

    # ------------------------------------------------------------------------------------------
    def set_options(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.opt, k):
                raise TypeError(f"unknown option {k!r}")
            setattr(self.opt, k, v)
        fn = self.lib.gbx_sbm_set_options if self.model == "sbm" else self.lib.gbx_mod_set_options
        check(fn(self._h, C.byref(self.opt)))

    def _tick(self, label):
        """GBX_PART_PROFILE=1: CUDA events between the steps of an iteration (profile_summary() adds them up)."""
        if self._profile is not None:
            ev = self.torch.cuda.Event(enable_timing=True)
            ev.record()
            self._profile.append((label, ev))

    def profile_summary(self):
        if not self._profile:
            return {}
        self.torch.cuda.synchronize()
        out = {}
        for (_, e0), (label, e1) in zip(self._profile[:-1], self._profile[1:]):
            out[label] = out.get(label, 0.0) + e0.elapsed_time(e1)
        self._profile = []
        return out

    def _sync(self, what):
        self._tick(what)
        if self._debug:
            try:
                self.torch.cuda.synchronize()
            except Exception as e:
                raise RuntimeError(f"rank {self.rank}: device error after {what}: {e}") from e

    def _step(self, step, arg=0):
        check(self.lib.gbx_sbm_step(self._h, step, arg))
        self._sync(f"gbx_sbm_step({step}, {arg})")

    def _allreduce(self, count=None):
        t = self.totals if count is None else self.totals[:count]
        self.dist.all_reduce(t)
        self._sync("all_reduce")

    def set_exchange(self, mode: int):
        """1 (default): boundary messages packed, pushed by the copy engine, unpacked after the all-reduce;
        0: stored one by one into the owner's buffer from the sweep kernel."""
        check(self.lib.gbx_sbm_part_set_exchange(self._h, int(mode)))

    @property
    def pull(self):
        """True when the run sweeps with the pull kernels (decided by init): what crosses NVLink is then one record per
        vertex and rank, stored by the sweep kernel itself; no unpack, no theta all-gather, no per-slot theta scales."""
        return bool(self.lib.gbx_sbm_sweep_kind(self._h))

    def _pass(self, mode):
        self._step(STEP_VERTEX_PASS, mode)
        self._step(STEP_LOCAL_TOTALS, mode)
        self._allreduce()                   # also the barrier after which every rank's peer stores have landed
        if mode == MODE_SWEEP and not self.pull:
            self._step(STEP_UNPACK)
        self._step(STEP_APPLY_TOTALS, mode)

    @property
    def _dc(self):
        return bool(self.opt.degreecorrection) if self.model == "sbm" else False

    def init(self, gr, Crs=None, psicav=None, seed: int = 0):
        gr = _f64(gr, (self.q,))
        psicav = _f64(psicav, (self.K_global, self.q))   # None: drawn on the device from `seed` (same on every rank)
        self.theta.fill_(1.0)
        if self.model == "sbm":
            Crs = _f64(Crs, (self.q, self.q))
            check(self.lib.gbx_sbm_init(self._h, _dp(gr), _dp(Crs), _dp(psicav), C.c_uint64(seed)))
        else:
            check(self.lib.gbx_mod_init(self._h, _dp(gr), _dp(psicav), C.c_uint64(seed)))
        self._step(STEP_PREPARE, 1)         # h = 0
        self._pass(MODE_MARGINAL)           # PSI = updatePSI(); gr = update_gr(PSI)

    def iterate(self, n: int = 1):
        q, c, r = self.q, self.chunk, self.rank
        for _ in range(n):
            self._step(STEP_PREPARE, 0)
            self._pass(MODE_SWEEP)          # the all-reduce inside also orders the peer stores of this sweep
            if self._dc:
                self._step(STEP_THETA_LOCAL)
                self._allreduce(q)
                self._step(STEP_THETA_APPLY)
                if not self.pull:
                    # in place: this rank's slice of the output is its input (NCCL's in-place all-gather layout)
                    self.dist.all_gather_into_tensor(self.theta, self.theta[r * c:(r + 1) * c])
                    self._sync("all_gather(theta)")
                    self._step(STEP_ESCALE)
            self._step(STEP_COUNT_ITERATION)
            self._tick("iteration end")

    def residual(self) -> float:
        v = C.c_double()
        check(self.lib.gbx_sbm_read_residual(self._h, C.byref(v)))
        return v.value

    def em(self, itrmax=None, tol=None):
        itrmax = self.opt.itrmax if itrmax is None else itrmax
        tol = self.opt.bp_conv_threshold if tol is None else tol
        cnv, itrnum = False, 0
        for itr in range(1, itrmax + 1):
            self.iterate(1)
            if self.residual() < tol:       # identical on every rank: it comes from the all-reduced totals
                cnv, itrnum = True, itr
                break
        res = self.finalize()
        res.cnv, res.itrnum = int(cnv), itrnum
        return res

    def finalize(self):
        self._step(STEP_PREPARE, 2 if self.model == "sbm" else 0)   # h = update_h(); gr = update_gr(PSI) (sbm.jl:364-365)
        self._pass(MODE_LOGZ)
        for p in (0, 1):
            self._step(STEP_FE_LOCAL, p)
            self._allreduce(5)
            self._step(STEP_FE_APPLY, p)
        self._step(STEP_FE_FINAL)
        if self.model == "sbm":
            res = SbmResult()
            check(self.lib.gbx_sbm_finalize(self._h, C.byref(res)))
        else:
            res = ModResult()
            check(self.lib.gbx_mod_finalize(self._h, C.byref(res)))
        return res

    # ------------------------------------------------------------------------------------------
    def state(self):
        """gr, Crs, h (replicated) and this rank's rows of PSI (vertices vbeg .. vbeg + n)."""
        q = self.q
        gr, Crs, h, P = np.empty(q), np.empty((q, q)), np.empty(q), np.empty((self.n, q))
        if self.model == "sbm":
            check(self.lib.gbx_sbm_get_state(self._h, _dp(gr), _dp(Crs), _dp(P), None, _dp(h)))
            return dict(gr=gr, Crs=Crs, PSI=P, h=h, vbeg=self.vbeg)
        par = np.empty(4)
        check(self.lib.gbx_mod_get_state(self._h, _dp(gr), _dp(P), _dp(h), _dp(par)))
        return dict(gr=gr, PSI=P, h=h, alpha=par[0], beta=par[1], omegain=par[2], omegaout=par[3], vbeg=self.vbeg)

    def messages_all(self):
        """All K_global messages in `links` order on every rank (each rank contributes the rows it owns)."""
        m = np.empty((self.K_global, self.q))
        check(self.lib.gbx_sbm_get_messages(self._h, _dp(m)))
        t = self.torch.from_numpy(m).cuda()
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    def marginals_all(self):
        """PSI of all N vertices on every rank."""
        torch = self.torch
        pad = torch.zeros((self.chunk, self.q), dtype=torch.float64, device="cuda")
        pad[: self.n] = torch.from_numpy(self.state()["PSI"]).cuda()
        out = torch.empty((self.world * self.chunk, self.q), dtype=torch.float64, device="cuda")
        self.dist.all_gather_into_tensor(out, pad)
        return out[: self.n_global].cpu().numpy()
