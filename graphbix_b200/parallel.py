"""One process per GPU: the q sweep (and the `initnum` samples) of sbm.jl / mod.jl are independent EM runs, so they
are dealt to the ranks round-robin and gathered on rank 0 -- no data-path collective (BASELINE.json configs[1], [4]).

    torchrun --nproc-per-node 8 -m graphbix_b200.parallel  ...   (see `sweep`)

`torch.distributed` is plumbing only: NCCL on GPUs, gloo in the CPU tests."""
from __future__ import annotations

import os
from typing import Callable, Iterable, List, Tuple


def my_share(tasks: List, rank: int, world: int) -> List[Tuple[int, object]]:
    """Round-robin deal; returns (position, task) so that results can be put back in order."""
    return [(i, t) for i, t in enumerate(tasks) if i % world == rank]


def sweep(tasks: Iterable, run_task: Callable[[object, int], object], *, backend: str | None = None):
    """Run `run_task(task, local_device)` for this rank's share of `tasks`; rank 0 returns the results in task order
    (other ranks return None).  Works without torch.distributed being initialised (single process)."""
    tasks = list(tasks)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    mine = [(i, run_task(t, local)) for i, t in my_share(tasks, rank, world)]
    if world == 1:
        return [r for _, r in sorted(mine, key=lambda x: x[0])]
    import torch.distributed as dist
    created = False
    if not dist.is_initialized():
        dist.init_process_group(backend or ("nccl" if _has_cuda() else "gloo"))
        created = True
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    out = None
    if rank == 0:
        flat = sorted((x for part in gathered for x in part), key=lambda x: x[0])
        assert [i for i, _ in flat] == list(range(len(tasks))), "a task was lost or duplicated"
        out = [r for _, r in flat]
    if created:
        dist.destroy_process_group()
    return out


def threaded_sweep(tasks: Iterable, run_task: Callable[[object, object], object], *, device: int = 0, max_workers: int = 0):
    """The q values of a sweep that land on ONE GPU, side by side: a host thread and a CUDA stream per task (ctypes
    releases the GIL inside libgraphbix_cuda, so the launches of the runs interleave and their kernels overlap on the
    GPU; every q has its own parameter block in constant memory).  `run_task(task, stream)` must pass `stream` on to
    EM(..., stream=stream) / EM_mod(..., stream=stream).  Tasks of EQUAL q must not run concurrently on one device
    (include/graphbix_cuda.h): give each q one task, or loop over the samples of a q inside the task.  Returns the
    results in task order."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    tasks = list(tasks)
    if not tasks:
        return []
    workers = max_workers or len(tasks)

    def one(t):
        torch.cuda.set_device(device)
        st = torch.cuda.Stream(device=device)
        out = run_task(t, st)
        st.synchronize()
        return out

    with ThreadPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(one, tasks))


def _has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def sbm_q_sweep(Ntot, links, q_values, *, initial="uniformAssortative", degreecorrection=True, itrmax=256,
                priorlearning=True, learningrate=0.3, seed=0):
    """The q loop of sbm.jl:892-996 with one q per GPU: returns, on rank 0, a list of dicts (q, FE, CV errors, ...)."""
    import numpy as np

    def one(q, device):
        from .engine import EM
        rng = np.random.default_rng(seed + q)
        out = EM(Ntot, q, links, Ntot, initial, degreecorrection, itrmax, priorlearning, learningrate, rng=rng,
                 device=device)
        return dict(q=q, gr=out[0], Crs=out[1], block=np.argmax(out[2], axis=1) + 1, FE=out[3], CVBayes=out[4],
                    CVGP=out[5], CVGT=out[6], CVMAP=out[7], cnv=out[13], itrnum=out[14])

    return sweep(list(q_values), one)
