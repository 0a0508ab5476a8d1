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
