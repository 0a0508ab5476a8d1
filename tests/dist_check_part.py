"""Run under torchrun on >= 2 GPUs (scripts/gpu_cycle.sh part): the vertex-partitioned engine must reproduce the
single-GPU engine sweep for sweep (same kernels, same schedule; only the order of the totals' summation differs)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from helpers import assortative_init, planted_case  # noqa: E402
from graphbix_b200.engine import ModEngine, SbmEngine  # noqa: E402
from graphbix_b200.engine_dist import PartEngine  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    # (vertices, q, degree correction, exchange mode, hubs): 120k vertices -> the packed sweep runs in several pieces
    cases = ((4000, 3, True, 1, False), (4000, 8, True, 1, False), (4000, 8, False, 0, False), (4000, 8, True, 0, False),
             (3000, 8, True, 1, True), (3000, 5, True, 1, True), (4000, 5, False, 0, False),
             (120000, 8, True, 1, False), (120000, 2, False, 1, False))
    if os.environ.get("GBX_PART_CASES"):
        import ast
        cases = ast.literal_eval(os.environ["GBX_PART_CASES"])
    for nv, q, dc, exchange, hubs in cases:
        if hubs:
            from test_gpu_sweep import _with_hubs
            n, links, psi0 = _with_hubs(nv, q, seed=5)
        else:
            n, links, psi0, _ = planted_case(nv, q, 10.0, 0.1, seed=90 + q)
        gr0, C0 = assortative_init(q)
        ref = SbmEngine(n, links, q, device=local)
        ref.set_options(degreecorrection=int(dc), priorlearning=1, learningrate=0.3)
        ref.init(gr0, C0, psi0)
        part = PartEngine(n, links, q, "sbm")
        part.set_exchange(exchange)
        part.set_options(degreecorrection=int(dc), priorlearning=1, learningrate=0.3)
        part.init(gr0, C0, psi0)
        for it in range(6 if nv < 100000 else 4):
            ref.iterate(1)
            part.iterate(1)
            sr, sp = ref.state(), part.state()
            dP = np.abs(part.marginals_all() - sr["PSI"]).max()
            dM = np.abs(part.messages_all() - ref.messages()).max()
            dC = np.abs(sp["Crs"] / sr["Crs"] - 1).max()
            dg = np.abs(sp["gr"] / sr["gr"] - 1).max()
            good = dP < 1e-11 and dM < 1e-11 and dC < 1e-10 and dg < 1e-10
            ok &= good
            if rank == 0:
                print(f"sbm n={nv} q={q} dc={dc} exchange={exchange} hubs={hubs} it={it}: dPSI={dP:.2e} dmsg={dM:.2e} dCrs={dC:.2e} dgr={dg:.2e} {'ok' if good else 'FAIL'}")
        rr, rp = ref.finalize(), part.finalize()
        for k in ("FE", "CVBayes", "CVGP", "CVGT", "CVMAP", "varCVBayes", "varCVGP", "varCVGT", "varCVMAP"):
            a, b = getattr(rr, k), getattr(rp, k)
            good = abs(a / b - 1) < 1e-9
            ok &= good
            if rank == 0 and not good:
                print(f"   {k}: single {a} partitioned {b} FAIL")
        if rank == 0:
            print(f"   free energy / CV errors {'ok' if ok else 'FAIL'} (FE {rr.FE:.12f} vs {rp.FE:.12f})")
        part.close()
        ref.close()
    # mod.jl model
    q = 3
    n, links, psi0, _ = planted_case(3000, q, 9.0, 0.1, seed=77)
    gr = np.ones(q) / q
    for exchange in (1, 0):
        ref = ModEngine(n, links, q, device=local)
        ref.init(gr, psi0)
        part = PartEngine(n, links, q, "mod")
        part.set_exchange(exchange)
        part.init(gr, None, psi0)
        for it in range(5):
            ref.iterate(1)
            part.iterate(1)
            sr, sp = ref.state(), part.state()
            dP = np.abs(part.marginals_all() - sr["PSI"]).max()
            do = abs(sp["omegain"] / sr["omegain"] - 1) + abs(sp["omegaout"] / sr["omegaout"] - 1)
            good = dP < 1e-11 and do < 1e-10
            ok &= good
            if rank == 0:
                print(f"mod q={q} exchange={exchange} it={it}: dPSI={dP:.2e} domega={do:.2e} {'ok' if good else 'FAIL'}")
        rr, rp = ref.finalize(), part.finalize()
        good = abs(rr.FE / rp.FE - 1) < 1e-9 and abs(rr.CVBayes / rp.CVBayes - 1) < 1e-9
        ok &= good
        if rank == 0:
            print(f"   mod FE {rr.FE:.12f} vs {rp.FE:.12f} {'ok' if good else 'FAIL'}")
        part.close()
        ref.close()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if rank == 0:
        print("PARTITIONED CHECK", "PASSED" if int(flag) == 0 else "FAILED")
    sys.exit(0 if int(flag) == 0 else 1)


if __name__ == "__main__":
    main()
