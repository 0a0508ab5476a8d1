"""All-to-all over NVLink/NVSwitch with NCCL (torchrun, one rank per GPU): what the fabric gives a boundary exchange
of `mb` MB per rank, in one call and in `pieces` calls."""
import os
import sys

import torch
import torch.distributed as dist

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
mb = float(sys.argv[1]) if len(sys.argv) > 1 else 896.0
n = int(mb * 1e6 / 8) // (world * 64) * (world * 64)
send = torch.zeros(n, dtype=torch.float64, device="cuda")
recv = torch.empty_like(send)
for pieces in (1, 8):
    m = n // pieces
    for rep in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for p in range(pieces):
            dist.all_to_all_single(recv[p * m:(p + 1) * m], send[p * m:(p + 1) * m])
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        sent = n * 8 * (world - 1) / world
        print(f"NCCL all_to_all_single world={world} {n * 8 / 1e6:.0f} MB per rank ({sent / 1e6:.0f} MB leave the GPU), "
              f"{pieces} call(s): {t.item():.3f} ms  {sent / t.item() * 1e-6:.0f} GB/s egress per GPU")
dist.destroy_process_group()
