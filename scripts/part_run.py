"""Partitioned leg of bench.py alone (torchrun): python -m torch.distributed.run ... scripts/part_run.py [bench args]."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench

a = bench.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29542")
dist.init_process_group("nccl", device_id=torch.device("cuda", local), rank=rank, world_size=world)
out = bench.run_partitioned(a, torch, dist, world, rank, local)
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
