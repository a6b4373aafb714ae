#!/usr/bin/env python
"""Times the optional NCCL all-gather of depth maps by itself (torchrun, N ranks): preallocated output vs gather_envs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from sapien_b200 import sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1024
c0, c1 = sharding.env_range(n, rank, world)
loc = torch.full((c1 - c0, 256, 256), float(rank), dtype=torch.float32, device="cuda")
full = torch.empty((n, 256, 256), dtype=torch.float32, device="cuda")


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


a = timed(lambda: dist.all_gather_into_tensor(full, loc))
b = timed(lambda: sharding.gather_envs(loc, n))
if rank == 0:
    nb = loc.numel() * 4
    print(f"world {world}: {nb / 1e6:.1f} MB per rank; all_gather_into_tensor preallocated {a:.3f} ms ({nb * world / a / 1e6:.0f} GB/s algbw), gather_envs {b:.3f} ms")
dist.destroy_process_group()
