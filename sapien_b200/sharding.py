"""Environment sharding of batched stereo sensors across the GPUs of one box.

The reference has no batch dimension and no multi-GPU support (one ``DepthSensorEngine`` = one
stereo pair on the current device).  Many-environment simulators (ManiSkill-style)
own N independent sensors, and every (environment, frame) pair is an independent instance of the
pipeline, so the path shards with NO data-path collective: rank r owns the contiguous environment
block ``env_range(N, r, world)`` and runs one batched engine on its own GPU.

The only exchange is the OPTIONAL epilogue for a single host-side consumer: gathering the per-rank
depth maps ``[n_r, H, W]`` (float32) with ``torch.distributed`` -- NCCL over NVLink/NVSwitch for
CUDA tensors, gloo for the CPU tests.  It is outside the frames/s metric.
"""
from __future__ import annotations

from typing import List, Optional, Tuple


def env_range(n_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced block [start, stop) of environments owned by `rank`.
    The first ``n_envs % world`` ranks own one environment more."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world: {rank}/{world}")
    if n_envs < 0:
        raise ValueError("n_envs must be >= 0")
    base, extra = divmod(n_envs, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def env_counts(n_envs: int, world: int) -> List[int]:
    return [b - a for a, b in (env_range(n_envs, r, world) for r in range(world))]


def owner_of(env: int, n_envs: int, world: int) -> int:
    """Rank that owns environment `env`."""
    if not 0 <= env < n_envs:
        raise ValueError("env out of range")
    base, extra = divmod(n_envs, world)
    cut = extra * (base + 1)
    return env // (base + 1) if env < cut else extra + (env - cut) // max(base, 1)


def gather_envs(local, n_envs: int, dst: Optional[int] = None, group=None):
    """Gathers per-rank results ``[n_r, ...]`` into ``[n_envs, ...]`` in environment order.

    ``dst=None``: every rank receives the full tensor (all-gather); otherwise only rank `dst`
    does and the others get ``None``.  Works for uneven shards (padded to the largest block).
    """
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if local.shape[0] != n_envs:
            raise ValueError("single-process gather needs all environments locally")
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = env_counts(n_envs, world)
    if local.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} environments, expected {counts[rank]}")
    cmax = max(counts)
    tail = tuple(local.shape[1:])
    if counts[rank] == cmax:
        padded = local.contiguous()
    else:
        padded = torch.zeros((cmax,) + tail, dtype=local.dtype, device=local.device)
        padded[: counts[rank]] = local
    if dst is None:
        full = torch.empty((world * cmax,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(full, padded, group=group)
    else:
        parts = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
        dist.gather(padded, parts, dst=dst, group=group)
        if rank != dst:
            return None
        full = torch.cat(parts, dim=0)
    if all(c == cmax for c in counts):
        return full
    full = full.view((world, cmax) + tail)
    return torch.cat([full[r, : counts[r]] for r in range(world)], dim=0)


class ShardedStereoDepth:
    """One batched engine per process (= per GPU) over this rank's block of environments.

    ``engine_args`` are the 40 positional ``DepthSensorEngine`` arguments (all environments share
    one sensor model).  ``compute`` takes this rank's ``[n_r, H, W, 4]`` float32 RGBA CUDA tensors
    (or ``[n_r, H, W]`` uint8 arrays) and returns the local depth as a torch CUDA tensor view.
    """

    def __init__(self, engine_args, n_envs: int, rank: int = 0, world: int = 1, device: Optional[int] = None,
                 engine_factory=None):
        self.n_envs, self.rank, self.world = n_envs, rank, world
        self.start, self.stop = env_range(n_envs, rank, world)
        self.local = self.stop - self.start
        self.engine = None
        if self.local > 0:
            if engine_factory is None:
                from .simsense import DepthSensorEngine  # fails loudly without the CUDA extension

                engine_factory = DepthSensorEngine
            self.engine = engine_factory(*engine_args, device=-1 if device is None else device, batch=self.local)

    def compute(self, left, right, *bbox, **kw):
        if self.engine is None:
            return None
        self.engine.compute(left, right, *bbox, **kw)
        return self.engine.get_cuda()

    def gather_depth(self, local_depth, dst: Optional[int] = None, group=None):
        return gather_envs(local_depth, self.n_envs, dst=dst, group=group)
