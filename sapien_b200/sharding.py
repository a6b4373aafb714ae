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
    if cmax == 0:
        return local
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
    """This rank's block of environments on its GPU (one process per GPU).

    ``engine_args`` are the 40 positional ``DepthSensorEngine`` arguments (all environments share
    one sensor model).  ``compute`` takes this rank's ``[n_r, H, W, 4]`` float32 RGBA CUDA tensors
    (or ``[n_r, H, W]`` uint8 arrays) -- the leading dimension is always present, also for
    ``n_r == 1`` -- and returns the local depth ``[n_r, out_H, out_W]`` as a torch CUDA tensor.

    ``pipelines=K`` splits the local block once more into K contiguous sub-blocks with one engine
    each.  The K engines own their streams, so the HBM-bound aggregation passes of one sub-block
    overlap the issue-bound kernels (cost volume, final pass, post-processing) of another; the
    result is the same tensor (sub-blocks are gathered into one output buffer on the device).
    A rank that owns no environment (``n_envs < world``) has no engine and contributes an empty
    tensor to the gather.
    """

    def __init__(self, engine_args, n_envs: int, rank: int = 0, world: int = 1, device: Optional[int] = None,
                 engine_factory=None, pipelines: int = 1):
        self.n_envs, self.rank, self.world = n_envs, rank, world
        self.start, self.stop = env_range(n_envs, rank, world)
        self.local = self.stop - self.start
        self.engines, self.blocks = [], []
        self._out = None
        if self.local > 0:
            if engine_factory is None:
                from .simsense import DepthSensorEngine  # fails loudly without the CUDA extension

                engine_factory = lambda *a, **k: DepthSensorEngine(*a, batched=True, **k)  # noqa: E731
            k = max(1, min(int(pipelines), self.local))
            for a, b in (env_range(self.local, i, k) for i in range(k)):
                self.blocks.append((a, b))
                self.engines.append(engine_factory(*engine_args, device=-1 if device is None else device, batch=b - a))

    @property
    def engine(self):
        """The engine of a single-pipeline shard (None on a rank without environments)."""
        return self.engines[0] if self.engines else None

    def enqueue(self, left, right, *bbox, inputs_ready: bool = False, **kw):
        """Enqueues the local block on every pipeline without waiting (device inputs only).
        ``inputs_ready=True`` states that the inputs are complete: no ordering against the caller's stream
        is established and the frames simply queue behind each engine's previous frame."""
        kw.setdefault("sync", False)
        for eng, (a, b) in zip(self.engines, self.blocks):
            if inputs_ready:
                kw["stream"] = eng.cuda_stream
            eng.compute(left[a:b], right[a:b], *bbox, **kw)

    def synchronize(self):
        for eng in self.engines:
            eng.synchronize()

    def depth(self):
        """Local depth ``[n_r, out_H, out_W]`` (torch tensor; aliases engine memory for one pipeline)."""
        import torch

        if not self.engines:
            return None
        views = [self._as_tensor(eng.get_cuda()) for eng in self.engines]
        if len(views) == 1:
            return views[0]
        if self._out is None:
            self._out = torch.empty((self.local,) + tuple(views[0].shape[1:]), dtype=views[0].dtype, device=views[0].device)
        for v, (a, b) in zip(views, self.blocks):
            self._out[a:b].copy_(v)
        return self._out

    @staticmethod
    def _as_tensor(x):
        return x.torch() if hasattr(x, "torch") and not hasattr(x, "dtype") else x

    def compute(self, left, right, *bbox, **kw):
        if not self.engines:
            return None
        if len(self.engines) == 1:
            self.engines[0].compute(left, right, *bbox, **kw)
        else:
            sync = kw.pop("sync", True)
            self.enqueue(left, right, *bbox, **kw)
            if sync:
                self.synchronize()
        return self.depth()

    def gather_depth(self, local_depth, dst: Optional[int] = None, group=None, like=None):
        """All-gather (``dst=None``) or gather of the local depth maps in environment order.  A rank without
        environments passes ``None`` and contributes ``[0, ...]`` shaped like ``like`` (any rank's depth
        shape/dtype/device: ``like=(tail_shape, dtype, device)``)."""
        if local_depth is None:
            import torch

            if like is None:
                raise ValueError("a rank without environments needs like=(tail_shape, dtype, device) to join the gather")
            tail, dtype, device = like
            local_depth = torch.zeros((0,) + tuple(tail), dtype=dtype, device=device)
        return gather_envs(local_depth, self.n_envs, dst=dst, group=group)
