"""Multi-GPU sharding of the hot path: one process per GPU, no collective on the data path.

Rays are independent, so a depth map is split into contiguous row blocks of the
ray grid (screen-space locality for L2) and a multi-image sweep into whole images, round-robin.  The only
communication is one gather of the rendered ``[rays, 4]`` (depth_z, r, g, b) block to rank 0 per depth map,
through ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests).
The reference itself has no multi-GPU support (main.py:108,206-215: ``devices=[0]``).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_rows(H: int, W: int, world: int, rank: int) -> Tuple[int, int]:
    """(ray_begin, n_rays) of this rank's block of whole image rows; blocks differ by at most one row."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    base, extra = divmod(H, world)
    r0 = rank * base + min(rank, extra)
    rows = base + (1 if rank < extra else 0)
    return r0 * W, rows * W


def shard_counts(H: int, W: int, world: int) -> List[int]:
    return [shard_rows(H, W, world, r)[1] for r in range(world)]


def shard_images(n_images: int, world: int, rank: int) -> List[int]:
    """Round-robin image assignment for sweeps (BASELINE config 5: 49 depth maps over 8 GPUs)."""
    return list(range(rank, n_images, world))


def gather_depth_rgb(depth_z: torch.Tensor, rgb: torch.Tensor, counts: Sequence[int], dst: int = 0,
                     group=None) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
    """Gather per-rank ``depth_z [n_r]`` and ``rgb [n_r, 3]`` to ``dst``; returns the concatenation there.

    Shards are padded to the largest count so that one fixed-size collective suffices.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return depth_z, rgb
    rank = dist.get_rank(group)
    n_max = max(counts)
    n = depth_z.shape[0]
    if n != counts[rank]:
        raise ValueError(f"rank {rank}: have {n} rays, plan says {counts[rank]}")
    packed = torch.zeros(n_max, 4, dtype=torch.float32, device=depth_z.device)
    packed[:n, 0] = depth_z
    packed[:n, 1:] = rgb
    if dist.get_backend(group) == "nccl" or rank == dst:
        bufs = [torch.empty_like(packed) for _ in range(world)] if rank == dst else None
    else:
        bufs = None
    dist.gather(packed, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    full = torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)
    return full[:, 0].contiguous(), full[:, 1:].contiguous()
