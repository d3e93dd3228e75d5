"""Multi-GPU plumbing. Games are independent units (``src/cpp/selfplaymanager.cpp:133-149``): rank r owns
a disjoint range of game slots with its own trees, RNG streams and replica of the tower, and the hot path
has NO collective. The only exchange is once per generation: the finished sample records are gathered to
rank 0 (what the reference does over its Twisted TCP link, ``src/ggpzero/distributed/worker.py:162-206``,
``server.py:237-278``) — here one padded all-gather of fixed-size int32 records over NCCL (NVLink/NVSwitch),
or gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(rank: int, world: int, total_games: int):
    """[lo, hi) of the game slots owned by ``rank`` (contiguous, sizes differ by at most one)."""
    base, extra = divmod(total_games, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_seed(seed: int, rank: int) -> int:
    """Distinct, reproducible engine seed per rank (global game id = rank's lo + local slot)."""
    return (seed + 7919 * rank) & 0xFFFFFFFFFFFF


def gather_sample_records(records: torch.Tensor, dst: int = 0, group=None):
    """records: int32 [n_local, srec] on this rank's device (n_local may differ per rank).
    Returns on ``dst`` the concatenation in rank order (int32 [sum n, srec]); None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return records
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = records.device
    srec = records.shape[1]
    count = torch.tensor([records.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts)
    if m == 0:
        return records[:0] if rank == dst else None
    padded = torch.zeros((m, srec), dtype=records.dtype, device=dev)
    padded[:records.shape[0]] = records
    bufs = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
