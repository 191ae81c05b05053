"""Read sharding for multi-GPU runs: the index is replicated on every GPU,
the read set is cut into contiguous blocks, one per rank, and results are reassembled in read
order on rank 0.  There is no exchange step on the data path; torch.distributed only carries
the (small) gathered results and the barrier."""
from __future__ import annotations

from typing import Callable, List, Sequence


def shard_bounds(n_items: int, world: int, rank: int):
    """Contiguous block [a, b) of rank `rank`: sizes differ by at most one, order preserved."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    a = rank * base + min(rank, extra)
    return a, a + base + (1 if rank < extra else 0)


def run_sharded(items: Sequence, engine: Callable[[Sequence], List], rank: int, world: int, group=None):
    """Runs `engine` on this rank's block and returns, on rank 0, the concatenation of all ranks'
    per-item results in the original order (None elsewhere).  `engine` is the per-GPU query
    (GpuIndex.query wrapped by the caller)."""
    a, b = shard_bounds(len(items), world, rank)
    local = engine(items[a:b])
    if len(local) != b - a:
        raise RuntimeError("engine returned %d results for %d items" % (len(local), b - a))
    if world == 1:
        return list(local)
    import torch.distributed as dist
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(list(local), gathered, dst=0, group=group)
    if rank != 0:
        return None
    out = []
    for part in gathered:
        out.extend(part)
    return out
