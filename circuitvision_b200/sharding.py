"""Image-wise sharding of a batch over the GPUs of one node.

Images / crops are independent in both halves of the path, so there is no data-path collective: every rank (one
process per GPU, `torch.distributed`) takes a contiguous block of ceil(N / world) image indices, runs the hot path on
its own device, and the small per-image results (node tables, contours, netlist inputs) are gathered on the HOST of
rank 0 by index.  NVLink / NCCL never carries image data; the only collective is the object gather of the results."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence


def shard_range(n_items: int, world: int, rank: int) -> range:
    """Contiguous block of item indices owned by `rank` (the last ranks may own fewer, possibly zero, items)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    per = -(-n_items // world) if n_items > 0 else 0
    lo = min(n_items, rank * per)
    return range(lo, min(n_items, lo + per))


def run_sharded(items: Sequence, fn: Callable[[Sequence, range], List], group=None, dst: int = 0) -> Optional[List]:
    """Apply `fn(items[lo:hi], range(lo, hi))` on this rank's shard (it must return one result per item) and gather
    the per-item results on rank `dst` in item order.  Returns the full list on `dst`, None elsewhere.  Works without
    an initialised process group (single process: returns fn's list)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        out = fn(items, range(len(items)))
        if len(out) != len(items):
            raise ValueError("fn must return one result per item")
        return list(out)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    r = shard_range(len(items), world, rank)
    mine = fn(items[r.start:r.stop], r)
    if len(mine) != len(r):
        raise ValueError("fn must return one result per item")
    gathered = [None] * world if rank == dst else None
    dist.gather_object((r.start, list(mine)), gathered, dst=dst, group=group)
    if rank != dst:
        return None
    full: List = [None] * len(items)
    for start, part in gathered:
        full[start:start + len(part)] = part
    return full
