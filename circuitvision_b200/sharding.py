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


def _as_list(out) -> List:
    """fn may return a list or an iterator of lists of consecutive results (one list per device batch)."""
    if isinstance(out, list):
        return out
    flat: List = []
    for piece in out:
        flat.extend(piece)
    return flat


def run_sharded(items: Sequence, fn: Callable[[Sequence, range], List], group=None, dst: int = 0,
                stream_chunk: Optional[int] = None) -> Optional[List]:
    """Apply `fn(items[lo:hi], range(lo, hi))` on this rank's shard (it must return one result per item, as a list or as an
    iterator of lists of consecutive results) and gather the per-item results on rank `dst` in item order.  Returns the full
    list on `dst`, None elsewhere.  Works without an initialised process group (single process: returns fn's list).

    stream_chunk = k: results travel to `dst` in pieces of k items while `fn` is still producing the rest (a helper thread owns
    `group` for the duration of the call and runs one object gather per piece), so that only the last piece's pickling /
    unpickling is left after the last device batch instead of the whole shard's.  Every rank runs the same number of gathers
    (short shards send empty pieces).  `group` must not be used by any other thread until the call returns."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        out = _as_list(fn(items, range(len(items))))
        if len(out) != len(items):
            raise ValueError("fn must return one result per item")
        return list(out)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    r = shard_range(len(items), world, rank)
    if not stream_chunk or stream_chunk <= 0:
        mine = _as_list(fn(items[r.start:r.stop], r))
        if len(mine) != len(r):
            raise ValueError("fn must return one result per item")
        gathered = [None] * world if rank == dst else None
        dist.gather_object((r.start, list(mine)), gathered, dst=dst, group=group)
        if rank != dst:
            return None
        parts = gathered
    else:
        import queue
        import threading
        per = -(-len(items) // world) if len(items) > 0 else 0
        n_chunks = -(-per // stream_chunk) if per else 0
        q: "queue.Queue" = queue.Queue()
        parts: List = []
        errors: List = []

        def sender():
            try:
                for _ in range(n_chunks):
                    piece = q.get()
                    if piece is None:  # the producer failed: leave (the other ranks will time out on the gather)
                        return
                    gathered = [None] * world if rank == dst else None
                    dist.gather_object(piece, gathered, dst=dst, group=group)
                    if rank == dst:
                        parts.extend(gathered)
            except BaseException as e:  # noqa: BLE001 - re-raised on the calling thread
                errors.append(e)

        th = threading.Thread(target=sender, name="cv-shard-gather", daemon=True)
        th.start()
        mine: List = []
        sent = 0
        try:
            out = fn(items[r.start:r.stop], r)
            for piece in ([out] if isinstance(out, list) else out):
                mine.extend(piece)
                while sent < n_chunks - 1 and len(mine) >= (sent + 1) * stream_chunk:
                    q.put((r.start + sent * stream_chunk, mine[sent * stream_chunk:(sent + 1) * stream_chunk]))
                    sent += 1
            if len(mine) != len(r):
                raise ValueError("fn must return one result per item")
            while sent < n_chunks:  # the rest (possibly empty pieces: every rank runs n_chunks gathers)
                q.put((r.start + sent * stream_chunk, mine[sent * stream_chunk:(sent + 1) * stream_chunk]))
                sent += 1
        except BaseException:
            q.put(None)
            raise
        th.join()
        if errors:
            raise errors[0]
        if rank != dst:
            return None
    full: List = [None] * len(items)
    for start, part in parts:
        full[start:start + len(part)] = part
    return full
