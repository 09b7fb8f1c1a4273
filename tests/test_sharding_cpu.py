"""Host logic of the image-wise multi-GPU path on CPU: world_size-2 gloo processes shard a batch, run a stand-in for the
per-rank work and gather the per-image results on rank 0 in image order (no data-path collective exists to test)."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from circuitvision_b200.sharding import run_sharded, shard_range


def test_shard_range_covers_every_item_once():
    for n in (0, 1, 7, 64, 65, 8192):
        for world in (1, 2, 4, 8):
            seen = []
            for rank in range(world):
                seen += list(shard_range(n, world, rank))
            assert seen == list(range(n)), (n, world)
    assert len(shard_range(65, 8, 0)) == 9 and len(shard_range(65, 8, 7)) == 2
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_single_process_passthrough():
    assert run_sharded(list("abc"), lambda xs, r: [x.upper() + str(i) for x, i in zip(xs, r)]) == ["A0", "B1", "C2"]


def _worker(rank, world, port, n, q, stream_chunk=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        items = [{"seed": i} for i in range(n)]

        def fn(xs, r):  # stand-in for segment + node analysis on this rank's device
            res = [{"image": i, "rank": rank, "nodes": [x["seed"] % 3, i]} for x, i in zip(xs, r)]
            if stream_chunk is None:
                return res
            return (res[k:k + 3] for k in range(0, len(res), 3))  # an iterator of device batches of 3 images

        full = run_sharded(items, fn, stream_chunk=stream_chunk)
        dist.barrier()
        if rank == 0:
            q.put(full)
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,stream_chunk", [(7, None), (64, None), (7, 2), (64, 8), (5, 4)])
def test_two_rank_gloo_gather_in_image_order(n, stream_chunk):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q, stream_chunk)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [f["image"] for f in full] == list(range(n))
    half = -(-n // 2)
    assert [f["rank"] for f in full] == [0] * half + [1] * (n - half)
    assert all(f["nodes"] == [i % 3, i] for i, f in enumerate(full))
