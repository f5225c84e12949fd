"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes exercise the weight-blob broadcast and the
clip sharding that bench.py / a serving process use over NCCL on the GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nunet_b200.sharding import broadcast_blob, shard_range


def test_shard_ranges_partition_everything():
    for n in (0, 1, 7, 8, 255, 256, 8192):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        blob = rng.integers(0, 256, size=100_003, dtype=np.uint8).tobytes() if rank == 0 else None
        got = broadcast_blob(blob, src=0)
        b, e = shard_range(n_clips, rank, world)
        # every rank "processes" its clips; the per-rank counts are summed like bench.py sums frames
        t = torch.tensor([e - b], dtype=torch.int64)
        dist.all_reduce(t)
        q.put((rank, len(got), int(np.frombuffer(got, np.uint8).astype(np.int64).sum()), (b, e), int(t.item())))
    finally:
        dist.destroy_process_group()


def test_blob_broadcast_and_sharding_world2_gloo():
    world, n_clips = 2, 513
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] == 100_003 and res[0][2] == res[1][2]
    assert res[0][3] == (0, 257) and res[1][3] == (257, 513)
    assert res[0][4] == res[1][4] == n_clips


def test_single_process_passthrough():
    assert broadcast_blob(b"abc") == b"abc"
