"""Multi-GPU host logic on CPU: sharding rules, and a world_size-2 gloo run of the bench's
barrier / max-over-ranks / whole-job aggregation (no data-path collective exists to test)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jmcodec_b200.shard import frames_for_rank, streams_for_rank


def test_streams_partition():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            mine = streams_for_rank(32, r, world)
            assert all(s % world == r for s in mine)
            seen += mine
        assert sorted(seen) == list(range(32))
    with pytest.raises(ValueError):
        streams_for_rank(32, 2, 2)


def test_frames_partition_ragged():
    for n in (0, 1, 7, 64, 300):
        for world in (1, 2, 3, 8):
            parts = [frames_for_rank(n, r, world) for r in range(world)]
            flat = [f for p in parts for f in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    mine = streams_for_rank(32, rank, world)
    frames = len(mine) * 300
    ms = 10.0 * (rank + 1)                       # pretend rank 1 is slower
    total_frames, max_ms = bench.aggregate(frames, ms, torch.device("cpu"))
    dist.barrier()
    q.put((rank, mine, total_frames, max_ms))
    dist.destroy_process_group()


def test_gloo_world2_aggregation():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(res[0][1] + res[1][1]) == list(range(32))
    for _, _, total, mx in res:
        assert total == 32 * 300          # whole-job units
        assert mx == 20.0                 # max over ranks, not rank 0's own time
