"""CPU, world_size 2 over gloo: the multi-GPU plumbing bench.py uses (frame ranges, MAX/SUM reductions,
gather of per-frame results). No collective touches point data; frames are independent."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_range_partitions():
    from stair_step_detector_b200.sharding import frame_range
    for total in (4096, 1024, 7, 1):
        for world in (1, 2, 4, 8):
            parts = [frame_range(r, world, total) for r in range(world)]
            assert sum(c for _, c in parts) == total
            nxt = 0
            for first, count in parts:
                assert first == nxt
                nxt += count
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from stair_step_detector_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = sharding.frame_range(rank, world, 10)
    # pretend per-frame step counts = global frame id % 5, per-rank time = 10 + rank ms
    local = [(sharding.global_frame_index(rank, 5, i)) % 5 for i in range(5)]
    tmax, tsum = sharding.reduce_timing(dist, [10.0 + rank, 3.0], [7 * (rank + 1), count])
    allc = sharding.gather_step_counts(dist, local)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, first, count, tmax, tsum, allc))


def test_two_ranks_gloo():
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 5), (5, 5)]
    for r in res:
        assert r[3] == [11.0, 3.0] and r[4] == [21, 10]
        assert r[5] == [i % 5 for i in range(10)]
