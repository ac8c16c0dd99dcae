"""World-size-2 gloo test of the N>1 host logic (sharding of pairs, barrier, max-over-ranks timing)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from pmt_learning_for_semantic_segmentation_and_disparity_b200 import sharding


def test_shard_range_partitions_everything():
    for total in (0, 1, 7, 32, 33):
        for ws in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def test_single_process_world():
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        os.environ.pop(k, None)
    w = sharding.init_world()
    assert (w.rank, w.world_size, w.distributed) == (0, 1, False)
    assert sharding.max_over_ranks(w, 3.5) == 3.5
    assert sharding.throughput(w, 8, 2.0) == (4000.0, 2.0)


def _worker(rank, world_size, port, q):
    os.environ.update({"RANK": str(rank), "LOCAL_RANK": str(rank), "WORLD_SIZE": str(world_size),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    w = sharding.init_world("gloo")
    b, e = sharding.shard_range(9, w.rank, w.world_size)
    sharding.barrier(w)
    # rank r "took" (r+1) ms for its (e-b) pairs: whole-job throughput uses the slowest rank
    val, ms = sharding.throughput(w, e - b, float(rank + 1))
    q.put((rank, b, e, val, ms, sharding.max_over_ranks(w, 10.0 * (rank + 1))))
    sharding.shutdown(w)


def test_two_rank_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 5), (5, 9)]
    for r in res:
        assert r[4] == 2.0 and abs(r[3] - 9 / 2e-3) < 1e-6 and r[5] == 20.0
