"""CPU suite: the N>1 host logic (stream sharding + result gather) with world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ygz_slam_b200 import dist as ydist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_streams, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = ydist.shard_streams(n_streams, rank, world)
    recs = [ydist.make_record(s, 300, 1000 + s, 800 + s, [s, 0, 0, 1, 0, 0, 0], 10.0 * (rank + 1) + s) for s in mine]
    allr, frames, ms = ydist.gather_records(recs, n_streams)
    q.put((rank, mine, allr, frames, ms))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_streams", [8, 5])
def test_shard_and_gather_world2(n_streams):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_streams, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda t: t[0])
    owned = sorted(res[0][1] + res[1][1])
    assert owned == list(range(n_streams))                      # every stream exactly once
    assert np.array_equal(res[0][2], res[1][2])                 # both ranks hold the same gathered table
    assert np.array_equal(res[0][2][:, 0], np.arange(n_streams))
    assert res[0][3] == res[1][3] == 300.0 * n_streams
    want_ms = max(sum(10.0 * (r + 1) + s for s in ydist.shard_streams(n_streams, r, 2)) for r in range(2))
    assert res[0][4] == res[1][4] == want_ms


def test_single_process_gather():
    recs = [ydist.make_record(s, 10, 1, 2, [0] * 7, 1.0) for s in range(3)]
    allr, frames, ms = ydist.gather_records(recs, 3)
    assert allr.shape == (3, ydist.RECORD_DOUBLES) and frames == 30.0 and ms == 3.0
