"""Host logic of the multi-GPU path on CPU: world_size 2 over gloo (image sharding + the all-gather)."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from labelany3d_b200 import dist as la_dist
from labelany3d_b200.records import O_STATUS, REC


def test_shard_range_covers_everything_once():
    for total in (1, 2, 7, 8, 256, 2048, 1000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                a, b, per = la_dist.shard_range(total, r, world)
                assert 0 <= a <= b <= total and b - a <= per and per * world >= total
                seen += list(range(a, b))
            assert seen == list(range(total))


def _worker(rank, world, port, total, I, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b, per = la_dist.shard_range(total, rank, world)
        # every rank fabricates the records of ITS images: value = global image index
        local = torch.zeros((b - a, I, REC), dtype=torch.float64)
        for j in range(b - a):
            local[j] = float(a + j)
        local[..., O_STATUS] = 0
        out = la_dist.all_gather_records(local, total=total)
        q.put((rank, out.numpy()))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_all_gather_records_world2():
    for total in (6, 7):          # even split and a ragged one (rank 1 owns one image fewer)
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, total, 3, q)) for r in range(2)]
        for p in procs:
            p.start()
        got = dict(q.get(timeout=120) for _ in procs)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        for r in (0, 1):
            out = got[r]
            assert out.shape == (total, 3, REC)
            want = np.repeat(np.arange(total, dtype=np.float64), 3 * REC).reshape(total, 3, REC)
            want[..., O_STATUS] = 0
            np.testing.assert_array_equal(out, want)     # image order preserved, padding trimmed
