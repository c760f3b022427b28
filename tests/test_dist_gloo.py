"""CPU, world_size 2 over gloo: the N>1 bookkeeping of bench.py (max-over-ranks timing, summed work) and the
largest-first round-robin sharding rule the multi-GPU engine uses."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from ampliconreconstructorom_b200 import dist_util
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cells = np.array([10, 500, 30, 70, 90, 20, 400, 60], dtype=np.int64)
    mine = dist_util.shard_problems(cells, world, rank)
    t, w = dist_util.reduce_over_ranks([100.0 + 50.0 * rank, 7.0], [float(cells[mine].sum()), float(mine.size)])
    q.put((rank, mine.tolist(), t, w))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduction_and_sharding():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, t0, w0), (r1, m1, t1, w1) = out
    assert sorted(m0 + m1) == list(range(8)) and not set(m0) & set(m1)
    assert m0[0] == 1 and m1[0] == 6  # the two largest problems go to different ranks
    assert t0 == t1 == [150.0, 7.0]   # max over ranks
    assert w0 == w1 == [1180.0, 8.0]  # summed work
