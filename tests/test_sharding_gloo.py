"""N > 1 host logic on CPU: world_size 2 over gloo (stream ownership, timing reductions)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ros2_mono_vo_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.weak_scaling_streams(4, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    slow = sharding.max_over_ranks(10.0 + 5.0 * rank)
    total = sharding.total_over_ranks(len(mine) * 7)
    out.put((rank, mine, gathered, slow, total))
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_ownership():
    assert sharding.streams_for_rank(256, 3, 8) == list(range(3, 256, 8))
    cover = sorted(s for r in range(8) for s in sharding.streams_for_rank(256, r, 8))
    assert cover == list(range(256))
    assert sharding.streams_for_rank(5, 1, 2) == [1, 3]
    with pytest.raises(ValueError):
        sharding.streams_for_rank(4, 2, 2)
    assert sharding.max_over_ranks(3.5) == 3.5 and sharding.total_over_ranks(9) == 9   # no process group


def test_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, g0, s0, t0), (r1, m1, g1, s1, t1) = res
    assert m0 == [0, 2, 4, 6] and m1 == [1, 3, 5, 7]
    assert g0 == g1 == [m0, m1]
    assert s0 == s1 == 15.0                      # max over ranks
    assert t0 == t1 == 56                        # whole-job frame count
