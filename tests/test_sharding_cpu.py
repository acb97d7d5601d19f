"""world_size-2 gloo test of the N>1 host logic (frame ownership + the single counter all-gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vehicle_counting_b200.sharding import frames_of_rank, gather_counters, max_over_ranks
    mine = list(frames_of_rank(total, rank, world))
    dets = sum(f % 7 for f in mine)                      # stand-in for per-frame detection counts
    per_rank = gather_counters([len(mine), dets, 64 * len(mine)])
    slowest = max_over_ranks([10.0 + rank, 3.0 - rank])
    q.put((rank, mine, per_rank, slowest))
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_sharding_and_counter_allgather():
    world, total = 2, 37
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = sorted(f for _, mine, _, _ in res for f in mine)
    assert owned == list(range(total))                              # every frame exactly once
    assert all(f % world == rank for rank, mine, _, _ in res for f in mine)
    for rank, mine, per_rank, slowest in res:
        assert per_rank == res[0][2]                                # identical on every rank
        assert [c[0] for c in per_rank] == [19, 18]
        assert sum(c[1] for c in per_rank) == sum(f % 7 for f in range(total))
        assert slowest == [11.0, 3.0]


def test_single_process_degenerates_cleanly():
    from vehicle_counting_b200.sharding import frames_per_rank, gather_counters, max_over_ranks
    assert frames_per_rank(10, 4) == [3, 3, 2, 2]
    assert gather_counters([1, 2, 3]) == [[1, 2, 3]]
    assert max_over_ranks([1.5]) == [1.5]
