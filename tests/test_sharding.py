"""CPU, world_size 2 on gloo: the N>1 plumbing of bench.py / dataset generation -- whole reference batches per
rank, every batch exactly once, max-over-ranks timing, whole-job throughput = all ranks' units / max time."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torch_fdtd_string_b200.parallel import rank_batches, max_over_ranks, sum_over_ranks


def test_rank_batches_partition():
    for n in (0, 1, 4, 7, 100, 1184):
        for w in (1, 2, 3, 4, 8):
            parts = [list(rank_batches(n, w, r)) for r in range(w)]
            flat = [b for p in parts for b in p]
            assert flat == list(range(n)), (n, w)
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, n_batches, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = list(rank_batches(n_batches, world, rank))
    # each rank "simulates" its batches: unit count and a fake device time that differs per rank
    units = 24 * len(mine)
    t_local = 1.0 + 0.5 * rank
    t_max = max_over_ranks(t_local)
    total = sum_over_ranks(units)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    dist.barrier()
    q.put((rank, t_max, total, gathered))
    dist.destroy_process_group()


def test_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_batches, world = 7, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_batches, q)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs: p.join(timeout=60)
    for rank, t_max, total, gathered in res:
        assert t_max == 1.5                       # max over ranks, not the local time
        assert total == 24 * n_batches            # whole-job units
        assert sorted(b for part in gathered for b in part) == list(range(n_batches))
