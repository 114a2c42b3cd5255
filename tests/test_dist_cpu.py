"""World-size-2 gloo test of the N>1 host logic: request sharding + the single result all_gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vitron_b200.dist import gather_results, shard_range, shard_requests


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    reqs = list(range(n_total))
    mine = shard_requests(reqs)
    local = torch.tensor([[r * 10 + t for t in range(4)] for r in mine], dtype=torch.int64).reshape(len(mine), 4)
    full = gather_results(local, n_total)
    q.put((rank, mine, full.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 8, 64):
        for w in (1, 2, 3, 8):
            got = [i for r in range(w) for i in range(*shard_range(n, r, w))]
            assert got == list(range(n))


def test_two_rank_gather_restores_request_order():
    for n_total in (8, 5):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        want = [[r * 10 + t for t in range(4)] for r in range(n_total)]
        seen = []
        for rank, mine, full in res:
            assert full == want
            seen += mine
        assert sorted(seen) == list(range(n_total))
