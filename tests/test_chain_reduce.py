"""world_size-2 gloo test of the order-fixed cross-rank sum."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from base_b200.chain_reduce import ordered_allreduce_sum
    g = torch.Generator().manual_seed(100 + rank)
    # magnitudes chosen so that (a+b)+c != a+(b+c) in the last bits
    part = torch.randn(1024, dtype=torch.float64, generator=g) * (10.0 ** (3 * rank))
    out = ordered_allreduce_sum(part)
    q.put((rank, part, out))
    dist.barrier()
    dist.destroy_process_group()


def _run(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    got = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    return got


def test_every_rank_gets_the_same_bits_in_rank_order():
    got = _run(2)
    want = got[0][1].clone()
    for _, part, _ in got[1:]:
        want += part
    for _, _, out in got:
        assert torch.equal(out.view(torch.int64), want.view(torch.int64))


def test_three_ranks_left_to_right():
    got = _run(3)
    want = (got[0][1] + got[1][1]) + got[2][1]
    for _, _, out in got:
        assert torch.equal(out.view(torch.int64), want.view(torch.int64))
