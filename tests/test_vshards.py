"""The world-size-independent sum over virtual shards, on CPU.

The claim under test: the same global per-star values give the same BITS at world sizes
1, 2 and 4 (gloo), equal to the CPU checker's, whereas a rank-ordered sum of per-rank
partials — what round 1 shipped — does not.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from base_b200 import vshards

N_STARS, CHAINS, V = 1003, 37, 64          # ragged: 1003 is not a multiple of 64


def global_values():
    """[chains, stars], magnitudes spread over 12 decades so that grouping changes the last bits."""
    rng = np.random.default_rng(2024)
    return rng.normal(size=(CHAINS, N_STARS)) * 10.0 ** rng.integers(-6, 6, size=(CHAINS, N_STARS))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_partials(ref, values, rank, world):
    own = vshards.owned_shards(rank, world, V)
    bounds = vshards.shard_bounds(N_STARS, V)
    lo0, hi0 = vshards.local_star_range(rank, world, N_STARS, V)
    local = np.ascontiguousarray(values[:, lo0:hi0])      # a rank holds its own stars only
    P = np.empty((len(own), CHAINS))
    for k, v in enumerate(own):
        lo, hi = bounds[v]
        for c in range(CHAINS):
            P[k, c] = ref.shard_partial(local[c], lo - lo0, hi - lo0)
    return torch.from_numpy(P), local


def _worker(rank, world, port, ref_path, q):
    from tests import _ref
    ref = _ref.load(ref_path)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, local = _rank_partials(ref, global_values(), rank, world)
    total = vshards.allgather_ordered_sum(P)
    # round 1's method, for contrast: one partial per RANK, added in rank order
    mine = torch.from_numpy(np.array([ref.serial_sum(local[c]) for c in range(CHAINS)]))
    flat = torch.empty(world * CHAINS, dtype=torch.float64)
    dist.all_gather_into_tensor(flat, mine)
    naive = flat.view(world, CHAINS)[0].clone()
    for r in range(1, world):
        naive += flat.view(world, CHAINS)[r]
    q.put((rank, total.numpy().copy(), naive.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, ref_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, str(ref_path), q)) for r in range(world)]
    [p.start() for p in ps]
    got = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    return got


def bits(a):
    return np.asarray(a, dtype=np.float64).view(np.int64)


def test_same_bits_at_world_1_2_4_and_equal_to_the_checker(built, ref):
    values = global_values()
    _, want = ref.vshard_total(values, V)
    P, _ = _rank_partials(ref, values, 0, 1)
    w1 = vshards.allgather_ordered_sum(P).numpy()           # world 1, no process group
    assert (bits(w1) == bits(want)).all()
    naive_by_world = {}
    for world in (2, 4):
        got = _run(world, built.REF_LIB)
        for _, total, naive in got:
            assert (bits(total) == bits(want)).all(), f"world {world}"
            assert (bits(naive) == bits(got[0][2])).all()   # rank-ordered: same on every rank...
        naive_by_world[world] = got[0][2]
    # ...but NOT the same across world sizes: this is what the virtual shards fix
    assert (bits(naive_by_world[2]) != bits(naive_by_world[4])).any()


def test_bounds_partition_the_stars_and_agree_with_c_and_checker(built, ref):
    from base_b200 import groundwork as gw
    for n, v in [(0, 4), (1, 4), (3, 8), (1003, 64), (10_000, 64), (50_000, 128), (2 ** 40 + 7, 128)]:
        b = vshards.shard_bounds(n, v)
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(v - 1))
        assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
        for s in (0, v // 2, v - 1):
            assert gw.vshard_bounds(n, v, s) == b[s]          # pure host arithmetic in the .so
            assert (ref.shard_lo(n, v, s), ref.shard_lo(n, v, s + 1)) == b[s]


def test_ownership_is_contiguous_and_covers_every_shard():
    for world in (1, 2, 4, 8):
        own = [vshards.owned_shards(r, world, 64) for r in range(world)]
        assert [v for o in own for v in o] == list(range(64))
        edges = [vshards.local_star_range(r, world, 10_000, 64) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == 10_000
        assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))


@pytest.mark.parametrize("v,world", [(64, 3), (48, 2), (2, 1), (256, 2), (64, 0)])
def test_bad_layouts_are_rejected(v, world):
    with pytest.raises(ValueError):
        vshards.check_layout(v, world)


def test_c_abi_rejects_bad_comm_arguments_before_touching_a_device(built):
    import ctypes as C
    from base_b200 import groundwork as gw
    L, h, buf = gw.lib(), C.c_void_p(), (C.c_char * 64)()
    for rank, world, v, chains in [(0, 3, 64, 8), (2, 2, 64, 8), (0, 1, 48, 8), (0, 1, 64, 0),
                                   (0, 17, 64, 8), (0, 1, 256, 8)]:
        assert L.b9gw_comm_create(0, rank, world, v, chains, C.byref(h), buf) == gw.E_ARG
        assert not h.value
    assert L.b9gw_ordered_allreduce(None, None, None, 1, None) == gw.E_ARG
    assert L.b9gw_vshard_bounds(10, 48, 0, None, None) == gw.E_ARG
