"""N-rank check of the peer-memory all-reduce (one rank per GPU; launch with torchrun).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/multigpu/peer_comm_ranks.py

Never run this with more ranks than GPUs: ranks that wait on one another cannot share a
device (B200_PROFILING.md).  Checks, on every rank:
  1. the N-rank total has the bits of the checker's world-independent total;
  2. 300 back-to-back steps with changing data stay correct (mailbox parity, step counter);
  3. a CUDA-graph replay of the step stays correct;
  3b. a star-sharded step (log-sum-exp of the rank's own shards, then the cross-rank sum) has
     the bits of the same job run by one rank alone;
  4. a peer that skips a step costs the others one timeout, NaN outputs and a sticky status —
     not a hang.
Prints one JSON line per rank-0 result.
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from base_b200 import build, groundwork as gw, vshards  # noqa: E402
from tests import _ref  # noqa: E402


def bits(t):
    return t.detach().cpu().numpy().view(np.int64)


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    assert torch.cuda.device_count() >= world, "one GPU per rank is required"
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ref = _ref.load(build.REF_LIB)
    chains, n, V = 513, 10_007, 64
    rng = np.random.default_rng(77)
    values = rng.normal(size=(chains, n)) * 10.0 ** rng.integers(-6, 6, size=(chains, n))
    _, want = ref.vshard_total(values, V)
    lo, hi = vshards.local_star_range(rank, world, n, V)
    dv = torch.from_numpy(np.ascontiguousarray(values[:, lo:hi])).cuda()
    report = {"world": world}

    with vshards.PeerComm(local, rank, world, V, max_chains=1024) as comm:
        P = comm.shard_partials(dv, n)
        total = comm.allreduce(P)
        comm.status()
        assert (bits(total) == want.view(np.int64)).all(), f"rank {rank}: bits differ from the checker"
        assert torch.equal(vshards.allgather_ordered_sum(P).view(torch.int64), total.view(torch.int64))
        report["bits_equal_checker"] = True

        # 2. many steps, data changing every step: scale the partials by exact powers of two
        out = torch.empty(chains, dtype=torch.float64, device="cuda")
        for step in range(300):
            k = float(2 ** (step % 7))
            comm.allreduce(P * k, out)
            if step % 50 == 49:
                assert (bits(out / k) == want.view(np.int64)).all(), f"rank {rank} step {step}"
        comm.status()
        report["steps_ok"] = 301

        # 3. graph replay
        g, s = torch.cuda.CUDAGraph(), torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            comm.allreduce(P, out)
            with torch.cuda.graph(g, stream=s):
                comm.allreduce(P, out)
        for _ in range(5):
            out.zero_()
            g.replay()
            torch.cuda.synchronize()
            assert (bits(out) == want.view(np.int64)).all()
        report["graph_replay_ok"] = True
        # 3b. a star-sharded step end to end: every rank computes ITS shards' log-sum-exp rows and
        # partials on chip, then the cross-rank sum; the total must have the bits this rank gets
        # alone, from all 64 shards in one launch (a 1-rank job on its own GPU)
        sn, sc, sch = 3_001, 160, 9
        step = comm.sharded_step(sn, sc, sch, warmup=2, reps=10)
        alone = gw.lse_generated_shards(sn, sc, sch, V, 0, V, local)["total"]
        assert (step["total"].view(np.int64) == alone.view(np.int64)).all(), f"rank {rank}: sharded step"
        assert (step["total_fused"].view(np.int64) == alone.view(np.int64)).all(), f"rank {rank}: fused step"
        # fewer stars than shards: most ranks own no star at all and only pull
        tiny = comm.sharded_step(5, 40, 3, warmup=1, reps=3)
        tiny_alone = gw.lse_generated_shards(5, 40, 3, V, 0, V, local)["total"]
        assert (tiny["total"].view(np.int64) == tiny_alone.view(np.int64)).all()
        assert (tiny["total_fused"].view(np.int64) == tiny_alone.view(np.int64)).all(), f"rank {rank}: tiny fused"
        report["fused_step_us"] = round(step["us_fused_step"], 2)
        # 3c. the fused step through caller-held tensors, replayed from a CUDA graph (the step
        # counters live in device memory), mixed with two-launch steps on the same comm
        per = V // world
        lo_s, hi_s = vshards.local_star_range(rank, world, sn, V)
        rows_t = torch.empty(sch, max(hi_s - lo_s, 1), dtype=torch.float64, device="cuda")
        P_t = torch.empty(per, sch, dtype=torch.float64, device="cuda")
        tot_t = torch.empty(sch, dtype=torch.float64, device="cuda")
        work_t = torch.zeros(gw.lib().b9gw_lse_workspace_bytes(sch, per) // 4, dtype=torch.int32, device="cuda")
        g2, s2 = torch.cuda.CUDAGraph(), torch.cuda.Stream()
        s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s2):
            comm.lse_generated_step(sn, sc, sch, rows_t, P_t, tot_t, work_t)
            with torch.cuda.graph(g2, stream=s2):
                comm.lse_generated_step(sn, sc, sch, rows_t, P_t, tot_t, work_t)
                comm.lse_generated_step(sn, sc, sch, rows_t, P_t, tot_t, work_t)
        torch.cuda.synchronize()
        for _ in range(4):
            tot_t.fill_(float("nan"))
            torch.cuda.synchronize()
            g2.replay()
            torch.cuda.synchronize()
            assert (bits(tot_t) == alone.view(np.int64)).all(), f"rank {rank}: fused graph replay"
            assert (bits(comm.allreduce(P_t)) == alone.view(np.int64)).all()
        comm.status()
        report["fused_graph_replay_ok"] = True
        flat = gw.lse_generated(sch * sn, sc, local)["row_lse"].reshape(sch, sn)
        assert (ref.vshard_total(flat, V)[1].view(np.int64) == alone.view(np.int64)).all()
        report["sharded_step_bits_equal_world_1"] = True
        report["sharded_step_us"] = round(step["us_step"], 2)
        lat = comm.latency(chains, warmup=20, reps=400)
        report["us_stream"], report["us_graph"] = round(lat["us_stream"], 2), round(lat["us_graph"], 2)
        dist.barrier()

    # 4. a missing peer: the last rank sits one step out
    with vshards.PeerComm(local, rank, world, V, max_chains=1024, timeout_ms=300) as comm:
        P = comm.shard_partials(dv, n)
        if rank != world - 1:
            out = comm.allreduce(P)
            try:
                comm.status()
                raised = False
            except gw.GroundworkError as e:
                raised = e.code == gw.E_TIMEOUT
            assert raised and torch.isnan(out).all(), f"rank {rank}: missing peer was not reported"
        dist.barrier()
    report["missing_peer_reported"] = True

    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(json.dumps(report))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
