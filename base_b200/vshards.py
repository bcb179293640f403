"""World-size-independent sum of per-chain FP64 scalars over ranks — host side.

north_star's only collective is "a single NCCL allreduce of per-chain
log-likelihood scalars over NVLink per step" (SURVEY.md §8e).  The accept/reject
decision downstream compares that sum with a uniform draw, so its last bit
matters, and a sum of per-rank partials — in rank order or in whatever order
NCCL picks — groups the stars differently at every world size.  Here the sum is
defined over V fixed VIRTUAL SHARDS instead (include/b9_groundwork.h):

    shard v      = stars [floor(v*N/V), floor((v+1)*N/V))        -- (N, V) only
    rank r owns  = shards r*V/W .. (r+1)*V/W - 1                  -- V % W == 0
    total[chain] = (((0 + P[0]) + P[1]) + ...) + P[V-1]           -- no W anywhere

so the same per-star values give the same bits at W = 1, 2, 4, 8.

Two implementations of the cross-rank step, bit-identical by construction:

`PeerComm`              the product path: libb9_groundwork.so's single kernel over
                        NVLink peer memory (csrc/vshard.cu), called through the
                        C-ABI with raw device pointers and a raw stream.  torch is
                        used for device memory and for publishing the 64-byte
                        handles only.  No fallback: it raises without the library
                        or without a GPU.
`allgather_ordered_sum` the same sum stated with torch.distributed (all-gather of
                        the [V, chains] partials, then V-1 ordered adds).  It
                        exists so the ownership arithmetic and the order can be
                        tested on CPU with gloo at several world sizes, and as
                        the NCCL comparison line in bench.py.  It is not called
                        on the step path.

This is reference-independent plumbing (the hot path is BLOCKED); it
does not depend on what the per-star values mean.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

DEFAULT_VSHARDS = 64


def check_layout(n_vshards: int, world: int) -> None:
    if n_vshards < 4 or n_vshards > 128 or n_vshards & (n_vshards - 1):
        raise ValueError("n_vshards must be a power of two in [4, 128]")
    if world < 1 or n_vshards % world:
        raise ValueError(f"world size {world} does not divide n_vshards {n_vshards}")


def shard_lo(n_stars: int, n_vshards: int, shard: int) -> int:
    return shard * n_stars // n_vshards


def shard_bounds(n_stars: int, n_vshards: int) -> list[tuple[int, int]]:
    """[lo, hi) of every virtual shard; a function of (n_stars, n_vshards) alone."""
    return [(shard_lo(n_stars, n_vshards, v), shard_lo(n_stars, n_vshards, v + 1))
            for v in range(n_vshards)]


def owned_shards(rank: int, world: int, n_vshards: int) -> range:
    check_layout(n_vshards, world)
    per = n_vshards // world
    return range(rank * per, (rank + 1) * per)


def local_star_range(rank: int, world: int, n_stars: int, n_vshards: int) -> tuple[int, int]:
    own = owned_shards(rank, world, n_vshards)
    return shard_lo(n_stars, n_vshards, own.start), shard_lo(n_stars, n_vshards, own.stop)


def allgather_ordered_sum(partials: torch.Tensor, group=None) -> torch.Tensor:
    """partials: this rank's [V/W, chains] float64 -> total [chains], same bits on every rank
    and for every W.  One all-gather, then shards 0..V-1 added left to right starting from +0."""
    if partials.dtype != torch.float64 or partials.dim() != 2:
        raise ValueError("partials must be a [V/W, chains] float64 tensor")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    per, chains = partials.shape
    if world == 1:
        gathered = partials
    else:
        flat = torch.empty(world * per * chains, dtype=partials.dtype, device=partials.device)
        dist.all_gather_into_tensor(flat, partials.contiguous().view(-1), group=group)
        gathered = flat.view(world * per, chains)      # rank-major == shard order
    total = torch.zeros(chains, dtype=partials.dtype, device=partials.device)
    for v in range(gathered.shape[0]):                 # explicit order; torch.sum's is unspecified
        total += gathered[v]
    return total


class PeerComm:
    """b9gw_comm_* / b9gw_ordered_allreduce through the C-ABI (include/b9_groundwork.h)."""

    def __init__(self, device: int, rank: int, world: int, n_vshards: int = DEFAULT_VSHARDS,
                 max_chains: int = 1024, group=None, timeout_ms: int = 2000):
        from . import groundwork as gw                 # raises if the .so is missing: no fallback
        check_layout(n_vshards, world)
        self._gw, self._L = gw, gw.lib()
        self.device, self.rank, self.world = device, rank, world
        self.n_vshards, self.max_chains = n_vshards, max_chains
        self._h = C.c_void_p()
        handle = (C.c_char * gw.IPC_HANDLE_BYTES)()
        gw._ck(self._L.b9gw_comm_create(device, rank, world, n_vshards, max_chains,
                                        C.byref(self._h), handle))
        gw._ck(self._L.b9gw_comm_set_timeout_ms(self._h, timeout_ms))
        if world > 1:
            # publish the opaque handles; any transport would do, this one is already up
            mine = torch.frombuffer(bytearray(bytes(handle)), dtype=torch.uint8).to(f"cuda:{device}")
            every = torch.empty(world * gw.IPC_HANDLE_BYTES, dtype=torch.uint8, device=mine.device)
            dist.all_gather_into_tensor(every, mine, group=group)
            blob = bytes(every.cpu().numpy().tobytes())
            gw._ck(self._L.b9gw_comm_connect(self._h, blob))
            dist.barrier(group=group)                  # everyone mapped before anyone pushes

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def shard_partials(self, values: torch.Tensor, n_stars_total: int) -> torch.Tensor:
        """values: [chains, n_local] float64 on this rank's GPU (its own stars only)."""
        if values.dtype != torch.float64 or values.dim() != 2 or not values.is_cuda:
            raise ValueError("values must be a [chains, n_local] float64 CUDA tensor")
        own = owned_shards(self.rank, self.world, self.n_vshards)
        chains = values.shape[0]
        out = torch.empty(len(own), chains, dtype=torch.float64, device=values.device)
        self._gw._ck(self._L.b9gw_shard_partials(
            self.device, values.data_ptr(), chains, values.stride(0), n_stars_total, self.n_vshards,
            own.start, len(own), out.data_ptr(), self._stream()))
        return out

    def allreduce(self, partials: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """partials [V/W, chains] -> total [chains]; one kernel on the current stream, not synced."""
        per = self.n_vshards // self.world
        if partials.dtype != torch.float64 or partials.shape[0] != per or not partials.is_contiguous():
            raise ValueError(f"partials must be a contiguous [{per}, chains] float64 tensor")
        chains = partials.shape[1]
        if out is None:
            out = torch.empty(chains, dtype=torch.float64, device=partials.device)
        self._gw._ck(self._L.b9gw_ordered_allreduce(self._h, partials.data_ptr(), out.data_ptr(),
                                                    chains, self._stream()))
        return out

    def latency(self, chains: int, warmup: int = 20, reps: int = 200) -> dict:
        a, b = C.c_float(), C.c_float()
        self._gw._ck(self._L.b9gw_allreduce_latency(self._h, chains, warmup, reps,
                                                    C.byref(a), C.byref(b)))
        return {"us_stream": a.value, "us_graph": b.value,
                "launches": warmup + reps + 8 * (1 + (reps + 7) // 8)}

    def sharded_step(self, n_stars_total: int, cols: int, chains: int, warmup: int = 3,
                     reps: int = 20) -> dict:
        """b9gw_sharded_step: this rank's share of the star-sharded log-sum-exp job, then the
        cross-rank sum; total has the same bits on every rank and at every world size."""
        import numpy as np
        total, fused = np.empty(chains, dtype=np.float64), np.empty(chains, dtype=np.float64)
        a, b, f = C.c_float(), C.c_float(), C.c_float()
        pd = C.POINTER(C.c_double)
        self._gw._ck(self._L.b9gw_sharded_step(self._h, n_stars_total, cols, chains, warmup, reps,
                                               total.ctypes.data_as(pd), fused.ctypes.data_as(pd),
                                               C.byref(a), C.byref(b), C.byref(f)))
        return {"total": total, "total_fused": fused, "us_step": a.value, "us_lse_alone": b.value,
                "us_fused_step": f.value, "launches": 4 * (warmup + reps)}

    def lse_generated_step(self, n_stars_total: int, cols: int, chains: int, row_lse: torch.Tensor,
                           partials: torch.Tensor, total: torch.Tensor, workspace: torch.Tensor) -> None:
        """b9gw_lse_generated_step on the current stream: ONE kernel = this rank's share of the
        log-sum-exp job + the cross-rank sum.  Buffers are the caller's CUDA tensors (workspace
        zero before the first launch)."""
        self._gw._ck(self._L.b9gw_lse_generated_step(
            self._h, n_stars_total, cols, chains, row_lse.data_ptr(), partials.data_ptr(),
            total.data_ptr(), workspace.data_ptr(), self._stream()))

    def status(self) -> dict:
        """Synchronises; raises GroundworkError(E_TIMEOUT) if any step gave up waiting."""
        flag, steps = C.c_int(), C.c_ulonglong()
        self._gw._ck(self._L.b9gw_comm_status(self._h, C.byref(flag), C.byref(steps)))
        return {"timed_out": bool(flag.value), "steps": steps.value}

    def close(self) -> None:
        if self._h:
            self._L.b9gw_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
