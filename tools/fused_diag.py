"""Where does the fused step's time go at W > 1?  Timing-only diagnostic (totals are wrong
under the flags) against a -DB9GW_DIAG build of the library.  Launch with torchrun:

    B9GW_LIB=build/libb9_diag.so torchrun --nproc-per-node 2 ... tools/fused_diag.py

flags: 1 = the finishing warp does not wait for remote shards; 2 = no packets to peers;
4 = __threadfence_system() after a warp's pushes; 8 = packets as two 8-byte system-scope exchanges.
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from base_b200 import groundwork as gw, vshards  # noqa: E402

if os.environ.get("B9GW_LIB"):
    gw.LIB_PATH = Path(os.environ["B9GW_LIB"]).resolve()

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
with vshards.PeerComm(local, rank, world, 64, max_chains=1024) as comm:
    for flags in (0, 3, 0):   # 2 alone would wait for packets nobody sends
        for chains in (16, 128):
            os.environ["B9GW_DIAG_FLAGS"] = str(flags)
            dist.barrier()
            r = comm.sharded_step(10_000, 1024, chains, warmup=5, reps=40)
            us = torch.tensor([r["us_lse_alone"], r["us_step"], r["us_fused_step"]], device="cuda")
            dist.all_reduce(us, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f"flags={flags} chains={chains}: lse {us[0]:.1f}  two-launch {us[1]:.1f}  fused {us[2]:.1f} us", flush=True)
            if chains == 16 and flags == 0 and os.environ.get("B9GW_DIAG_TRACE"):
                # device timestamps of the last 8 fused steps (ring indexed by step & 7), chains 0 and 15
                import ctypes as C
                import numpy as np
                t = np.zeros(8 * 16 * 4, dtype=np.uint64)
                L = gw.lib()
                L.b9gw_diag_dump.argtypes = [C.c_int, C.c_longlong, C.c_void_p]
                assert L.b9gw_diag_dump(local, 16, t.ctypes.data) == 0
                t = t.reshape(8, 16, 4).astype(np.int64)
                order = np.argsort(t[:, 0, 3])                 # steps in time order
                base = t[order[0], 0, 3]
                for r_ in range(world):
                    dist.barrier()
                    if r_ == rank:
                        print(f"  rank {rank} (us on this GPU's clock; globaltimer base {base}):")
                        print("    step: kernel start | ch0 local done, lane0 packets, total | ch15 first CTA, local done, lane0 packets, total | next start - this total")
                        for i, st_ in enumerate(order):
                            u = lambda c, k: (t[st_, c, k] - base) / 1e3
                            nxt = (t[order[i + 1], 0, 3] - t[st_, 15, 2]) / 1e3 if i + 1 < 8 else float("nan")
                            print(f"    {i}: {u(0,3):8.1f} | {u(0,0):8.1f} {u(0,1):8.1f} {u(0,2):8.1f} | {u(15,3):8.1f} {u(15,0):8.1f} {u(15,1):8.1f} {u(15,2):8.1f} | {nxt:6.1f}", flush=True)
    dist.barrier()
dist.destroy_process_group()
