#!/usr/bin/env python
"""bench.py — the driver's measurement contract, answered honestly.

BLOCKED: BASELINE.json's metric ("cluster log-likelihood evals/sec and MCMC
steps/sec") cannot be measured, because the likelihood it refers to does not
exist in this container: /root/reference is a relocation notice
(/root/reference/README.md:1-4), base-cpp is not staged, and north_star forbids
reconstructing it from memory.  So `value`, `e2e`, `roofline` and
`cpu_baseline` are null and `blocked` says why.  No stand-in workload is reported
under the headline metric.

What this run DOES measure, under the separate key `groundwork`, are the
reference-independent denominators north_star demands before any roofline
fraction can be quoted: the B200's FP64 DFMA peak (absent from
MEASURED_PEAKS.json), FP64 exp/log/exp10/log10 rates, the host round trip of one
dependent step, a fixed-order row log-sum-exp fed from memory and from registers,
and the world-size-independent cross-rank sum of 1024 per-chain scalars as one
peer-memory kernel behind the C-ABI (its bits are asserted equal to the 1-rank
sum on every rank, at every N), and a star-sharded step that puts the two together
(each rank's share of a fixed synthetic log-sum-exp job, then the cross-rank sum;
same bits at every N, asserted).  There is no "step": K and W only size the timing
loops (CUDA events on the launching stream, W warm-up launches first).
`gpu_launches` counts those groundwork kernels and nothing else.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from base_b200 import roofline, staging  # noqa: E402


def _baseline_metric() -> str:
    try:
        return json.loads((ROOT / "BASELINE.json").read_text())["metric"]
    except Exception:
        return "cluster log-likelihood evals/sec and MCMC steps/sec at 1/2/4/8 B200 vs host CPU"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "100"], stdout=fd, stderr=subprocess.DEVNULL)
            os.close(fd)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self) -> dict | None:
        if not self.path or not os.path.exists(self.path):
            return None
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in Path(self.path).read_text().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def reference_arm(args) -> int:
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"impl": "reference", "unavailable": staging.BLOCKED_ONE_LINE}))
    return 0


def _vshard_section(gw, rank: int, world: int, local: int, warmup: int, steps: int) -> tuple[dict, int]:
    """The one collective, through the C-ABI: W-independence asserted on real bits, latency timed."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from base_b200 import vshards

    chains, n_stars, V = 1024, 10_000, vshards.DEFAULT_VSHARDS
    rng = np.random.default_rng(20260)               # every rank draws the same global values
    values = rng.normal(size=(chains, n_stars)) * 10.0 ** rng.integers(-6, 6, size=(chains, n_stars))
    lo, hi = vshards.local_star_range(rank, world, n_stars, V)
    launches = 0
    out: dict = {"chains": chains, "n_stars": n_stars, "n_vshards": V, "world": world}
    with vshards.PeerComm(local, rank, world, V, max_chains=chains) as comm:
        dv = torch.from_numpy(np.ascontiguousarray(values[:, lo:hi])).to(f"cuda:{local}")
        P = comm.shard_partials(dv, n_stars)
        total = comm.allreduce(P)
        launches += 2
        comm.status()                                # raises on a timeout
        # the world-1 answer, computed by this rank alone on its own GPU from ALL the stars
        alone = gw.vshard_total(values, V, device=local)["total"]
        launches += 2
        same = bool((total.cpu().numpy().view(np.int64) == alone.view(np.int64)).all())
        if not same:
            raise SystemExit(f"rank {rank}: the {world}-rank sum differs in bits from the 1-rank sum")
        out["bits_equal_world_1"] = True
        if world > 1:
            dist.barrier()
        lat = comm.latency(chains, warmup=max(warmup, 20), reps=max(steps, 200))
        launches += lat["launches"]
        us = torch.tensor([lat["us_stream"], lat["us_graph"]], device=f"cuda:{local}")
        if world > 1:
            # comparison lines (eager torch over NCCL, launched from Python): the all-gather of the
            # [V, chains] partials alone, and the same sum stated as all-gather + V ordered adds
            def timed(fn, n=50):
                for _ in range(max(warmup, 5)):
                    r_ = fn()
                dist.barrier(); torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(n):
                    r_ = fn()
                b.record()
                torch.cuda.synchronize()
                return r_, a.elapsed_time(b) * 1e3 / n
            flat = torch.empty(V * chains, dtype=torch.float64, device=P.device)
            _, us_gather = timed(lambda: dist.all_gather_into_tensor(flat, P.view(-1)))
            ref, us_sum = timed(lambda: vshards.allgather_ordered_sum(P))
            if not torch.equal(ref.view(torch.int64), total.view(torch.int64)):
                raise SystemExit("peer kernel and NCCL all-gather statement disagree in bits")
            us = torch.cat([us, torch.tensor([us_gather, us_sum], device=us.device)])
            dist.all_reduce(us, op=dist.ReduceOp.MAX)   # device-timed, max over ranks
            out["nccl_allgather_alone_us"] = round(us[2].item(), 2)
            out["nccl_allgather_plus_ordered_adds_us"] = round(us[3].item(), 2)
            dist.barrier()                              # nobody frees a mailbox a peer still writes
        out["peer_kernel_us_stream"] = round(us[0].item(), 2)
        out["peer_kernel_us_graph"] = round(us[1].item(), 2)
        comm.status()
        out["sharded_step"], n = _sharded_step(gw, comm, rank, world, local, warmup, steps)
        launches += n
    return out, launches


def _sharded_step(gw, comm, rank: int, world: int, local: int, warmup: int, steps: int) -> tuple[dict, int]:
    """Star sharding with real work per rank (still NOT a likelihood: the synthetic generator).
    A fixed job — 10 000 stars x 1024 terms x `chains` chains — is cut into the 64 virtual shards;
    each rank runs the fixed-order log-sum-exp of its own 64/W shards and then the cross-rank sum —
    as two launches, and as ONE kernel with the sum fused into the LSE kernel's tail (the warp that
    finishes a chain locally pushes its shards to the peers and pulls theirs).  Strong scaling: the job does not grow with W.  The total's bits are
    asserted equal, on every rank, to those the rank gets alone from all 64 shards."""
    import numpy as np
    import torch
    import torch.distributed as dist

    n_stars, cols, V = 10_000, 1_024, comm.n_vshards
    out, launches = {"n_stars": n_stars, "cols": cols, "scaling": "strong"}, 0
    for chains in (16, 128):
        r = comm.sharded_step(n_stars, cols, chains, warmup=max(warmup, 3), reps=steps)
        launches += r["launches"]
        alone = gw.lse_generated_shards(n_stars, cols, chains, V, 0, V, local)["total"]
        launches += 1
        if not (r["total"].view(np.int64) == alone.view(np.int64)).all():
            raise SystemExit(f"rank {rank}: the {world}-rank sharded step differs in bits from the 1-rank job")
        if not (r["total_fused"].view(np.int64) == alone.view(np.int64)).all():
            raise SystemExit(f"rank {rank}: the one-kernel step differs in bits from the 1-rank job")
        us = torch.tensor([r["us_step"], r["us_lse_alone"], r["us_fused_step"]], device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(us, op=dist.ReduceOp.MAX)   # device-timed, max over ranks
            dist.barrier()
        sec = us[2].item() * 1e-6
        out[f"chains_{chains}"] = {
            "us_fused_step_1_launch": round(us[2].item(), 2), "us_step_2_launches": round(us[0].item(), 2),
            "us_lse_share_alone": round(us[1].item(), 2),
            "gterms_per_s_whole_job_fused": round(n_stars * cols * chains / sec * 1e-9, 1),
            "bits_equal_world_1": True}
    return out, launches


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    warmup, steps = max(args.warmup, 0), max(args.steps, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    status = staging.probe()

    line = {
        "metric": _baseline_metric(), "value": None, "unit": "evals/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": None, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "none (blocked)",
        "config": {"workload": "BLOCKED — no BASE-9 likelihood exists to run; see `blocked`",
                   "step": "undefined: there is no likelihood evaluation to call a step; --steps/--warmup "
                           "only size the timing loops of the groundwork kernels below",
                   "l2": "n/a for the headline; lse rows_10000x1024 (82 MB) is under the 126 MB L2, "
                         "rows_40000x1024 (328 MB) is not, generated_* read no input at all"},
        "blocked": status.reason if status.blocked else
        "source now staged — SURVEY.md must be redone from it before a hot path exists: " + status.reason,
        "e2e": None, "roofline": None, "cpu_baseline": None,
        "gpu_launches": 0, "clocks": None, "groundwork": None,
    }
    if warmup < 3:
        line["warmup_note"] = "fewer than the 3 warm-up launches the timing rules ask for; as requested"

    from base_b200 import groundwork as gw  # raises if the .so is not built: no fallback
    if gw.device_count() == 0:
        line["error"] = "no CUDA device visible; groundwork not measured (there is no CPU fallback)"
        if rank == 0:
            print(json.dumps(line))
        return 1

    import torch
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()

    import numpy as np
    launches = 0
    few = max(3, steps // 4)
    with ClockSampler(local) as cs:
        r = gw.dfma_peak(local, ctas_per_sm=8, iters=1 << 16, warmup=warmup, reps=steps)
        launches += r["launches"]
        occ = {}
        for c in (1, 2, 4):
            o = gw.dfma_peak(local, ctas_per_sm=c, iters=1 << 16, warmup=warmup, reps=few)
            occ[str(c * 8)] = round(o["tflops"], 3)
            launches += o["launches"]
        occ["64"] = round(r["tflops"], 3)
        rates = {}
        for name, reps in (("exp", steps), ("log", steps), ("exp10", few), ("log10", few),
                           ("exp_spread", steps), ("log_spread", few)):
            t = gw.transcendental_rate(name, local, iters=1 << 12, warmup=warmup, reps=reps)
            rates[name] = round(t["gevals_per_s"], 2)
            launches += t["launches"]
        lat = gw.step_latency(local, warmup=50, reps=2000)
        launches += lat["launches"]
        # dependent-issue latency of the FP64 pipe: 2 chains in flight per scheduler
        l1 = gw.dfma_peak(local, ctas_per_sm=1, ilp=1, iters=1 << 16, warmup=warmup, reps=few)
        launches += l1["launches"]
        # what other instructions cost next to FP64 work: DFMA rate with k integer companions each
        company = {}
        for k, name in ((1, "1_imad"), (2, "2_imad"), (-2, "2_alu"), (-4, "4_alu")):
            c_ = gw.dfma_peak(local, ctas_per_sm=8, int_per_fma=k, iters=1 << 14, warmup=warmup, reps=few)
            company[name] = round(c_["tflops"] / r["tflops"], 3)
            launches += c_["launches"]
        cols = 1_024
        lse = {}
        # 10 000 rows = one proposal over a cfg2-sized cluster: latency-bound (one wave of work);
        # 160 000 rows = 16 proposals batched in one launch, which is what north_star prescribes
        for rows in (10_000, 160_000):
            g_ = gw.lse_generated(rows, cols, local, warmup=warmup, reps=steps)
            lse[f"generated_{rows}x{cols}"] = (rows, g_, False)
            launches += g_["launches"]
        rng = np.random.default_rng(1234)
        for rows in (10_000, 40_000):      # 82 MB sits in the 126 MB L2; 328 MB does not
            x = rng.normal(-40.0, 12.0, size=(rows, cols))
            s_ = gw.lse_rows(x, local, warmup=warmup, reps=steps)
            lse[f"rows_{rows}x{cols}"] = (rows, s_, True)
            launches += s_["launches"]
            del x
        vs, n = _vshard_section(gw, rank, world, local, warmup, steps)
        launches += n
    clocks = cs.summary()

    info = gw.device_info(local)
    mhz = info["sm_clock_mhz"]
    sm_mhz = clocks["sm_mhz"] if clocks else float(mhz)   # median SM clock sampled under load

    def lse_line(rows, res, reads_matrix):
        sec = res["ms_per_launch"] * 1e-3
        d = {"ms_per_launch": round(res["ms_per_launch"], 4),
             "gterms_per_s": round(rows * cols / sec * 1e-9, 2),
             "frac_of_exp_spread_rate": round(rows * cols / sec * 1e-9 / rates["exp_spread"], 3)}
        if reads_matrix:
            d["algorithmic_gb_per_s"] = round(rows * cols * 8 / sec * 1e-9, 1)
        else:
            # the kernel's own bound: its instruction mix (from the committed ncu capture of this
            # source) through the issue model, at the clock this run sampled
            k = roofline.LSE_STAGED_INSTR_PER_32_TERMS
            r_ = roofline.fp64_issue_roofline(rows * cols, sec, k["fp64"] / 32, k["other"] / 32,
                                              info["sm_count"], sm_mhz, "terms/s")
            d["frac_of_issue_bound"] = round(r_["frac"], 3)
            d["issue_bound_gterms_per_s"] = round(r_["peak"] * 1e-9, 1)
        return d

    g = {
        "note": "reference-independent denominators and plumbing; NOT the BASE-9 hot path",
        "fp64_dfma_tflops": round(r["tflops"], 3), "dfma_ms_per_launch": round(r["ms_per_launch"], 4),
        "dfma_tflops_by_warps_per_sm": occ,
        # 2 chains in flight per scheduler: warp-DFMAs per cycle per scheduler = 2 / latency
        "fp64_dependent_issue_latency_clk": round(
            2.0 / (l1["tflops"] * 1e12 / 2 / 32 / (info["sm_count"] * 4) / (mhz * 1e6)), 1),
        "dfma_rate_with_integer_company_per_dfma": company,
        "issue_model": "a DFMA holds a scheduler's issue port 2 cycles, any other instruction 1: "
                       "cycles ~ 2*n_fp64 + n_other (fits every ncu capture in profiles/r02_groundwork.md)",
        "fp64_gevals_per_s": rates,
        "rates_note": "exp/log/exp10/log10 are single-argument mid-range rates (contractions); "
                      "*_spread take a fresh log-uniform argument per evaluation (b9_groundwork.h)",
        "dependent_step_latency_us": {k: round(v, 2) for k, v in lat.items() if k.startswith("us_")},
        "lse": {name: lse_line(*v) for name, v in lse.items()},
        "vshard_sum": vs,
    }
    if dist is not None:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, g["fp64_dfma_tflops"])
        g["fp64_dfma_tflops_per_rank"] = per_rank
        dist.barrier()
        dist.destroy_process_group()

    line.update({"groundwork": g, "gpu_launches": launches, "clocks": clocks})
    broken = roofline.check_line(line)                   # the driver's contract, checked before it is
    if broken:
        line["contract_violations"] = broken
    if rank == 0:
        print(json.dumps(line))
    return 1 if broken else 0


if __name__ == "__main__":
    sys.exit(main())
