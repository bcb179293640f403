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
dependent step, a fixed-order row log-sum-exp, and at N>1 the latency of the order-fixed cross-rank sum of 1024 per-chain scalars.
A "step" here is one launch of the DFMA kernel; K steps are timed with CUDA
events on the launching stream after W warm-up launches.  `gpu_launches` counts
those groundwork kernels and nothing else.

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

from base_b200 import staging  # noqa: E402


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


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    status = staging.probe()

    line = {
        "metric": _baseline_metric(), "value": None, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": None, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "none (blocked)",
        "config": {"workload": "BLOCKED — no BASE-9 likelihood exists to run; see `blocked`",
                   "l2": "n/a (groundwork kernels are register-resident; lse input < L2 noted below)"},
        "blocked": status.reason if status.blocked else
        "source now staged — SURVEY.md must be redone from it before a hot path exists: " + status.reason,
        "e2e": None, "roofline": None, "cpu_baseline": None,
        "gpu_launches": 0, "clocks": None, "groundwork": None,
    }

    from base_b200 import groundwork as gw  # raises if the .so is not built: no fallback
    if gw.device_count() == 0:
        line["error"] = "no CUDA device visible; groundwork not measured (there is no CPU fallback)"
        if rank == 0:
            print(json.dumps(line))
        return 1

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()

    import numpy as np
    launches = 0
    with ClockSampler(local) as cs:
        # step = one DFMA launch: 148*8 CTAs x 256 thr x 8 chains x 65536 fma
        r = gw.dfma_peak(local, ctas_per_sm=8, iters=1 << 16, warmup=warmup, reps=args.steps)
        launches += r["launches"]
        occ = {}
        for c in (1, 2, 4):
            o = gw.dfma_peak(local, ctas_per_sm=c, iters=1 << 16, warmup=warmup, reps=max(3, args.steps // 4))
            occ[str(c * 8)] = round(o["tflops"], 3)
            launches += o["launches"]
        occ["64"] = round(r["tflops"], 3)
        e = gw.transcendental_rate("exp", local, iters=1 << 12, warmup=warmup, reps=args.steps)
        l = gw.transcendental_rate("log", local, iters=1 << 12, warmup=warmup, reps=args.steps)
        e10 = gw.transcendental_rate("exp10", local, iters=1 << 12, warmup=warmup, reps=max(3, args.steps // 4))
        l10 = gw.transcendental_rate("log10", local, iters=1 << 12, warmup=warmup, reps=max(3, args.steps // 4))
        lat = gw.step_latency(local, warmup=50, reps=2000)
        launches += e["launches"] + l["launches"] + e10["launches"] + l10["launches"] + lat["launches"]
        rng = np.random.default_rng(1234)
        rows, cols = 10_000, 1_024  # 82 MB: under the 126 MB L2, second pass is an L2 hit
        x = rng.normal(-40.0, 12.0, size=(rows, cols))
        s = gw.lse_rows(x, local, warmup=warmup, reps=args.steps)
        launches += s["launches"]
    clocks = cs.summary()

    g = {
        "note": "reference-independent denominators; NOT the BASE-9 hot path",
        "fp64_dfma_tflops": round(r["tflops"], 3), "dfma_ms_per_launch": round(r["ms_per_launch"], 4),
        "dfma_tflops_by_warps_per_sm": occ,
        "fp64_exp_gevals_per_s": round(e["gevals_per_s"], 2),
        "fp64_log_gevals_per_s": round(l["gevals_per_s"], 2),
        "fp64_exp10_gevals_per_s": round(e10["gevals_per_s"], 2),
        "fp64_log10_gevals_per_s": round(l10["gevals_per_s"], 2),
        "dependent_step_latency_us": {k: round(v, 2) for k, v in lat.items() if k.startswith("us_")},
        "lse_rows": {"rows": rows, "cols": cols, "ms_per_launch": round(s["ms_per_launch"], 4),
                     "gelem_per_s": round(rows * cols / (s["ms_per_launch"] * 1e-3) * 1e-9, 2),
                     "algorithmic_gb_per_s": round(rows * cols * 8 / (s["ms_per_launch"] * 1e-3) * 1e-9, 1)},
    }

    if dist is not None:
        import torch
        from base_b200.chain_reduce import ordered_allreduce_sum
        t = torch.arange(1024, dtype=torch.float64, device=f"cuda:{local}") * (rank + 1)
        for _ in range(warmup):
            ordered_allreduce_sum(t)
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(200):
            out = ordered_allreduce_sum(t)
        b.record()
        torch.cuda.synchronize(); dist.barrier()
        us = torch.tensor([a.elapsed_time(b) * 1e3 / 200], device=f"cuda:{local}")
        dist.all_reduce(us, op=dist.ReduceOp.MAX)
        expect = torch.arange(1024, dtype=torch.float64) * (world * (world + 1) // 2)
        if not torch.equal(out.cpu(), expect):
            raise SystemExit("ordered_allreduce_sum returned a wrong sum")
        per_rank = [None] * world
        dist.all_gather_object(per_rank, g["fp64_dfma_tflops"])
        g["fp64_dfma_tflops_per_rank"] = per_rank
        g["ordered_allreduce_1024xf64_us_max_over_ranks"] = round(us.item(), 2)
        dist.barrier()
        dist.destroy_process_group()

    line.update({"groundwork": g, "gpu_launches": launches, "clocks": clocks})
    if rank == 0:
        print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
