"""Roofline arithmetic and the bench line's contract, as code.

Reference-independent (the BASE-9 hot path is BLOCKED).  Two things live here so
that unblock day only has to plug a kernel in:

1. The FP64 issue model measured on B200 (profiles/r02_groundwork.md §1): a scheduler's issue
   port is held 2 cycles by a DFMA/DADD/DMUL and 1 cycle by anything else, so a kernel whose
   warps execute n_fp64 FP64 and n_other other warp-instructions per scheduler needs at least
   2*n_fp64 + n_other cycles.  tests/test_issue_model.py holds this against every
   `ncu --set full` capture committed under profiles/: throughput-bound launches sit within
   4 % of it, no launch beats it by more than that.
2. `check_line`: what the driver's contract asks of bench.py's JSON line, for the blocked
   line printed today and for the measured line that will replace it.

Pure Python, no GPU, no torch.
"""
from __future__ import annotations

import math

SCHEDULERS_PER_SM = 4
FP64_ISSUE_CYCLES = 2          # issue-port cycles of one FP64 warp-instruction
OTHER_ISSUE_CYCLES = 1

# Warp-instructions lse_staged_kernel<1, false> executes per 32 terms (one warp-wide column
# step) at 1024 columns, from the committed capture of this source (profiles/r02b_ncu_full.csv,
# 160 000 x 1024: 392 272 per scheduler, 197 346 of them FP64, over 160000*1024/32/592 steps).
LSE_STAGED_INSTR_PER_32_TERMS = {"fp64": 22.82, "other": 22.54}


def issue_bound_cycles(n_fp64: float, n_other: float) -> float:
    """Least cycles a scheduler needs to issue n_fp64 FP64 and n_other other warp-instructions."""
    if n_fp64 < 0 or n_other < 0:
        raise ValueError("instruction counts cannot be negative")
    return FP64_ISSUE_CYCLES * n_fp64 + OTHER_ISSUE_CYCLES * n_other


def fp64_issue_roofline(units: float, seconds: float, fp64_per_unit: float, other_per_unit: float,
                        sm_count: int, sm_mhz: float, unit: str, traffic=None) -> dict:
    """`roofline` object for an FP64 issue-bound kernel.

    units            what one launch processes (terms, star x grid points ...)
    *_per_unit       warp-instructions per unit PER SCHEDULER-VISIBLE WARP, i.e. counted per warp
                     and divided by the units that warp-instruction covers (read off the SASS or an
                     ncu capture)
    peak             units/s if every scheduler issued back to back at sm_mhz
    """
    if units <= 0 or seconds <= 0 or sm_count <= 0 or sm_mhz <= 0:
        raise ValueError("units, seconds, sm_count and sm_mhz must be positive")
    cycles_per_unit = issue_bound_cycles(fp64_per_unit, other_per_unit)
    if cycles_per_unit <= 0:
        raise ValueError("a unit of work must cost at least one instruction")
    peak = sm_count * SCHEDULERS_PER_SM * sm_mhz * 1e6 / cycles_per_unit
    achieved = units / seconds
    return {"bound": "fp64-issue", "achieved": achieved, "peak": peak, "unit": unit,
            "frac": achieved / peak, "traffic": traffic,
            "fp64_pipe_frac_at_bound": FP64_ISSUE_CYCLES * fp64_per_unit / cycles_per_unit}


def hbm_roofline(algorithmic_bytes: float, seconds: float, peak_gbs: float, traffic=None) -> dict:
    """`roofline` object for an HBM-bound kernel; peak_gbs from MEASURED_PEAKS.json."""
    if algorithmic_bytes <= 0 or seconds <= 0 or peak_gbs <= 0:
        raise ValueError("bytes, seconds and peak must be positive")
    achieved = algorithmic_bytes / seconds * 1e-9
    return {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
            "frac": achieved / peak_gbs, "traffic": traffic}


# ------------------------------------------------------------------ what batching buys

# One launch of the stand-in kernel (fixed-order log-sum-exp, 1024 terms per row, terms produced on
# chip) against the rows it covers: profiles/r02b_lse_sizes.log.  NOT a likelihood; it stands for
# "a kernel that does ~1000 grid points of FP64 work per star" when sizing batches.
LSE_LAUNCH_US = [(0, 8.4), (1250, 12.09), (2500, 14.49), (5000, 18.65), (10_000, 29.11),
                 (20_000, 47.77), (40_000, 85.29), (160_000, 304.31)]
HOST_ROUND_TRIP_US = 14.2      # launch + 8-byte D2H + sync of one dependent step (profiles/r02_groundwork.md)


def launch_us(rows: float) -> float:
    """Piecewise-linear read of LSE_LAUNCH_US; beyond the table the last slope continues."""
    if rows < 0:
        raise ValueError("rows cannot be negative")
    t = LSE_LAUNCH_US
    for (r0, u0), (r1, u1) in zip(t, t[1:]):
        if rows <= r1:
            return u0 + (u1 - u0) * (rows - r0) / (r1 - r0)
    (r0, u0), (r1, u1) = t[-2], t[-1]
    return u1 + (u1 - u0) / (r1 - r0) * (rows - r1)


def speculative_steps_per_s(stars: int, depth: int, acceptance: float, chains: int = 1,
                            host_us: float = HOST_ROUND_TRIP_US) -> float:
    """Chain-steps per second of include/b9_spec_chain.hpp's rounds on one GPU, if a proposal cost
    what `stars` rows of the stand-in kernel cost: a round evaluates depth x chains candidates in
    one launch plus one host round trip, and completes (1 - (1-a)^depth) / a steps per chain."""
    if stars < 1 or depth < 1 or chains < 1 or not 0.0 < acceptance <= 1.0:
        raise ValueError("need stars, depth, chains >= 1 and 0 < acceptance <= 1")
    steps = chains * (1.0 - (1.0 - acceptance) ** depth) / acceptance
    return steps / (launch_us(stars * depth * chains) + host_us) * 1e6


def best_depth(stars: int, acceptance: float, chains: int = 1, max_depth: int = 64) -> int:
    return max(range(1, max_depth + 1), key=lambda k: speculative_steps_per_s(stars, k, acceptance, chains))


# ------------------------------------------------------------------ the bench line

_NUM = (int, float)
_BOUNDS = {"hbm", "tensor", "fp64-issue"}
_SLOWDOWNS = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def _num(x) -> bool:
    return isinstance(x, _NUM) and not isinstance(x, bool) and math.isfinite(x)


def check_line(line: dict) -> list[str]:
    """Every way `line` falls short of the driver's contract for bench.py ([] = none).

    A blocked line (value null) must say why and must not carry any performance object; a
    measured line must carry all of them, consistent with one another.
    """
    bad: list[str] = []

    def need(key, ok, what):
        if key not in line:
            bad.append(f"missing key {key!r}")
        elif not ok(line[key]):
            bad.append(f"{key!r} must be {what}, got {line[key]!r}")

    need("metric", lambda v: isinstance(v, str) and v, "a non-empty string")
    need("unit", lambda v: isinstance(v, str) and v, "a non-empty string")
    need("n_gpus", lambda v: isinstance(v, int) and v >= 1, "an int >= 1")
    need("steps", lambda v: isinstance(v, int) and v >= 1, "an int >= 1")
    need("warmup", lambda v: isinstance(v, int) and v >= 0, "an int >= 0")
    need("higher_is_better", lambda v: isinstance(v, bool), "a bool")
    need("scaling", lambda v: v in ("weak", "strong"), "'weak' or 'strong'")
    need("dtype", lambda v: isinstance(v, str) and v, "a non-empty string")
    need("data", lambda v: isinstance(v, str) and v, "a non-empty string")
    need("config", lambda v: isinstance(v, dict) and isinstance(v.get("workload"), str) and "model" not in v,
         "a dict naming the workload (no model keys)")
    need("vs_baseline", lambda v: v is None or _num(v), "null or a number")
    need("gpu_launches", lambda v: isinstance(v, int) and v >= 0, "an int")
    if bad:
        return bad

    if line.get("value") is None:
        if not (isinstance(line.get("blocked"), str) and line["blocked"]):
            bad.append("a line without a value must say why in 'blocked'")
        for k in ("ms_per_step", "e2e", "roofline", "cpu_baseline"):
            if line.get(k) is not None:
                bad.append(f"{k!r} must be null on a blocked line")
        return bad

    if not _num(line["value"]) or line["value"] <= 0:
        bad.append("'value' must be a positive number")
    if not _num(line.get("ms_per_step")) or line["ms_per_step"] <= 0:
        bad.append("'ms_per_step' must be a positive number")
    if line["warmup"] < 3:
        bad.append("timing rules ask for >= 3 warm-up steps")
    if line["gpu_launches"] <= 0:
        bad.append("'gpu_launches' must count the kernels launched in the timed region")
    if "blocked" in line and line["blocked"]:
        bad.append("a measured line cannot also be 'blocked'")

    e = line.get("e2e")
    if not isinstance(e, dict):
        bad.append("'e2e' must be an object")
    else:
        if not _num(e.get("value")) or e["value"] <= 0:
            bad.append("e2e.value must be a positive number")
        if e.get("unit") != line["unit"]:
            bad.append("e2e.unit must be the line's unit")
        for k in ("h2d_bytes_per_step", "d2h_bytes_per_step"):
            if not (isinstance(e.get(k), int) and e[k] > 0):
                bad.append(f"e2e.{k} must be a positive int: host buffers cross the bus inside the timed region")
        if _num(e.get("value")) and _num(line["value"]) and e["value"] > line["value"] * 1.001:
            bad.append("e2e.value cannot exceed the device-resident value")

    r = line.get("roofline")
    if not isinstance(r, dict):
        bad.append("'roofline' must be an object")
    else:
        if r.get("bound") not in _BOUNDS:
            bad.append(f"roofline.bound must be one of {sorted(_BOUNDS)}")
        for k in ("achieved", "peak", "frac"):
            if not _num(r.get(k)) or r[k] <= 0:
                bad.append(f"roofline.{k} must be a positive number")
        if not isinstance(r.get("unit"), str):
            bad.append("roofline.unit must be a string")
        if "traffic" not in r or not (r["traffic"] is None or _num(r["traffic"])):
            bad.append("roofline.traffic must be null or dram bytes per launch")
        if all(_num(r.get(k)) and r[k] > 0 for k in ("achieved", "peak", "frac")):
            if abs(r["frac"] - r["achieved"] / r["peak"]) > 1e-3 * max(1.0, r["frac"]):
                bad.append("roofline.frac must be achieved / peak")
            if r["frac"] > 1.05:
                bad.append("roofline.frac above 1: the peak or the algorithmic work is wrong")

    c = line.get("cpu_baseline")
    if not isinstance(c, dict):
        bad.append("'cpu_baseline' must be an object")
    else:
        if not _num(c.get("value")) or c["value"] <= 0:
            bad.append("cpu_baseline.value must be a positive number")
        if not (isinstance(c.get("cores"), int) and c["cores"] >= 1):
            bad.append("cpu_baseline.cores must be the threads actually used")
        if c.get("kind") not in ("reference", "port"):
            bad.append("cpu_baseline.kind must be 'reference' or 'port'")
        if not (isinstance(c.get("sample"), str) and c["sample"]):
            bad.append("cpu_baseline.sample must say what was timed")
        if not isinstance(c.get("unit"), str):
            bad.append("cpu_baseline.unit must be a string")

    k = line.get("clocks")
    if not isinstance(k, dict):
        bad.append("'clocks' must be sampled during the timed region")
    else:
        if not _num(k.get("sm_mhz")) or not _num(k.get("sm_max_mhz")):
            bad.append("clocks.sm_mhz and clocks.sm_max_mhz must be numbers")
        if not isinstance(k.get("reasons"), list):
            bad.append("clocks.reasons must be a list")
        elif _SLOWDOWNS & set(k["reasons"]):
            bad.append("the run saw a slowdown reason and must be re-measured")
    return bad
