"""The FP64 issue model (base_b200/roofline.py) against the ncu captures committed under profiles/,
and the bench-line contract checker against good and bad lines.  CPU only: reads CSV files."""
import copy
import csv
import json
from pathlib import Path

import pytest

from base_b200 import roofline as rf

ROOT = Path(__file__).resolve().parent.parent
SCHEDULERS = 148 * rf.SCHEDULERS_PER_SM
CAPTURES = ["r02_ncu_full.csv", "r02_ncu_lse_variants.csv", "r02b_ncu_full.csv", "r02b_ncu_fused.csv"]


def launches():
    """(file, kernel, waves, predicted cycles, measured active cycles) per captured launch."""
    out = []
    for name in CAPTURES:
        for r in csv.DictReader(open(ROOT / "profiles" / name)):
            active = float(r["sm_cycles_active_avg"])
            total = float(r["warp_inst_executed"]) / SCHEDULERS
            # FP64 warp-instructions per scheduler: the pipe's peak is one every 2 cycles
            pct = float(r.get("fp64_inst_pct_of_peak_active") or r["fp64_pipe_pct_of_active"])
            n_fp64 = pct / 100.0 * active / rf.FP64_ISSUE_CYCLES
            out.append((name, r["kernel"], float(r["waves_per_sm"]),
                        rf.issue_bound_cycles(n_fp64, total - n_fp64), active))
    return out


def test_captures_are_there():
    got = launches()
    assert len(got) >= 40 and {n for n, *_ in got} == set(CAPTURES)


def test_throughput_bound_launches_sit_on_the_issue_bound():
    """>= 10 waves of CTAs per SM: nothing but issue slots limits the launch, and the model says how many."""
    full = [l for l in launches() if l[2] >= 10.0]
    assert len(full) >= 8
    for name, kernel, waves, predicted, active in full:
        assert abs(active / predicted - 1.0) <= 0.04, (name, kernel, waves, predicted, active)


def test_no_launch_beats_the_issue_bound():
    for name, kernel, waves, predicted, active in launches():
        assert active / predicted >= 0.96, (name, kernel, waves, predicted, active)


def test_latency_bound_launches_are_above_it():
    """A single wave (10 000 x 1024 and smaller) cannot reach the bound: that is what batching is for."""
    small = [l for l in launches() if l[2] < 3.0 and "lse" in l[1]]
    assert small and all(active / predicted > 1.05 for *_, predicted, active in small)


def test_lse_constants_match_the_capture_they_cite():
    rows = [r for r in csv.DictReader(open(ROOT / "profiles" / "r02b_ncu_full.csv"))
            if r["kernel"].startswith("lse_staged_kernel<1, 0>") and r["grid"] == "10000"]
    assert rows
    steps = 160_000 * 1024 / 32 / SCHEDULERS
    total = float(rows[0]["warp_inst_executed"]) / SCHEDULERS / steps
    fp64 = float(rows[0]["fp64_inst_pct_of_peak_active"]) / 100 * float(rows[0]["sm_cycles_active_avg"]) / 2 / steps
    assert fp64 == pytest.approx(rf.LSE_STAGED_INSTR_PER_32_TERMS["fp64"], abs=0.02)
    assert total - fp64 == pytest.approx(rf.LSE_STAGED_INSTR_PER_32_TERMS["other"], abs=0.02)


def test_fp64_issue_roofline_reproduces_the_quoted_fraction():
    """profiles/r02b_bench_n1.json: 160 000 x 1024 generated terms in 304.5 us at 1965 MHz."""
    g = json.loads((ROOT / "profiles" / "r02b_bench_n1.json").read_text().strip().splitlines()[-1])["groundwork"]
    ms = g["lse"]["generated_160000x1024"]["ms_per_launch"]
    k = rf.LSE_STAGED_INSTR_PER_32_TERMS
    r = rf.fp64_issue_roofline(160_000 * 1024, ms * 1e-3, k["fp64"] / 32, k["other"] / 32, 148, 1965.0, "terms/s")
    assert r["bound"] == "fp64-issue" and r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    assert 0.97 <= r["frac"] <= 1.0                      # "0.985 of the issue bound"
    assert 0.66 <= r["fp64_pipe_frac_at_bound"] <= 0.68  # what sm__pipe_fp64_cycles_active can show at best
    # a pure DFMA stream: the bound is the DFMA peak, 148 x 64 FMA/clk
    d = rf.fp64_issue_roofline(1e12, 1.0, 1 / 32, 0.0, 148, 1965.0, "fma/s")
    assert d["peak"] == pytest.approx(148 * 64 * 1.965e9)


def test_rooflines_reject_nonsense():
    with pytest.raises(ValueError):
        rf.issue_bound_cycles(-1, 0)
    with pytest.raises(ValueError):
        rf.fp64_issue_roofline(0, 1, 1, 1, 148, 1965, "x")
    with pytest.raises(ValueError):
        rf.fp64_issue_roofline(1, 1, 0, 0, 148, 1965, "x")
    with pytest.raises(ValueError):
        rf.hbm_roofline(1, 0, 6552.6)
    h = rf.hbm_roofline(81.92e6, 33.9e-6, 6552.6, traffic=81.97e6)
    assert h["bound"] == "hbm" and h["frac"] == pytest.approx(2416.5 / 6552.6, rel=1e-3)


def test_what_batching_buys_on_the_measured_launch_curve():
    """Speculative depth pays where one proposal is a latency-bound launch, not on a 10 000-star
    cluster on one GPU; independent chains pay until the launch reaches the throughput bound."""
    assert rf.launch_us(10_000) == pytest.approx(29.11) and rf.launch_us(160_000) == pytest.approx(304.31)
    assert rf.launch_us(7_500) == pytest.approx((18.65 + 29.11) / 2)
    a = 0.25
    one = {s: rf.speculative_steps_per_s(s, 1, a) for s in (100, 1_250, 10_000)}
    assert one[10_000] == pytest.approx(1e6 / (29.11 + 14.2))
    # 10 000 stars: the launch is already most of a wave; depth buys < 25 %, and deep speculation loses
    assert rf.speculative_steps_per_s(10_000, rf.best_depth(10_000, a), a) < 1.25 * one[10_000]
    assert rf.speculative_steps_per_s(10_000, 16, a) < one[10_000]
    assert rf.best_depth(10_000, a) <= 3
    # 1 250 stars (a 10 000-star cluster sharded over 8 GPUs): about 2.3x at depth 6
    assert 2.0 < rf.speculative_steps_per_s(1_250, rf.best_depth(1_250, a), a) / one[1_250] < 2.6
    assert 4 <= rf.best_depth(1_250, a) <= 8
    # 100 stars (cfg1): the launch is all latency, depth 16 gives > 3x
    assert rf.speculative_steps_per_s(100, 16, a) > 3.0 * one[100] and rf.best_depth(100, a) >= 8
    # independent chains: 64 chains of 100 stars in one launch, > 30x one chain
    assert rf.speculative_steps_per_s(100, 1, a, chains=64) > 30 * one[100]
    with pytest.raises(ValueError):
        rf.speculative_steps_per_s(100, 0, a)


# ------------------------------------------------------------------ the bench line

BLOCKED = {
    "metric": "m", "value": None, "unit": "evals/s", "n_gpus": 1, "steps": 20, "warmup": 3, "ms_per_step": None,
    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "none (blocked)",
    "config": {"workload": "BLOCKED"}, "blocked": "why", "e2e": None, "roofline": None, "cpu_baseline": None,
    "gpu_launches": 7040, "clocks": None,
}
MEASURED = {
    **BLOCKED, "value": 1.0e6, "ms_per_step": 1.0, "data": "synthetic", "config": {"workload": "cfg2"},
    "e2e": {"value": 0.9e6, "unit": "evals/s", "h2d_bytes_per_step": 4096, "d2h_bytes_per_step": 1024},
    "roofline": {"bound": "fp64-issue", "achieved": 5.0, "peak": 10.0, "unit": "G/s", "frac": 0.5, "traffic": None},
    "cpu_baseline": {"value": 1.0e3, "unit": "evals/s", "cores": 1, "kind": "reference", "sample": "200 steps"},
    "clocks": {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": []},
}
del MEASURED["blocked"]


def test_good_lines_pass():
    assert rf.check_line(BLOCKED) == []
    assert rf.check_line(MEASURED) == []


@pytest.mark.parametrize("edit,needle", [
    (lambda l: l.pop("blocked"), "must say why"),
    (lambda l: l.update(roofline={"bound": "hbm"}), "'roofline' must be null"),
    (lambda l: l.update(scaling="none"), "'weak' or 'strong'"),
    (lambda l: l["config"].update(model="x"), "no model keys"),
    (lambda l: l.pop("gpu_launches"), "missing key 'gpu_launches'"),
])
def test_bad_blocked_lines_are_named(edit, needle):
    line = copy.deepcopy(BLOCKED)
    edit(line)
    assert any(needle in b for b in rf.check_line(line)), rf.check_line(line)


@pytest.mark.parametrize("edit,needle", [
    (lambda l: l.update(warmup=1), "3 warm-up"),
    (lambda l: l.update(gpu_launches=0), "gpu_launches"),
    (lambda l: l["e2e"].update(h2d_bytes_per_step=0), "h2d_bytes_per_step"),
    (lambda l: l["e2e"].update(value=2.0e6), "cannot exceed"),
    (lambda l: l["roofline"].update(frac=0.7), "achieved / peak"),
    (lambda l: l["roofline"].update(achieved=20.0, frac=2.0), "above 1"),
    (lambda l: l["roofline"].pop("traffic"), "traffic"),
    (lambda l: l["cpu_baseline"].update(kind="guess"), "'reference' or 'port'"),
    (lambda l: l["cpu_baseline"].update(cores=0), "cores"),
    (lambda l: l["clocks"].update(reasons=["hw_thermal_slowdown"]), "re-measured"),
    (lambda l: l.update(blocked="still"), "cannot also be"),
    (lambda l: l.update(cpu_baseline=None), "'cpu_baseline' must be an object"),
])
def test_bad_measured_lines_are_named(edit, needle):
    line = copy.deepcopy(MEASURED)
    edit(line)
    assert any(needle in b for b in rf.check_line(line)), rf.check_line(line)


def test_the_committed_bench_lines_satisfy_the_contract():
    for name in ("r02b_bench_n1.json", "r02b_bench_n2.json", "r02b_bench_n4.json", "r02b_bench_n8.json"):
        line = json.loads((ROOT / "profiles" / name).read_text().strip().splitlines()[-1])
        assert rf.check_line(line) == [], name
