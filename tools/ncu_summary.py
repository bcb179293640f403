#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the few columns the profiles/ notes quote.

    python tools/ncu_summary.py gpurun_out/r02_lse.ncu-rep > profiles/r02_ncu_full.csv

One row per profiled launch.  Needs `ncu` on PATH (reading a report needs no GPU).
"""
import csv
import subprocess
import sys

COLS = [
    ("kernel", "Kernel Name"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
    ("regs", "launch__registers_per_thread"), ("dyn_smem_B", "launch__shared_mem_per_block_dynamic"),
    ("waves_per_sm", "launch__waves_per_multiprocessor"),
    ("occ_limit_smem", "launch__occupancy_limit_shared_mem"), ("occ_limit_regs", "launch__occupancy_limit_registers"),
    ("duration_us", "gpu__time_duration.sum"),
    ("sm_cycles_active_avg", "sm__cycles_active.avg"), ("sm_cycles_elapsed_avg", "sm__cycles_elapsed.avg"),
    ("warps_active_per_sched", "smsp__warps_active.avg.per_cycle_active"),
    ("issue_active_per_cycle", "smsp__issue_active.avg.per_cycle_active"),
    ("fp64_pipe_pct_of_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    ("fp64_pipe_pct_of_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("warp_inst_executed", "smsp__inst_executed.sum"),
    ("fp64_inst_pct_of_peak_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_pct_of_peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("per_issue_active.ratio")]
    w = csv.writer(sys.stdout)
    w.writerow([c for c, _ in COLS] + ["top_stalls(per issue)"])
    for r in data:
        vals = []
        for c, h in COLS:
            v = r[hdr.index(h)] if h in hdr else ""
            if c == "kernel":
                v = v.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            elif c in ("dram_read_MB", "dram_write_MB") and v:
                u = units[hdr.index(h)]
                scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                v = f"{float(v) * scale:.3f}"
            vals.append(v)
        st = sorted(((float(r[hdr.index(h)] or 0), h.split("stalled_")[1].split("_per_issue")[0]) for h in stall_cols),
                    reverse=True)[:4]
        vals.append(" ".join(f"{n}={v:.2f}" for v, n in st))
        w.writerow(vals)


if __name__ == "__main__":
    main(sys.argv[1])
