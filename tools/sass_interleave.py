"""Static check of the LSE kernels' pass-2 schedule, read from the built library's SASS.

    python tools/sass_interleave.py [base_b200/libb9_groundwork.so]

Pass 2 of lse_staged_kernel exponentiates four terms per lane per iteration with a
branch-free copy of libm's exp fast path (csrc/common.cuh).  Each exp is a chain of ~17
dependent DFMAs at ~10 clk dependent-issue latency, so the four chains must be INTERLEAVED in
the instruction stream: with 6 warps per scheduler, one chain per warp keeps the FP64 pipe
~87 % fed, four keep it full (profiles/r02_groundwork.md "chains in flight").  The PTX is
interleaved as written; whether the SASS stays interleaved is ptxas's decision and changed
once without any change to the loop (a launch bound decided it, profiles/r02b_groundwork.md).
This prints, per kernel, the destination registers of the fast path's DFMAs as runs; a run
`Rn x11` is one exp executed on its own.
"""
from __future__ import annotations

import re
import subprocess
import sys
from pathlib import Path


def fast_path_runs(lib: str | Path, kernel_substr: str) -> list[tuple[str, int]]:
    """Runs of consecutive DFMAs with the same destination between the kernel's warp vote
    (the range test of pass 2) and the branch that ends the fast path."""
    sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    name, body = None, []
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name and kernel_substr in name:
                break
            name, body = m.group(1), []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m:
            body.append(m.group(1).strip())
    if not name or kernel_substr not in name:
        raise LookupError(f"no kernel matching {kernel_substr!r} in {lib}")
    # the fast path starts at the first DFMA that rounds with 1.5 * 2^52, just behind the vote
    first = next(j for j, t in enumerate(body) if t.startswith("DFMA") and "6.75539944105574400000e+15" in t)
    i = max(j for j in range(first) if body[j].startswith("VOTE.ALL"))
    j = next(j for j in range(first, len(body)) if body[j].startswith("BRA"))
    runs: list[tuple[str, int]] = []
    for t in body[i:j]:
        m = re.match(r"DFMA (R\d+)", t)
        if not m:
            continue
        if runs and runs[-1][0] == m.group(1):
            runs[-1] = (m.group(1), runs[-1][1] + 1)
        else:
            runs.append((m.group(1), 1))
    return runs


if __name__ == "__main__":
    lib = sys.argv[1] if len(sys.argv) > 1 else Path(__file__).resolve().parent.parent / "base_b200" / "libb9_groundwork.so"
    for k in ("lse_staged_kernelILi0ELb0", "lse_staged_kernelILi1ELb0", "lse_staged_kernelILi1ELb1"):
        runs = fast_path_runs(lib, k)
        print(k, "longest run", max(n for _, n in runs), "of", sum(n for _, n in runs), "DFMAs")
        print("  ", " ".join(f"{r}x{n}" if n > 1 else r for r, n in runs))
