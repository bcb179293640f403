"""Times the two LSE kernels alone (CUDA events inside the C-ABI) — the quick loop for tuning.

    python tools/lse_probe.py [rows cols reps]
"""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from base_b200 import groundwork as gw  # noqa: E402

if os.environ.get("B9GW_LIB"):            # a variant build of the library, for tuning sweeps
    gw.LIB_PATH = Path(os.environ["B9GW_LIB"]).resolve()

rows, cols, reps = (int(a) for a in (sys.argv[1:4] + ["10000", "1024", "20"][len(sys.argv) - 1:]))
x = gw.generate_terms(rows, cols)
g = gw.lse_generated(rows, cols, warmup=3, reps=reps)
m = gw.lse_rows(x, warmup=3, reps=reps)
assert (g["row_lse"].view(np.int64) == m["row_lse"].view(np.int64)).all() and g["total"] == m["total"]
for name, r in (("lse_generated", g), ("lse_rows", m)):
    us = r["ms_per_launch"] * 1e3
    print(f"{name:14s} {rows}x{cols}: {us:8.2f} us/launch  {rows * cols / us * 1e-3:8.1f} G terms/s")
