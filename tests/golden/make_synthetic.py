"""Writes tests/golden/synthetic.b9dump — a format fixture, NOT reference data.

The values are arithmetic on the record key (no model of anything): they exist so that the
loader, the C writer and the comparator are tested against a committed file whose exact
bits are known.  Real golden vectors can only come from the staged reference
(tools/unblock.sh); this directory holds none yet — parity is UNPINNED.

    python tests/golden/make_synthetic.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from tests.golden_io import Record, dump  # noqa: E402


def records():
    out = []
    for stage, n in (("stageA", 5), ("stageB", 3), ("total", 1)):
        for star in ([-1] if stage == "total" else range(4)):
            k = np.arange(n, dtype=np.float64)
            v = np.ldexp(1.0 + (k + 1) / 7.0, int(star) * 3 - 5) * (-1.0) ** k      # spread of exponents and signs
            out.append(Record(stage, star, v))
    out.append(Record("edge", 0, np.array([0.0, -0.0, 5e-324, -1.7976931348623157e308, np.inf, -np.inf, np.nan])))
    out.append(Record("empty", 1, np.array([])))
    return out


if __name__ == "__main__":
    dump(ROOT / "tests" / "golden" / "synthetic.b9dump", records(),
         meta={"source": "tests/golden/make_synthetic.py (synthetic; not reference data)", "tolerance": "n/a"})
