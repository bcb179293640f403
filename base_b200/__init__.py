"""base_b200 — B200-native groundwork for BASE-9's cluster log-likelihood.

STATUS: the hot path is BLOCKED.  The mounted reference is a
relocation notice (/root/reference/README.md:1-4); base-cpp is not staged and
BASELINE.json's north_star forbids reconstructing it from memory.  This package
therefore holds only (i) the staging gate and (ii) reference-independent FP64
measurements the path will need on day one: `staging`, `groundwork`, `build`.
"""
from . import staging  # noqa: F401

__all__ = ["staging", "groundwork", "build"]
