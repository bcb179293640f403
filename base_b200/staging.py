"""The gate BASELINE.json's north_star puts in front of all other work.

north_star: "Before any other work, the engineer must confirm that the base-cpp
(BASE-9) source and its model tables are staged offline; if they are not, report
BLOCKED in BASELINE.md and do not reconstruct a reference from memory."

`probe()` answers that question mechanically, so that tests, `bench.py` and
`__graft_entry__` all report the same status and flip together the day the
source is staged.  It looks only at the two places SURVEY.md names for staging:
the reference mount and the git-ignored `baseline/_ref/` slot.  It never reads a
file's content on the GPU box path (`/root/reference` does not exist there); the
absence of both roots is simply "not staged".

Criteria (all three are needed to unblock, per SURVEY.md "What must be staged"):
  source  — at least one C/C++ translation unit or header under the root;
  build   — a build file (CMakeLists.txt / Makefile / configure.ac / meson.build);
  tables  — at least one directory of model tables: any file whose path
            mentions one of the model families north_star names.
Nothing here encodes knowledge of base-cpp's real layout, because none has been
read; the criteria are the weakest ones that distinguish "a source tree with
data" from "/root/reference/README.md:1-4".
"""
from __future__ import annotations

import dataclasses
import json
import os
import sys
from pathlib import Path

REPO_ROOT = Path(__file__).resolve().parent.parent
CANDIDATE_ROOTS = (Path("/root/reference"), REPO_ROOT / "baseline" / "_ref")

_SOURCE_EXT = {".c", ".cc", ".cpp", ".cxx", ".h", ".hh", ".hpp", ".hxx"}
_BUILD_FILES = {"CMakeLists.txt", "Makefile", "makefile", "GNUmakefile",
                "configure.ac", "configure", "meson.build"}
# Model families exactly as BASELINE.json's north_star spells them.
_TABLE_FAMILIES = ("dsed", "parsec", "yale", "girardi", "montgomery", "renedo",
                   "bergeron", "althaus")
_MAX_FILES = 200_000  # bound the walk; a staged tree is far smaller


@dataclasses.dataclass(frozen=True)
class RootReport:
    root: str
    exists: bool
    n_files: int
    n_source: int
    build_files: tuple[str, ...]
    table_hits: tuple[str, ...]

    @property
    def staged(self) -> bool:
        return self.n_source > 0 and bool(self.build_files) and bool(self.table_hits)


@dataclasses.dataclass(frozen=True)
class StagingStatus:
    blocked: bool
    reason: str
    roots: tuple[RootReport, ...]

    def to_json(self) -> str:
        return json.dumps(dataclasses.asdict(self), indent=1)


def _scan(root: Path) -> RootReport:
    if not root.is_dir():
        return RootReport(str(root), False, 0, 0, (), ())
    n_files = n_source = 0
    builds: list[str] = []
    tables: list[str] = []
    for dirpath, dirnames, filenames in os.walk(root):
        dirnames[:] = [d for d in dirnames if d != ".git"]
        for name in filenames:
            n_files += 1
            if n_files > _MAX_FILES:
                break
            rel = os.path.relpath(os.path.join(dirpath, name), root)
            ext = os.path.splitext(name)[1].lower()
            if ext in _SOURCE_EXT:
                n_source += 1
            elif name in _BUILD_FILES and len(builds) < 8:
                builds.append(rel)
            low = rel.lower()
            if ext not in _SOURCE_EXT and len(tables) < 8 and any(f in low for f in _TABLE_FAMILIES):
                tables.append(rel)
        if n_files > _MAX_FILES:
            break
    return RootReport(str(root), True, n_files, n_source, tuple(builds), tuple(tables))


def probe(roots=CANDIDATE_ROOTS) -> StagingStatus:
    reports = tuple(_scan(Path(r)) for r in roots)
    if any(r.staged for r in reports):
        where = next(r.root for r in reports if r.staged)
        return StagingStatus(
            False,
            f"a source tree with build files and model tables is staged at {where}: "
            "redo SURVEY.md from source (checklists in its sections 1-8), then build "
            "the oracle before any kernel",
            reports)
    parts = []
    for r in reports:
        if not r.exists:
            parts.append(f"{r.root}: absent")
        else:
            missing = [w for w, ok in (("source", r.n_source > 0), ("build files", bool(r.build_files)),
                                       ("model tables", bool(r.table_hits))) if not ok]
            parts.append(f"{r.root}: {r.n_files} file(s), missing {', '.join(missing)}")
    return StagingStatus(
        True,
        "BLOCKED: base-cpp (BASE-9) source and model tables are not staged offline ("
        + "; ".join(parts) + "); north_star forbids reconstructing the reference from memory",
        reports)


BLOCKED_ONE_LINE = ("reference is a relocation stub (/root/reference/README.md:1-4); base-cpp "
                    "source and model tables are not staged offline and there is no network")


def main(argv=None) -> int:
    st = probe()
    print(st.to_json())
    return 3 if st.blocked else 0


if __name__ == "__main__":
    sys.exit(main())
