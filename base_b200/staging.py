"""The gate BASELINE.json's north_star puts in front of all other work.

north_star: "Before any other work, the engineer must confirm that the base-cpp
(BASE-9) source and its model tables are staged offline; if they are not, report
BLOCKED in BASELINE.md and do not reconstruct a reference from memory."

`probe()` answers that question mechanically, so that tests, `bench.py` and
`__graft_entry__` all report the same status and flip together the day the
source is staged.  It looks only at the two places SURVEY.md names for staging:
the reference mount and the git-ignored `baseline/_ref/` slot.

A root counts as staged when it has all three of (SURVEY.md "What must be staged"):
  source  — at least MIN_SOURCES C/C++ translation units or headers;
  build   — a build file (CMakeLists.txt / Makefile / configure.ac / meson.build);
  tables  — a model-table DIRECTORY: one directory (at any depth) that directly
            holds at least MIN_TABLE_FILES non-source, non-document data files and
            whose path mentions one of the model families north_star names.
            (Round 1 accepted any single path containing a family name, so a note
            called girardi_notes.txt unblocked it; documents no longer count.)
or when the operator says so explicitly with `<root>/STAGED.json`:
  {"commit": "<sha>", "source_root": "<path>", "table_root": "<path>"}
(paths absolute or relative to the root; both must exist and be non-empty).  The
override is for trees laid out in a way the heuristics above cannot recognise,
e.g. tables staged as a sibling checkout under another name.

On the GPU box neither root exists (gpurun ships /root/repo only, and
`baseline/_ref/` may be too large to ship).  Re-probing there would always say
BLOCKED, whatever the build box saw.  So `build()` records this box's verdict in
`base_b200/STAGING_VERDICT.json` (tracked, deterministic), and a probe that finds
NO root at all returns that carried verdict, labelled as carried.

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
VERDICT_FILE = Path(__file__).resolve().parent / "STAGING_VERDICT.json"
OVERRIDE_NAME = "STAGED.json"

MIN_SOURCES = 5
MIN_TABLE_FILES = 3
_SOURCE_EXT = {".c", ".cc", ".cpp", ".cxx", ".h", ".hh", ".hpp", ".hxx"}
_DOC_EXT = {".md", ".txt", ".rst", ".pdf", ".html", ".htm", ".tex", ".doc", ".docx", ".json", ".yaml", ".yml"}
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
    table_dirs: tuple[str, ...]           # "<dir> (<n> data files)"
    override: str | None = None           # "ok: ..." or "invalid: ..." when STAGED.json is present

    @property
    def staged(self) -> bool:
        if self.override is not None:
            return self.override.startswith("ok")
        return self.n_source >= MIN_SOURCES and bool(self.build_files) and bool(self.table_dirs)


@dataclasses.dataclass(frozen=True)
class StagingStatus:
    blocked: bool
    reason: str
    roots: tuple[RootReport, ...]
    carried: bool = False                 # True when no root exists here and the build box's verdict is used

    def to_json(self) -> str:
        return json.dumps(dataclasses.asdict(self), indent=1)


def _nonempty_dir(p: Path) -> bool:
    try:
        return p.is_dir() and any(p.iterdir())
    except OSError:
        return False


def _check_override(root: Path) -> str | None:
    f = root / OVERRIDE_NAME
    if not f.is_file():
        return None
    try:
        spec = json.loads(f.read_text())
    except (OSError, ValueError) as e:
        return f"invalid: {OVERRIDE_NAME} unreadable ({e})"
    missing = [k for k in ("commit", "source_root", "table_root") if not str(spec.get(k, "")).strip()]
    if missing:
        return f"invalid: {OVERRIDE_NAME} lacks {', '.join(missing)}"
    for k in ("source_root", "table_root"):
        p = Path(spec[k])
        p = p if p.is_absolute() else root / p
        if not _nonempty_dir(p):
            return f"invalid: {k} {p} is not a non-empty directory"
    return f"ok: commit {spec['commit']}, source_root {spec['source_root']}, table_root {spec['table_root']}"


def _scan(root: Path) -> RootReport:
    if not root.is_dir():
        return RootReport(str(root), False, 0, 0, (), ())
    n_files = n_source = 0
    builds: list[str] = []
    tables: list[str] = []
    for dirpath, dirnames, filenames in os.walk(root):
        dirnames[:] = [d for d in dirnames if d != ".git"]
        data_here = 0
        for name in filenames:
            n_files += 1
            if n_files > _MAX_FILES:
                break
            ext = os.path.splitext(name)[1].lower()
            if ext in _SOURCE_EXT:
                n_source += 1
            elif name in _BUILD_FILES:
                if len(builds) < 8:
                    builds.append(os.path.relpath(os.path.join(dirpath, name), root))
            elif ext not in _DOC_EXT and name != OVERRIDE_NAME:
                data_here += 1
        rel = os.path.relpath(dirpath, root).lower()
        if data_here >= MIN_TABLE_FILES and len(tables) < 8 and any(f in rel for f in _TABLE_FAMILIES):
            tables.append(f"{os.path.relpath(dirpath, root)} ({data_here} data files)")
        if n_files > _MAX_FILES:
            break
    return RootReport(str(root), True, n_files, n_source, tuple(builds), tuple(tables), _check_override(root))


def _describe(r: RootReport) -> str:
    if not r.exists:
        return f"{r.root}: absent"
    if r.override is not None and not r.staged:
        return f"{r.root}: {r.override}"
    missing = [w for w, ok in ((f"source (>= {MIN_SOURCES} C/C++ files)", r.n_source >= MIN_SOURCES),
                               ("build files", bool(r.build_files)),
                               (f"model tables (a family-named directory with >= {MIN_TABLE_FILES} data files)",
                                bool(r.table_dirs))) if not ok]
    return f"{r.root}: {r.n_files} file(s), missing {', '.join(missing)}"


def probe(roots=CANDIDATE_ROOTS, use_carried: bool | None = None) -> StagingStatus:
    """use_carried: consult STAGING_VERDICT.json when no root exists (default: only for the
    default roots, so tests that pass their own roots are self-contained)."""
    reports = tuple(_scan(Path(r)) for r in roots)
    if use_carried is None:
        use_carried = tuple(map(str, roots)) == tuple(map(str, CANDIDATE_ROOTS))
    if use_carried and not any(r.exists for r in reports):
        carried = load_verdict()
        if carried is not None:
            return StagingStatus(carried["blocked"],
                                 carried["reason"] + " [verdict carried from the build box: neither staging "
                                 "root exists on this machine]", reports, True)
    if any(r.staged for r in reports):
        first = next(r for r in reports if r.staged)
        how = first.override if first.override else "source, build files and model tables found"
        return StagingStatus(
            False,
            f"a reference tree is staged at {first.root} ({how}): run tools/unblock.sh, redo SURVEY.md "
            "from source (checklists in its sections 1-8), then build the oracle before any kernel",
            reports)
    return StagingStatus(
        True,
        "BLOCKED: base-cpp (BASE-9) source and model tables are not staged offline ("
        + "; ".join(_describe(r) for r in reports)
        + "); north_star forbids reconstructing the reference from memory",
        reports)


def write_verdict(status: StagingStatus | None = None) -> Path:
    """Record this box's verdict for machines that cannot see the staging roots.  Deterministic
    (no timestamp, no per-box paths beyond the two roots) so the tracked file only changes
    when the verdict does."""
    st = status if status is not None else probe(use_carried=False)
    VERDICT_FILE.write_text(json.dumps({"blocked": st.blocked, "reason": st.reason}, indent=1) + "\n")
    return VERDICT_FILE


def load_verdict() -> dict | None:
    try:
        v = json.loads(VERDICT_FILE.read_text())
        return v if isinstance(v.get("blocked"), bool) and isinstance(v.get("reason"), str) else None
    except (OSError, ValueError):
        return None


BLOCKED_ONE_LINE = ("reference is a relocation stub (/root/reference/README.md:1-4); base-cpp "
                    "source and model tables are not staged offline and there is no network")


def main(argv=None) -> int:
    st = probe()
    print(st.to_json())
    return 3 if st.blocked else 0


if __name__ == "__main__":
    sys.exit(main())
