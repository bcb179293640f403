#!/usr/bin/env python
"""Unblock-day triage of a staged reference tree.  Reads the tree; never copies from it.

    tools/unblock.sh [TREE]        (default: the first staging root that exists)

What it does, in order, writing only under oracle/_ref/ (git-ignored):
  1. inventory   — commit SHA (if the tree carries .git), source / header / build-file counts,
                   the ten largest data directories (candidate model tables);
  2. headers     — every `#include <...>` the sources use that this container's g++ cannot
                   find (the non-vendored dependencies SURVEY.md says must be staged too);
  3. link deps   — find_package / pkg_check_modules / target_link_libraries / -l names in the
                   build files, each checked against `ldconfig -p` and the tree itself;
  4. compile     — `g++ -std=c++17 -O2 -fPIC -c` of every translation unit, from the sources
                   where they lie, objects into oracle/_ref/obj/ (no cmake run: SURVEY.md §8c);
  5. report      — oracle/_ref/UNBLOCK_REPORT.json + a STAGED.json template for the operator.
Exit status: 0 = everything compiled; 3 = nothing staged (no sources); 4 = staged, but headers
are missing or some units failed — the report lists exactly which.
It knows nothing about base-cpp's layout: every step is generic C/C++ tooling.
"""
from __future__ import annotations

import json
import os
import re
import shutil
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from base_b200 import staging  # noqa: E402

OUT = Path(os.environ.get("B9_UNBLOCK_OUT", ROOT / "oracle" / "_ref"))   # the override is for tests
SRC_EXT = {".c", ".cc", ".cpp", ".cxx"}
HDR_EXT = {".h", ".hh", ".hpp", ".hxx"}
STD_CXX = "-std=c++17"


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, **kw)


def inventory(tree: Path) -> dict:
    srcs, hdrs, builds, data = [], [], [], Counter()
    for dp, dn, fn in os.walk(tree):
        dn[:] = [d for d in dn if d != ".git"]
        for name in fn:
            p = Path(dp) / name
            ext = p.suffix.lower()
            if ext in SRC_EXT:
                srcs.append(p)
            elif ext in HDR_EXT:
                hdrs.append(p)
            elif name in staging._BUILD_FILES:
                builds.append(p)
            elif ext not in staging._DOC_EXT:
                data[str(Path(dp).relative_to(tree))] += 1
    sha = "unknown (no .git in the tree: record the commit by hand in STAGED.json)"
    if (tree / ".git").exists():
        r = run(["git", "-C", str(tree), "rev-parse", "HEAD"])
        if r.returncode == 0:
            sha = r.stdout.strip()
    return {"tree": str(tree), "commit": sha, "sources": srcs, "headers": hdrs, "build_files": builds,
            "data_dirs": data.most_common(10)}


def missing_headers(inv: dict) -> dict[str, list[str]]:
    """system-style includes the container cannot satisfy -> the files that ask for them."""
    users: dict[str, list[str]] = {}
    pat = re.compile(r'^\s*#\s*include\s*<([^>]+)>', re.M)
    own = {h.name for h in inv["headers"]} | {str(h.relative_to(inv["tree"])) for h in inv["headers"]}
    for f in inv["sources"] + inv["headers"]:
        try:
            text = f.read_text(errors="ignore")
        except OSError:
            continue
        for inc in set(pat.findall(text)):
            if inc in own or Path(inc).name in own:
                continue
            users.setdefault(inc, []).append(str(f.relative_to(inv["tree"])))
    missing = {}
    for inc in sorted(users):
        r = run(["g++", STD_CXX, "-fsyntax-only", "-x", "c++", "-"], input=f"#include <{inc}>\n")
        if r.returncode != 0:
            missing[inc] = sorted(users[inc])[:5]
    return missing


def link_deps(inv: dict) -> dict[str, str]:
    names: set[str] = set()
    for b in inv["build_files"]:
        text = b.read_text(errors="ignore")
        names |= set(re.findall(r'find_package\s*\(\s*([A-Za-z0-9_+\-]+)', text))
        names |= set(re.findall(r'pkg_check_modules\s*\(\s*\w+\s+(?:REQUIRED\s+)?([A-Za-z0-9_+\-.]+)', text))
        names |= set(re.findall(r'(?<![\w-])-l([A-Za-z0-9_+\-]+)', text))
    ld = run(["ldconfig", "-p"]).stdout.lower()
    tree_names = {p.name.lower() for p in Path(inv["tree"]).iterdir()}
    out = {}
    for n in sorted(names):
        low = n.lower()
        where = ("system library" if f"lib{low}" in ld or low in ld else
                 "vendored in the tree?" if any(low in t for t in tree_names) else "NOT FOUND in this container")
        out[n] = where
    return out


def compile_all(inv: dict) -> dict:
    obj = OUT / "obj"
    if obj.exists():
        shutil.rmtree(obj)
    obj.mkdir(parents=True)
    inc_dirs = sorted({str(h.parent) for h in inv["headers"]} | {inv["tree"]})
    flags = [STD_CXX, "-O2", "-fPIC", "-w"] + [f"-I{d}" for d in inc_dirs]
    ok, failed = [], {}
    for i, s in enumerate(inv["sources"]):
        cc = "gcc" if s.suffix.lower() == ".c" else "g++"
        fl = [f for f in flags if not (cc == "gcc" and f == STD_CXX)]
        r = run([cc, *fl, "-c", str(s), "-o", str(obj / f"{i:04d}_{s.stem}.o")])
        rel = str(s.relative_to(inv["tree"]))
        if r.returncode == 0:
            ok.append(rel)
        else:
            first = next((l for l in r.stderr.splitlines() if "error" in l), r.stderr.strip().splitlines()[-1:] or [""])
            failed[rel] = first if isinstance(first, str) else " ".join(first)
    return {"compiled": ok, "failed": failed, "include_dirs": len(inc_dirs)}


def main(argv) -> int:
    if len(argv) > 1:
        tree = Path(argv[1]).resolve()
    else:
        tree = next((Path(r) for r in staging.CANDIDATE_ROOTS if Path(r).is_dir()), None)
    if tree is None or not tree.is_dir():
        print("unblock: no staging root exists (/root/reference, baseline/_ref) and no TREE given", file=sys.stderr)
        return 3
    OUT.mkdir(parents=True, exist_ok=True)
    inv = inventory(tree)
    print(f"tree        {inv['tree']}\ncommit      {inv['commit']}")
    print(f"sources     {len(inv['sources'])} translation units, {len(inv['headers'])} headers")
    print("build files " + (", ".join(str(b.relative_to(tree)) for b in inv["build_files"][:8]) or "NONE"))
    print("data dirs   " + ("; ".join(f"{d} ({n})" for d, n in inv["data_dirs"]) or "NONE (no model tables?)"))
    st = staging.probe([tree])
    print(f"staging     {'BLOCKED' if st.blocked else 'staged'}: {st.reason}")
    if not inv["sources"]:
        print("\nNOT STAGED: the tree has no C/C++ translation units.  Needed (BASELINE.md, 'What unblocks this'):\n"
              "  1. base-cpp source at a pinned commit\n  2. its model tables\n  3. its non-vendored build "
              "dependencies as source\nat /root/reference or /root/repo/baseline/_ref.", file=sys.stderr)
        return 3
    miss = missing_headers(inv)
    deps = link_deps(inv)
    comp = compile_all(inv)
    print(f"\nheaders the container lacks ({len(miss)}):")
    for inc, users in miss.items():
        print(f"  <{inc}>   used by {', '.join(users)}")
    print(f"\nlink dependencies named in build files ({len(deps)}):")
    for n, where in deps.items():
        print(f"  {n:24s} {where}")
    print(f"\ncompile: {len(comp['compiled'])} ok, {len(comp['failed'])} failed (objects in {OUT / 'obj'})")
    for rel, err in list(comp["failed"].items())[:20]:
        print(f"  FAILED {rel}: {err}")
    report = {"tree": inv["tree"], "commit": inv["commit"], "n_sources": len(inv["sources"]),
              "n_headers": len(inv["headers"]), "build_files": [str(b.relative_to(tree)) for b in inv["build_files"]],
              "data_dirs": inv["data_dirs"], "missing_headers": miss, "link_deps": deps, **comp}
    (OUT / "UNBLOCK_REPORT.json").write_text(json.dumps(report, indent=1) + "\n")
    (OUT / "STAGED.json.template").write_text(json.dumps(
        {"commit": inv["commit"], "source_root": ".", "table_root": inv["data_dirs"][0][0] if inv["data_dirs"] else "?"},
        indent=1) + "\n")
    print(f"\nreport: {OUT / 'UNBLOCK_REPORT.json'}")
    print("next: redo SURVEY.md sections 1-8 from the source; link the likelihood's objects into oracle/_ref/; "
          "write the dump harness with oracle/b9_dump.h; pin oracle/ against tests/golden/ (tests/golden_io.py).")
    return 0 if not miss and not comp["failed"] else 4


if __name__ == "__main__":
    sys.exit(main(sys.argv))
