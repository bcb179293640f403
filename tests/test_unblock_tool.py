"""tools/unblock.sh on synthetic trees: it must fail loudly on the stub, list what a staged tree
lacks, and compile what can be compiled — without knowing anything about the reference."""
import json
import os
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def run(tree, out):
    env = dict(os.environ, B9_UNBLOCK_OUT=str(out))
    return subprocess.run([str(ROOT / "tools" / "unblock.sh"), str(tree)], capture_output=True, text=True, env=env,
                          timeout=300)


def make_tree(root, with_missing_dep):
    (root / "src").mkdir(parents=True)
    (root / "include").mkdir()
    (root / "include" / "own.hpp").write_text("#pragma once\ninline int own() { return 1; }\n")
    for extra in ("a.hpp", "b.hpp", "c.hpp"):
        (root / "include" / extra).write_text("#pragma once\n")
    (root / "src" / "good.cpp").write_text('#include <vector>\n#include "own.hpp"\nint good() { return own(); }\n')
    (root / "src" / "plain.c").write_text("#include <math.h>\ndouble plain(double x) { return sqrt(x); }\n")
    cm = "project(x)\nfind_package(Threads)\n"
    if with_missing_dep:
        (root / "src" / "needs.cpp").write_text("#include <not_a_real_lib/thing.hpp>\nint needs() { return 0; }\n")
        cm += "find_package(NotARealLib REQUIRED)\ntarget_link_libraries(x -lnotareallib)\n"
    (root / "CMakeLists.txt").write_text(cm)
    (root / "models" / "dsed").mkdir(parents=True)
    for i in range(4):
        (root / "models" / "dsed" / f"iso{i}.dat").write_text("0 0 0\n")


def test_stub_fails_with_the_operator_request(tmp_path):
    (tmp_path / "stub").mkdir()
    (tmp_path / "stub" / "README.md").write_text("moved\n")
    r = run(tmp_path / "stub", tmp_path / "out")
    assert r.returncode == 3
    assert "NOT STAGED" in r.stderr and "model tables" in r.stderr and "baseline/_ref" in r.stderr


def test_staged_tree_with_a_missing_dependency_is_reported_exactly(tmp_path):
    make_tree(tmp_path / "t", with_missing_dep=True)
    r = run(tmp_path / "t", tmp_path / "out")
    assert r.returncode == 4, r.stdout + r.stderr
    rep = json.loads((tmp_path / "out" / "UNBLOCK_REPORT.json").read_text())
    assert list(rep["missing_headers"]) == ["not_a_real_lib/thing.hpp"]
    assert rep["missing_headers"]["not_a_real_lib/thing.hpp"] == ["src/needs.cpp"]
    assert rep["link_deps"]["NotARealLib"].startswith("NOT FOUND") and "notareallib" in rep["link_deps"]
    assert sorted(rep["compiled"]) == ["src/good.cpp", "src/plain.c"] and list(rep["failed"]) == ["src/needs.cpp"]
    assert rep["data_dirs"][0] == ["models/dsed", 4]
    assert "staging     staged" in r.stdout and "<not_a_real_lib/thing.hpp>   used by src/needs.cpp" in r.stdout
    assert len(list((tmp_path / "out" / "obj").glob("*.o"))) == 2


def test_clean_tree_compiles_and_exits_zero(tmp_path):
    make_tree(tmp_path / "t", with_missing_dep=False)
    r = run(tmp_path / "t", tmp_path / "out")
    assert r.returncode == 0, r.stdout + r.stderr
    tpl = json.loads((tmp_path / "out" / "STAGED.json.template").read_text())
    assert tpl["table_root"] == "models/dsed" and tpl["source_root"] == "."
