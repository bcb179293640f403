"""The north_star gate: BLOCKED unless source + build files + model tables are staged."""
import json
import subprocess
import sys
from pathlib import Path

from base_b200 import staging

ROOT = Path(__file__).resolve().parent.parent


def make_tree(root, sources=6, build=True, table_dir="models/parsec", table_files=4):
    (root / "src").mkdir(parents=True)
    for i in range(sources):
        (root / "src" / f"f{i}.cpp").write_text("int f();\n")
    if build:
        (root / "CMakeLists.txt").write_text("project(x)\n")
    if table_dir:
        (root / table_dir).mkdir(parents=True)
        for i in range(table_files):
            (root / table_dir / f"t{i}.dat").write_text("0 0\n")


def test_stub_tree_is_blocked(tmp_path):
    # exactly what /root/reference holds: one README pointing elsewhere
    (tmp_path / "README.md").write_text("The BASE-9 code is available at a new location\n")
    st = staging.probe([tmp_path])
    assert st.blocked and "not staged" in st.reason and "missing source" in st.reason


def test_absent_roots_are_blocked_when_nothing_is_carried(tmp_path):
    st = staging.probe([tmp_path / "nope", tmp_path / "nada"])
    assert st.blocked and not st.carried and st.reason.count(": absent") == 2


def test_source_without_tables_is_still_blocked(tmp_path):
    make_tree(tmp_path, table_dir=None)
    st = staging.probe([tmp_path])
    assert st.blocked and "model tables" in st.reason


def test_a_document_that_merely_names_a_family_does_not_unblock(tmp_path):
    # "girardi_notes.txt" used to count as tables
    make_tree(tmp_path, table_dir=None)
    (tmp_path / "docs").mkdir()
    (tmp_path / "docs" / "girardi_notes.txt").write_text("notes\n")
    (tmp_path / "docs" / "parsec_readme.md").write_text("notes\n")
    (tmp_path / "girardi").mkdir()
    (tmp_path / "girardi" / "one_table.dat").write_text("0\n")           # a directory, but too few files
    assert staging.probe([tmp_path]).blocked


def test_full_tree_unblocks_and_demands_a_resurvey(tmp_path):
    make_tree(tmp_path)
    st = staging.probe([tmp_path / "absent", tmp_path])
    assert not st.blocked and "redo SURVEY.md" in st.reason and "tools/unblock.sh" in st.reason
    assert st.roots[1].table_dirs == ("models/parsec (4 data files)",)


def test_explicit_override_accepts_a_layout_the_heuristics_cannot_see(tmp_path):
    root, tables = tmp_path / "ref", tmp_path / "elsewhere" / "grids"
    (root / "code").mkdir(parents=True)
    (root / "code" / "a.cpp").write_text("int main(){}\n")               # 1 source, no build file
    tables.mkdir(parents=True)
    (tables / "x.bin").write_bytes(b"\0")
    assert staging.probe([root]).blocked
    (root / "STAGED.json").write_text(json.dumps({"commit": "abc123", "source_root": "code", "table_root": str(tables)}))
    st = staging.probe([root])
    assert not st.blocked and "commit abc123" in st.reason
    # ... and rejects one that points nowhere, saying why
    (root / "STAGED.json").write_text(json.dumps({"commit": "abc123", "source_root": "code", "table_root": "missing"}))
    st = staging.probe([root])
    assert st.blocked and "table_root" in st.reason and "not a non-empty directory" in st.reason
    (root / "STAGED.json").write_text(json.dumps({"source_root": "code", "table_root": str(tables)}))
    assert "lacks commit" in staging.probe([root]).reason


def test_verdict_is_carried_to_a_box_without_the_roots(tmp_path, monkeypatch):
    make_tree(tmp_path / "ref")
    here = staging.probe([tmp_path / "ref"])
    monkeypatch.setattr(staging, "VERDICT_FILE", tmp_path / "verdict.json")
    staging.write_verdict(here)
    first = (tmp_path / "verdict.json").read_text()
    staging.write_verdict(here)
    assert (tmp_path / "verdict.json").read_text() == first               # deterministic
    there = staging.probe([tmp_path / "gone1", tmp_path / "gone2"], use_carried=True)
    assert there.carried and not there.blocked and "carried from the build box" in there.reason
    # a root that EXISTS always wins over the carried verdict
    (tmp_path / "stub").mkdir()
    assert staging.probe([tmp_path / "stub"], use_carried=True).blocked


def test_live_status_docs_and_carried_verdict_agree():
    st = staging.probe()
    if st.blocked:
        for doc in ("DESIGN.md", "BASELINE.md", "INTEGRATION.md", "README.md"):
            assert "BLOCKED" in (ROOT / doc).read_text(), doc
    carried = staging.load_verdict()
    assert carried is not None, "base_b200/STAGING_VERDICT.json is missing: run __graft_entry__.build()"
    if any(r.exists for r in st.roots):
        assert carried["blocked"] == st.blocked, "stale STAGING_VERDICT.json: run __graft_entry__.build()"
    r = subprocess.run([sys.executable, "-m", "base_b200.staging"], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == (3 if st.blocked else 0)
    assert json.loads(r.stdout)["blocked"] == st.blocked
