"""The north_star gate: BLOCKED unless source + build files + model tables are staged."""
import json
import subprocess
import sys
from pathlib import Path

from base_b200 import staging

ROOT = Path(__file__).resolve().parent.parent


def test_stub_tree_is_blocked(tmp_path):
    # exactly what /root/reference holds: one README pointing elsewhere
    (tmp_path / "README.md").write_text("The BASE-9 code is available at a new location\n")
    st = staging.probe([tmp_path])
    assert st.blocked and "not staged" in st.reason and "missing source" in st.reason


def test_absent_roots_are_blocked(tmp_path):
    # the GPU box: neither /root/reference nor baseline/_ref exists
    st = staging.probe([tmp_path / "nope", tmp_path / "nada"])
    assert st.blocked and st.reason.count(": absent") == 2


def test_source_without_tables_is_still_blocked(tmp_path):
    (tmp_path / "src").mkdir()
    (tmp_path / "src" / "a.cpp").write_text("int main(){}\n")
    (tmp_path / "CMakeLists.txt").write_text("project(x)\n")
    st = staging.probe([tmp_path])
    assert st.blocked and "model tables" in st.reason


def test_full_tree_unblocks_and_demands_a_resurvey(tmp_path):
    (tmp_path / "src").mkdir()
    (tmp_path / "src" / "a.cpp").write_text("int main(){}\n")
    (tmp_path / "CMakeLists.txt").write_text("project(x)\n")
    (tmp_path / "models" / "parsec").mkdir(parents=True)
    (tmp_path / "models" / "parsec" / "t.dat").write_text("0 0\n")
    st = staging.probe([tmp_path / "absent", tmp_path])
    assert not st.blocked and "redo SURVEY.md" in st.reason


def test_live_status_and_docs_agree():
    st = staging.probe()
    if st.blocked:
        for doc in ("DESIGN.md", "BASELINE.md", "INTEGRATION.md"):
            assert "BLOCKED" in (ROOT / doc).read_text(), doc
    r = subprocess.run([sys.executable, "-m", "base_b200.staging"], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == (3 if st.blocked else 0)
    assert json.loads(r.stdout)["blocked"] == st.blocked
