"""The golden-vector plumbing, end to end on synthetic data (no reference involved)."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from tests import golden_io as gio

ROOT = Path(__file__).resolve().parent.parent
FIXTURE = ROOT / "tests" / "golden" / "synthetic.b9dump"


def bits(a):
    return np.asarray(a, dtype=np.float64).view(np.int64)


def synthetic():
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    import make_synthetic
    return make_synthetic.records()


def test_committed_fixture_is_what_its_script_writes_and_round_trips_bit_for_bit(tmp_path):
    want = synthetic()
    got = gio.load(FIXTURE)
    assert got.meta["source"].startswith("tests/golden/make_synthetic.py")
    assert got.stages() == ["stageA", "stageB", "total", "edge", "empty"]
    assert len(got.records) == len(want)
    for g, w in zip(got.records, want):
        assert (g.stage, g.star) == (w.stage, w.star) and g.values.size == w.values.size
        nan = np.isnan(w.values)
        assert (np.isnan(g.values) == nan).all()
        assert (bits(g.values[~nan]) == bits(w.values[~nan])).all()          # -0.0, denormal, DBL_MAX, +-inf
    gio.dump(tmp_path / "again.b9dump", got.records, got.meta)
    assert (tmp_path / "again.b9dump").read_text() == FIXTURE.read_text()    # the text itself is canonical


def test_c_writer_and_python_loader_agree(tmp_path):
    src = tmp_path / "w.c"
    src.write_text(r'''
#include "b9_dump.h"
#include <float.h>
int main(int argc, char **argv) {
    b9dump_t d;
    double a[7] = {0.0, -0.0, 4.9406564584124654e-324, -DBL_MAX, INFINITY, -INFINITY, NAN};
    double b[5];
    for (int i = 0; i < 5; ++i) b[i] = ldexp(1.0 + (i + 1) / 7.0, 3 * i - 5) * (i % 2 ? -1.0 : 1.0);
    if (b9dump_open(&d, argv[1])) return 1;
    if (b9dump_meta(&d, "commit", "0000000 (synthetic)")) return 2;
    if (b9dump_record(&d, "edge", 0, a, 7) || b9dump_record(&d, "stageA", 3, b, 5)) return 3;
    if (b9dump_record(&d, "empty", -1, 0, 0)) return 4;
    if (b9dump_record(&d, "bad name", 0, a, 1) != -1) return 5;      /* rejected, nothing written */
    return b9dump_close(&d) ? 6 : 0;
}''')
    exe = tmp_path / "w"
    subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", f"-I{ROOT / 'oracle'}", "-o", str(exe), str(src),
                    "-lm"], check=True)
    out = tmp_path / "c.b9dump"
    assert subprocess.run([str(exe), str(out)]).returncode == 0
    d = gio.load(out)
    assert d.meta == {"commit": "0000000 (synthetic)"} and [r.stage for r in d.records] == ["edge", "stageA", "empty"]
    edge = d.records[0].values
    assert (bits(edge[:6]) == bits([0.0, -0.0, 5e-324, -1.7976931348623157e308, np.inf, -np.inf])).all()
    assert np.isnan(edge[6])
    k = np.arange(5.0)
    assert (bits(d.records[1].values) == bits(np.ldexp(1.0 + (k + 1) / 7.0, (3 * k - 5).astype(int)) * (-1.0) ** k)).all()
    assert d.records[2].star == -1 and d.records[2].values.size == 0


def test_comparator_reports_the_first_divergence_in_pipeline_order():
    want = gio.load(FIXTURE)
    recs = [gio.Record(r.stage, r.star, r.values.copy()) for r in want.records]
    assert gio.compare(gio.Dump({}, recs), want, tol=0.0).ok                 # identical -> ok even at tol 0
    # perturb a LATE stage by a lot and an EARLY stage by a little: the early one must be reported
    recs[-3].values[0] += 1.0                                                # "total"
    recs[2].values[4] = np.nextafter(recs[2].values[4], np.inf)              # "stageA" star 2, 1 ulp
    rep = gio.compare(gio.Dump({}, recs), want, tol=1e-10)
    assert not rep.ok and rep.first.stage == "total"                         # 1 ulp is inside 1e-10 ...
    assert rep.per_stage["stageA"]["max_ulps"] == 1 and rep.per_stage["stageA"]["failures"] == 0
    rep = gio.compare(gio.Dump({}, recs), want, tol=1e-10, bit_exact=["stageA"])
    assert rep.first.stage == "stageA" and rep.first.star == 2 and rep.first.index == 4 and rep.first.ulps == 1
    assert "FIRST DIVERGENCE: stage 'stageA' star 2 element 4" in rep.summary()
    # per-stage tolerances
    rep = gio.compare(gio.Dump({}, recs), want, tol={"*": 1e-10, "total": 10.0})
    assert rep.ok


def test_comparator_structure_errors_and_specials():
    want = gio.load(FIXTURE)
    recs = [gio.Record(r.stage, r.star, r.values.copy()) for r in want.records]
    dropped = recs.pop(1)
    recs.append(gio.Record("stageZ", 0, np.ones(2)))
    recs[0] = gio.Record(recs[0].stage, recs[0].star, recs[0].values[:-1])
    rep = gio.compare(gio.Dump({}, recs), want)
    assert not rep.ok and rep.missing == [(dropped.stage, dropped.star)] and rep.extra == [("stageZ", 0)]
    assert rep.shape_mismatch == [("stageA", 0, 4, 5)]
    # inf where a number is wanted, and a number where nan is wanted, are infinitely wrong
    a = gio.Dump({}, [gio.Record("s", 0, np.array([np.inf, 1.0, np.nan, -np.inf]))])
    b = gio.Dump({}, [gio.Record("s", 0, np.array([1.0, np.nan, np.nan, -np.inf]))])
    rep = gio.compare(a, b, tol=1e300)
    assert rep.per_stage["s"]["failures"] == 2 and rep.first.index == 0


@pytest.mark.parametrize("mutilate,msg", [
    (lambda t: t.replace("b9dump 1", "b9dump 2"), "version-1"),
    (lambda t: t[: t.rindex("end")], "no 'end' line"),
    (lambda t: t.replace("end 11", "end 12"), "truncated"),
    (lambda t: t.replace("rec stageB 0 3", "rec stageB 0 4"), "short"),
    (lambda t: t.replace("rec stageB 0 3", "rec stageB 0 2"), "record is long"),
    (lambda t: t.replace("0x1.2492492492492p-5", "0x1.2492492492492q-5", 1), "not a hex float"),
])
def test_loader_rejects_damaged_files(tmp_path, mutilate, msg):
    text = FIXTURE.read_text()
    bad = mutilate(text)
    assert bad != text
    p = tmp_path / "bad.b9dump"
    p.write_text(bad)
    with pytest.raises(ValueError, match=msg):
        gio.load(p)
