"""GPU tests: every groundwork kernel, through the C-ABI, against the CPU checker.

Tolerances, and why:
  dfma chain, ordered sum  — IEEE add/fma only            -> bit-exact
  exp / log elementwise    — CUDA libm documents <= 1 ULP (exp) and <= 1 ULP (log) in FP64,
                             glibc < 1 ULP                 -> distance <= 2 ULP
  row log-sum-exp          — same order as the kernel      -> <= 2e-14 of max(1,|lse|)
                             serial CPU order              -> <= 1e-13 of max(1,|lse|)
                             (mixed abs/rel: a row value near 0 makes pure relative error
                             ill-conditioned, see tests/_ref.py:mixed_err)
north_star's bar for the real path is 1e-10 relative; these show the margin available.
"""
import numpy as np
import pytest

from tests._ref import mixed_err, ulp_distance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gw(built):
    from base_b200 import groundwork
    assert groundwork.device_count() > 0, "no CUDA device: run under gpurun"
    return groundwork


def bits(a):
    return np.asarray(a, dtype=np.float64).view(np.int64)


def test_device_is_a_b200(gw):
    info = gw.device_info(0)
    assert info["sm_count"] == 148 and info["l2_bytes"] > 100 << 20


@pytest.mark.parametrize("iters,ctas", [(1, 1), (7, 1), (4096, 2), (100_000, 8)])
def test_dfma_chain_bit_exact(gw, ref, iters, ctas):
    a, b = 1.0 - 2.0 ** -12, 2.0 ** -12
    r = gw.dfma_peak(0, ctas_per_sm=ctas, iters=iters, a=a, b=b, warmup=0, reps=1, want_out=True)
    assert r["n_threads"] == 148 * ctas * 256
    want = ref.dfma_lanes(a, b, iters)
    assert (bits(r["out"].reshape(-1, 32)) == bits(want)).all()


def test_dfma_peak_is_plausible(gw):
    r = gw.dfma_peak(0, ctas_per_sm=8, iters=1 << 16, warmup=3, reps=5)
    # 148 SMs x 64 DFMA/clk x 2 x 1.965 GHz = 37.2 TF is the arithmetic ceiling
    assert 5.0 < r["tflops"] < 45.0, r


@pytest.mark.parametrize("which", ["exp", "log", "exp10", "log10"])
def test_transcendental_chain(gw, ref, which):
    r = gw.transcendental_rate(which, 0, ctas_per_sm=1, iters=200, warmup=0, reps=1, want_out=True)
    want = ref.trans_lanes(which, 200)
    got = r["out"].reshape(-1, 32)
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-14   # contraction: libm bits wash out
    assert r["gevals_per_s"] > 1.0


@pytest.mark.parametrize("which", ["exp", "log"])
def test_libm_distance_within_2_ulp(gw, ref, which):
    rng = np.random.default_rng(42)
    if which == "exp":   # the arguments a max-shifted log-sum-exp produces, plus the full finite range
        x = np.concatenate([-rng.exponential(20.0, 500_000), rng.uniform(-745.0, 709.0, 500_000),
                            [0.0, -0.0, -745.2, 709.7, -np.inf]])
    else:
        x = np.concatenate([rng.uniform(1.0, 1024.0, 500_000), 10.0 ** rng.uniform(-300, 300, 500_000),
                            [1.0, 5e-324, 1.7976931348623157e308]])
    got, want = gw.device_map(which, x), ref.map(which, x)
    d = ulp_distance(got, want)
    assert d.max() <= 2, (d.max(), x[d.argmax()])
    print(f"\n{which}: bit-identical {np.mean(d == 0):.4%}, max {d.max()} ulp")


def test_step_latency_is_measured_and_ordered(gw):
    r = gw.step_latency(0, warmup=20, reps=300)
    assert 1.0 < r["us_launch_sync"] < 500.0
    assert r["us_launch_d2h_sync"] >= 0.8 * r["us_launch_sync"]
    assert 1.0 < r["us_graph_d2h_sync"] < 500.0


def test_map_empty(gw):
    assert gw.device_map("exp", np.empty(0)).size == 0


@pytest.mark.parametrize("rows,cols", [(1, 1), (5, 31), (9, 32), (8, 33), (257, 1000), (10_000, 1024)])
def test_lse_rows_against_both_orders(gw, ref, rows, cols):
    rng = np.random.default_rng(rows * 7919 + cols)
    x = rng.normal(-40.0, 12.0, size=(rows, cols))
    r = gw.lse_rows(x)
    same_order, serial = ref.lse_rows(x, True), ref.lse_rows(x, False)
    assert mixed_err(r["row_lse"], same_order) < 2e-14
    assert mixed_err(r["row_lse"], serial) < 1e-13
    # the total is adds only, in a fixed order: bit-exact given the kernel's own row values
    assert bits(r["total"]) == bits(ref.ordered_sum(r["row_lse"]))
    assert mixed_err(r["total"], ref.serial_sum(serial)) < 1e-13


def test_lse_edge_cases(gw, ref):
    x = np.full((4, 70), -np.inf)
    x[1, 69] = -700.0
    x[2, :] = 700.0
    x[3, :] = np.linspace(-1e4, 0.0, 70)          # terms underflowing to exactly 0
    r = gw.lse_rows(x)
    assert r["row_lse"][0] == -np.inf and r["row_lse"][1] == -700.0
    assert r["row_lse"][2] == pytest.approx(700.0 + np.log(70.0), rel=1e-15)
    assert r["row_lse"][3] == pytest.approx(ref.lse_rows(x, False)[3], rel=1e-14)
    assert r["total"] == -np.inf
    e = gw.lse_rows(np.empty((6, 0)))              # empty grid: every row -inf
    assert (e["row_lse"] == -np.inf).all() and e["total"] == -np.inf
    z = gw.lse_rows(np.empty((0, 9)))              # no stars: empty sum
    assert z["row_lse"].size == 0 and z["total"] == 0.0


def test_lse_is_run_to_run_deterministic(gw):
    x = np.random.default_rng(5).normal(-40.0, 12.0, size=(3000, 777))
    a, b = gw.lse_rows(x), gw.lse_rows(x)
    assert (bits(a["row_lse"]) == bits(b["row_lse"])).all() and bits(a["total"]) == bits(b["total"])


def test_smoke_entry_point(gw):
    import __graft_entry__ as g
    g.smoke()
