"""GPU tests: every groundwork kernel, through the C-ABI, against the CPU checker.

Tolerances, and why:
  dfma chain, ordered sum, shard partials, virtual-shard totals, the term generator
                           — IEEE add/mul/div/fma only    -> bit-exact
  exp / log / exp10 / log10 elementwise
                           — CUDA libm documents <= 1 ULP for each in FP64, glibc < 1 ULP
                                                           -> distance <= 2 ULP
  row log-sum-exp          — same order as the kernel      -> <= 2e-14 of max(1,|lse|)
                             serial CPU order              -> <= 1e-13 of max(1,|lse|)
                             (mixed abs/rel: a row value near 0 makes pure relative error
                             ill-conditioned, see tests/_ref.py:mixed_err)
  lse_generated vs lse_rows on the materialised terms      -> bit-exact (same device libm)
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


@pytest.mark.parametrize("iters,ctas,ilp", [(1, 1, 8), (7, 1, 1), (4096, 2, 2), (333, 1, 4), (100_000, 8, 8)])
def test_dfma_chain_bit_exact(gw, ref, iters, ctas, ilp):
    a, b = 1.0 - 2.0 ** -12, 2.0 ** -12
    r = gw.dfma_peak(0, ctas_per_sm=ctas, iters=iters, a=a, b=b, warmup=0, reps=1, want_out=True, ilp=ilp)
    assert r["n_threads"] == 148 * ctas * 256
    want = ref.dfma_lanes(a, b, iters, ilp)
    assert (bits(r["out"].reshape(-1, 32)) == bits(want)).all()


@pytest.mark.parametrize("fill", [1, 2, -2, -4])
def test_dfma_chain_with_integer_company_is_still_bit_exact(gw, ref, fill):
    a, b = 1.0 - 2.0 ** -12, 2.0 ** -12
    r = gw.dfma_peak(0, ctas_per_sm=2, iters=777, a=a, b=b, warmup=0, reps=1, want_out=True, int_per_fma=fill)
    assert (bits(r["out"].reshape(-1, 32)) == bits(ref.dfma_lanes(a, b, 777))).all()


def test_dfma_peak_is_within_3_percent_of_the_recorded_denominator(gw):
    r = gw.dfma_peak(0, ctas_per_sm=8, iters=1 << 16, warmup=3, reps=5)
    # 148 SMs x 64 DFMA/clk x 2 x 1.965 GHz = 37.2 TF is the arithmetic ceiling; rounds 1 and 2
    # measured 36.99-37.10 (profiles/).  A box that misses this window is throttled or clock-locked.
    assert 37.0 * 0.97 < r["tflops"] < 37.0 * 1.03, r


@pytest.mark.parametrize("which", ["exp", "log", "exp10", "log10"])
def test_transcendental_chain(gw, ref, which):
    r = gw.transcendental_rate(which, 0, ctas_per_sm=1, iters=200, warmup=0, reps=1, want_out=True)
    want = ref.trans_lanes(which, 200)
    got = r["out"].reshape(-1, 32)
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-14   # contraction: libm bits wash out
    assert r["gevals_per_s"] > 1.0


@pytest.mark.parametrize("which", ["exp_spread", "log_spread"])
def test_spread_rate_sums_match_the_checker(gw, ref, which):
    iters = 300
    r = gw.transcendental_rate(which, 0, ctas_per_sm=1, iters=iters, warmup=0, reps=1, want_out=True)
    want = ref.trans_lanes(which, iters)
    got = r["out"].reshape(-1, 32)
    # sums of 4*iters libm values of magnitude <= 1 (exp) / <= 5.6 (log), each within 1 ulp of the
    # host's; a differing input can also move the rounding of the running sum (ulp(acc) ~ 3e-14)
    assert np.max(np.abs(got - want)) < 4 * iters * 6.0 * 2.3e-16 * 8
    assert (got == got[0]).all() and r["gevals_per_s"] > 1.0
    a = ref.spread_args(which, 5, 2, 4096)               # the distribution the header promises
    if which == "exp_spread":
        assert a.max() <= -2.0 ** -6 and a.min() > -1024.0 and 0.01 < np.mean(a < -745.2) < 0.06
    else:
        assert a.min() >= 2.0 ** -8 and a.max() < 256.0


LIBM_CASES = {
    # the arguments a max-shifted log-sum-exp produces, plus the full finite range
    "exp": lambda rng: np.concatenate([-rng.exponential(20.0, 500_000), rng.uniform(-745.0, 709.0, 500_000),
                                       [0.0, -0.0, -745.2, 709.7, -np.inf]]),
    "log": lambda rng: np.concatenate([rng.uniform(1.0, 1024.0, 500_000), 10.0 ** rng.uniform(-300, 300, 500_000),
                                       [1.0, 5e-324, 1.7976931348623157e308]]),
    # 10^(-0.4 m) for magnitudes m in [-10, 40], plus the full finite range and exact powers
    "exp10": lambda rng: np.concatenate([-0.4 * rng.uniform(-10.0, 40.0, 500_000),
                                         rng.uniform(-323.0, 308.0, 500_000),
                                         np.arange(-20.0, 23.0), [0.0, -0.0, -323.4, 308.2, -np.inf]]),
    "log10": lambda rng: np.concatenate([rng.uniform(1e-3, 1e3, 500_000), 10.0 ** rng.uniform(-300, 300, 500_000),
                                         10.0 ** np.arange(-20.0, 23.0), [1.0, 5e-324, 1.7976931348623157e308]]),
}


@pytest.mark.parametrize("which", ["exp", "log", "exp10", "log10"])
def test_libm_distance_within_2_ulp(gw, ref, which):
    x = LIBM_CASES[which](np.random.default_rng(42))
    got, want = gw.device_map(which, x), ref.map(which, x)
    d = ulp_distance(got, want)
    assert d.max() <= 2, (d.max(), x[d.argmax()])
    print(f"\n{which}: bit-identical {np.mean(d == 0):.4%}, max {d.max()} ulp")


def test_lse_exp_fast_path_is_libm_exp_bit_for_bit(gw):
    # lse.cu evaluates exp through a branch-free copy of CUDA libm's fast path whenever a warp's
    # arguments are all above -708.  It must be the same function there, to the last bit.
    rng = np.random.default_rng(2026)
    x = np.concatenate([
        -rng.uniform(0.0, 707.999, 6_000_000),                      # the whole range, uniformly
        -rng.exponential(3.0, 2_000_000),                           # where the sum's mass is
        -(2.0 ** rng.uniform(-1074, 9.4, 2_000_000)),               # every binade incl. denormals
        -np.arange(0.0, 708.0, 0.5), -np.log(2.0) * np.arange(0, 1021),   # reduction boundaries
        [0.0, -0.0, -5e-324, -707.9999999999999, np.nextafter(-708.0, 0.0)]])
    x = x[x > -708.0]
    fast, libm = gw.device_map("exp_fast_path", x), gw.device_map("exp", x)
    bad = np.flatnonzero(bits(fast) != bits(libm))
    assert bad.size == 0, (bad.size, x[bad[:5]], fast[bad[:5]], libm[bad[:5]])
    print(f"\nexp fast path == exp() on {x.size} arguments")


def test_device_exp10_against_host_pow10(gw, ref):
    # ADVICE r1 / VERDICT weak 7: device exp10() and host pow(10, x) are different functions.
    # Measured, not assumed: both are compared with the correctly rounded value where it is
    # known exactly (10^k, k = 0..22) and with each other elsewhere.
    x = LIBM_CASES["exp10"](np.random.default_rng(43))
    dev, p10, e10 = gw.device_map("exp10", x), ref.map("pow10", x), ref.map("exp10", x)
    d_pow, d_e10 = ulp_distance(dev, p10), ulp_distance(dev, e10)
    assert d_pow.max() <= 2 and d_e10.max() <= 2, (d_pow.max(), d_e10.max())
    k = np.arange(0.0, 23.0)
    exact = np.array([float(10 ** int(i)) for i in k])
    assert (gw.device_map("exp10", k) == exact).all()     # CUDA exp10 is exact on exact powers
    print(f"\nexp10 vs pow(10,x): identical {np.mean(d_pow == 0):.4%}, max {d_pow.max()} ulp; "
          f"vs glibc exp10: identical {np.mean(d_e10 == 0):.4%}, max {d_e10.max()} ulp")


def test_step_latency_is_measured_and_ordered(gw):
    r = gw.step_latency(0, warmup=20, reps=300)
    assert 1.0 < r["us_launch_sync"] < 500.0
    assert r["us_launch_d2h_sync"] >= 0.8 * r["us_launch_sync"]
    assert 1.0 < r["us_graph_d2h_sync"] < 500.0


def test_map_empty(gw):
    assert gw.device_map("exp", np.empty(0)).size == 0


# cols straddle every dispatch edge of lse.cu: 128/129 (4 -> 16 terms per lane), 512/513
# (16 -> 32), 1024/1025 (registers -> streaming), 3000 (several streaming chunks + ragged tail)
@pytest.mark.parametrize("rows,cols", [(1, 1), (5, 31), (9, 32), (8, 33), (17, 128), (17, 129),
                                       (33, 512), (33, 513), (257, 1000), (10_000, 1024),
                                       (40, 1025), (23, 3000)])
def test_lse_rows_against_both_orders(gw, ref, rows, cols):
    rng = np.random.default_rng(rows * 7919 + cols)
    x = rng.normal(-40.0, 12.0, size=(rows, cols))
    r = gw.lse_rows(x)
    same_order, serial = ref.lse_rows(x, True), ref.lse_rows(x, False)
    assert mixed_err(r["row_lse"], same_order) < 2e-14
    assert mixed_err(r["row_lse"], serial) < 1e-13
    # shard partials and total are adds only, in a fixed order: bit-exact given the kernel's
    # own row values, for every shard count (rows < V leaves shards empty)
    for V in (4, 64, 128):
        rv = r if V == 64 else gw.lse_rows(x, n_vshards=V)
        want_p, want_t = ref.rows_total(rv["row_lse"], V)
        assert (bits(rv["partials"]) == bits(want_p)).all() and bits(rv["total"]) == bits(want_t)
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
    assert z["row_lse"].size == 0 and z["total"] == 0.0 and (z["partials"] == 0.0).all()


def test_lse_is_run_to_run_deterministic(gw):
    x = np.random.default_rng(5).normal(-40.0, 12.0, size=(3000, 777))
    a, b = gw.lse_rows(x), gw.lse_rows(x)
    assert (bits(a["row_lse"]) == bits(b["row_lse"])).all() and bits(a["total"]) == bits(b["total"])


@pytest.mark.parametrize("rows,cols", [(1, 1), (3, 2), (9, 33), (64, 128), (65, 200), (100, 513),
                                       (10_000, 1024), (50, 1025), (7, 2500)])
def test_generator_and_generated_lse_are_bit_identical_to_the_materialised_path(gw, ref, rows, cols):
    x = gw.generate_terms(rows, cols)
    assert (bits(x) == bits(ref.generate_terms(rows, cols))).all()      # arithmetic only: bit-exact
    assert x.max() <= 0.0 and x.min() > -1700.0
    mat, gen = gw.lse_rows(x), gw.lse_generated(rows, cols)
    assert (bits(mat["row_lse"]) == bits(gen["row_lse"])).all()
    assert bits(mat["total"]) == bits(gen["total"])
    assert mixed_err(gen["row_lse"], ref.lse_rows(x, True)) < 2e-14
    assert (bits(mat["partials"]) == bits(gen["partials"])).all()
    assert bits(gen["total"]) == bits(ref.rows_total(gen["row_lse"])[1])


def test_generated_lse_empty_shapes(gw):
    e = gw.lse_generated(6, 0)
    assert (e["row_lse"] == -np.inf).all() and e["total"] == -np.inf
    z = gw.lse_generated(0, 9)
    assert z["row_lse"].size == 0 and z["total"] == 0.0
    assert gw.generate_terms(0, 5).size == 0 and gw.generate_terms(5, 0).size == 0


def test_lse_total_survives_repeated_launches(gw, ref):
    # the shard tickets must reset themselves: reps > 1 on one stream, total still the shard sum
    for shape in ((1234, 300), (50, 1500), (3, 40)):
        x = np.random.default_rng(11).normal(-40.0, 12.0, size=shape)
        r = gw.lse_rows(x, warmup=2, reps=5)
        want_p, want_t = ref.rows_total(r["row_lse"])
        assert (bits(r["partials"]) == bits(want_p)).all() and bits(r["total"]) == bits(want_t)


def vshard_values(chains, n, seed):
    rng = np.random.default_rng(seed)
    return rng.normal(size=(chains, n)) * 10.0 ** rng.integers(-6, 6, size=(chains, n))


@pytest.mark.parametrize("V", [4, 8, 16, 32, 64, 128])
@pytest.mark.parametrize("chains,n", [(1, 1), (3, 5), (63, 1003), (64, 130), (65, 4096), (1000, 777)])
def test_vshard_total_bit_exact(gw, ref, V, chains, n):
    values = vshard_values(chains, n, V * 1000 + chains + n)      # n < V leaves some shards empty
    got = gw.vshard_total(values, V)
    want_p, want_t = ref.vshard_total(values, V)
    assert (bits(got["partials"]) == bits(want_p)).all()
    assert (bits(got["total"]) == bits(want_t)).all()


def test_vshard_total_no_stars(gw):
    got = gw.vshard_total(np.empty((5, 0)), 64)
    assert (got["partials"] == 0.0).all() and (got["total"] == 0.0).all()


def test_peer_comm_world_1_through_device_pointers_and_a_cuda_graph(gw, ref):
    import torch
    from base_b200.vshards import PeerComm
    chains, n, V = 300, 2000, 64
    values = vshard_values(chains, n, 99)
    _, want = ref.vshard_total(values, V)
    dv = torch.from_numpy(values).cuda()
    with PeerComm(0, 0, 1, V, max_chains=512) as comm:
        P = comm.shard_partials(dv, n)
        out = comm.allreduce(P)
        assert (bits(out.cpu().numpy()) == bits(want)).all()
        # fewer chains than max_chains, then the same launch captured and replayed: the step
        # counter lives in device memory, so replays keep working
        out2 = torch.empty(chains, dtype=torch.float64, device="cuda")
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            comm.allreduce(P, out2)                       # warm the stream outside capture
            with torch.cuda.graph(g, stream=s):
                comm.allreduce(P, out2)
        for _ in range(3):
            out2.zero_()
            g.replay()
            torch.cuda.synchronize()
            assert (bits(out2.cpu().numpy()) == bits(want)).all()
        st = comm.status()
        assert not st["timed_out"] and st["steps"] == 1 + 1 + 3
        lat = comm.latency(chains, warmup=5, reps=40)
        assert 0.5 < lat["us_stream"] < 200.0 and 0.5 < lat["us_graph"] < 200.0


@pytest.mark.parametrize("n_total,cols,chains,V", [
    (1003, 96, 5, 64),        # ragged shards, padded columns
    (10_000, 1024, 3, 64),    # cfg2-sized, full staged rows
    (37, 50, 4, 64),          # fewer stars than shards: empty shards, ranks with no star at all
    (700, 1500, 2, 16),       # rows longer than the staged limit (stream kernel)
    (129, 7, 1, 128),
])
def test_sharded_lse_has_the_same_bits_at_every_world_size(gw, ref, n_total, cols, chains, V):
    """b9gw_lse_generated_shards: W emulated ranks, one after the other on this GPU, each over its
    own V/W shards.  Row values are those of the one-launch kernel over all chains*n_total rows
    (bit-exact: same device code), the P[v][chain] they produce are the checker's over those row
    values (bit-exact: adds only), and neither depends on W."""
    flat = gw.lse_generated(chains * n_total, cols, 0)["row_lse"].reshape(chains, n_total)
    wantP, want_total = ref.vshard_total(flat, V)
    for W in (1, 2, 4, 8):
        if V % W:
            continue
        per = V // W
        P = np.empty((V, chains))
        for r in range(W):
            out = gw.lse_generated_shards(n_total, cols, chains, V, r * per, per, 0, launches=2)
            lo, hi = ref.shard_lo(n_total, V, r * per), ref.shard_lo(n_total, V, (r + 1) * per)
            assert (bits(out["row_lse"]) == bits(flat[:, lo:hi])).all(), (W, r)
            assert not out["workspace"].any(), "tickets must be zero again after a launch"
            P[r * per:(r + 1) * per] = out["partials"]
            if W == 1:
                assert (bits(out["total"]) == bits(want_total)).all()
        assert (bits(P) == bits(wantP)).all(), W
        assert (bits(gw.vshard_total(flat, V)["total"]) == bits(want_total)).all()


def test_sharded_lse_rejects_bad_jobs(gw):
    for args in [(10, 8, 1, 64, 60, 8), (10, 8, 70_000, 64, 0, 64), (1 << 24, 8, 256, 128, 0, 1),
                 (10, 8, 1, 48, 0, 8)]:
        with pytest.raises(gw.GroundworkError) as e:
            gw.lse_generated_shards(*args, 0)
        assert e.value.code == gw.E_ARG


def test_sharded_lse_no_stars_and_no_chains(gw):
    out = gw.lse_generated_shards(0, 16, 3, 64, 0, 64, 0)
    assert (bits(out["partials"]) == 0).all() and (bits(out["total"]) == 0).all()
    out = gw.lse_generated_shards(100, 16, 0, 64, 0, 64, 0)
    assert out["partials"].size == 0


def test_sharded_step_world_1(gw, ref):
    """b9gw_sharded_step on one rank: LSE share + cross-rank sum, total checked against the
    checker's sum over the one-launch kernel's row values."""
    from base_b200 import vshards
    n_total, cols, chains, V = 2_000, 256, 7, 64
    flat = gw.lse_generated(chains * n_total, cols, 0)["row_lse"].reshape(chains, n_total)
    _, want = ref.vshard_total(flat, V)
    with vshards.PeerComm(0, 0, 1, V, max_chains=16) as comm:
        r = comm.sharded_step(n_total, cols, chains, warmup=2, reps=5)
        assert (bits(r["total"]) == bits(want)).all()
        assert (bits(r["total_fused"]) == bits(want)).all(), "the one-kernel step differs from the two-launch step"
        assert 0 < r["us_lse_alone"] <= r["us_step"] * 1.5 and r["us_fused_step"] > 0
        assert comm.status()["steps"] == 14          # 7 two-launch steps + 7 fused ones on chain 0


@pytest.mark.parametrize("n_total,cols,chains,V", [(1003, 96, 5, 64), (37, 50, 4, 64), (700, 1500, 2, 16),
                                                   (0, 8, 3, 64), (5000, 1024, 33, 128)])
def test_fused_step_world_1_mixes_with_the_two_launch_step(gw, ref, n_total, cols, chains, V):
    """b9gw_lse_generated_step through caller-held device tensors, alternating with
    shard_partials + allreduce on the same comm (shared mailbox parity and step counters)."""
    import torch
    from base_b200 import vshards
    flat = gw.lse_generated(chains * n_total, cols, 0)["row_lse"].reshape(chains, n_total)
    wantP, want = ref.vshard_total(flat, V)
    dev = torch.device("cuda", 0)
    with vshards.PeerComm(0, 0, 1, V, max_chains=64) as comm:
        rows = torch.empty(chains, max(n_total, 1), dtype=torch.float64, device=dev)
        P = torch.empty(V, chains, dtype=torch.float64, device=dev)
        tot = torch.full((chains,), float("nan"), dtype=torch.float64, device=dev)
        work = torch.zeros(gw.lib().b9gw_lse_workspace_bytes(chains, V) // 4, dtype=torch.int32, device=dev)
        vals = torch.from_numpy(flat).to(dev)
        for rep in range(3):
            tot.fill_(float("nan"))
            comm.lse_generated_step(n_total, cols, chains, rows, P, tot, work)
            torch.cuda.synchronize()
            assert (bits(tot.cpu().numpy()) == bits(want)).all(), rep
            assert (bits(P.cpu().numpy()) == bits(wantP)).all(), rep
            assert not work.any()
            two = comm.allreduce(comm.shard_partials(vals, n_total)) if n_total else None
            if two is not None:
                assert (bits(two.cpu().numpy()) == bits(want)).all(), rep
        assert comm.status()["steps"] == (6 if n_total else 0)


def test_smoke_entry_point(gw):
    import __graft_entry__ as g
    g.smoke()
