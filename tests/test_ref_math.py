"""The CPU checker itself, pinned against exact arithmetic and scipy (no GPU)."""
from fractions import Fraction

import numpy as np
import pytest
from scipy.special import logsumexp

from tests._ref import mixed_err


def test_dfma_chain_is_correctly_rounded_fma(ref):
    # fma(x,a,b) = round(x*a+b): replay in exact rationals, round once per step
    a, b, iters = 1.0 - 2.0 ** -12, 2.0 ** -12, 64
    got = ref.dfma_lanes(a, b, iters)
    for lane in (0, 1, 17, 31):
        xs = [1.0 + 0.125 * j + lane * 2.0 ** -10 for j in range(8)]
        for _ in range(iters):
            xs = [float(Fraction(x) * Fraction(a) + Fraction(b)) for x in xs]
        s = xs[0]
        for x in xs[1:]:
            s += x
        assert got[lane] == s


def test_dfma_chain_converges_to_fixed_point(ref):
    a, b = 1.0 - 2.0 ** -12, 2.0 ** -12          # x* = b/(1-a) = 1 per chain, 8 chains
    assert np.allclose(ref.dfma_lanes(a, b, 1 << 17), 8.0, rtol=0, atol=1e-9)


@pytest.mark.parametrize("cols", [1, 2, 31, 32, 33, 1000, 4096])
def test_lse_orders_agree_with_scipy(ref, cols):
    rng = np.random.default_rng(cols)
    x = rng.normal(-50.0, 25.0, size=(64, cols))
    want = logsumexp(x, axis=1)
    for warp_order in (False, True):
        got = ref.lse_rows(x, warp_order)
        assert mixed_err(got, want) < 2e-14


def test_lse_edge_cases(ref):
    x = np.full((3, 40), -np.inf)
    x[1, 5] = -700.0                 # a single finite term, far below exp's underflow
    x[2, :] = 700.0                  # would overflow without the max shift
    for warp_order in (False, True):
        got = ref.lse_rows(x, warp_order)
        assert got[0] == -np.inf
        assert got[1] == -700.0
        assert got[2] == pytest.approx(700.0 + np.log(40.0), rel=1e-15)
    assert (ref.lse_rows(np.empty((4, 0)), True) == -np.inf).all()      # empty grid
    assert ref.lse_rows(np.empty((0, 7)), True).size == 0               # no stars


def test_reduction_order_effect_is_far_inside_1e_10(ref):
    # SURVEY.md §7 "hard parts": how far can a warp-tree order move a result vs a serial loop?
    rng = np.random.default_rng(0)
    x = rng.normal(-40.0, 12.0, size=(2000, 1024))
    s, w = ref.lse_rows(x, False), ref.lse_rows(x, True)
    assert np.max(np.abs(s - w)) < 2e-14          # absolute: a few ulp of the row max (~40)
    assert mixed_err(w, s) < 2e-14
    # ...but NOT small relative to a row value that happens to sit near zero:
    assert np.max(np.abs(s - w) / np.abs(s)) > 1e-13
    (_, tot_shards), tot_serial = ref.rows_total(s), ref.serial_sum(s)
    assert abs(tot_shards - tot_serial) / abs(tot_serial) < 1e-13


def test_rows_total_small_and_ragged(ref):
    P, t = ref.rows_total(np.empty(0))
    assert t == 0.0 and (P == 0.0).all()
    v = np.arange(1, 2050, dtype=np.float64)
    for V in (4, 64, 128):
        assert ref.rows_total(v, V)[1] == v.sum() == 2049 * 2050 / 2
    P, t = ref.rows_total(np.array([3.0, 5.0]), 8)       # fewer rows than shards: empty shards add +0
    assert t == 8.0 and sorted(P) == [0.0] * 6 + [3.0, 5.0]


def test_generator_is_the_stated_arithmetic_rounded_once_per_operation(ref):
    # replay include/b9_groundwork.h's formula in exact rationals, rounding where it says
    import math
    phi = 0.6180339887498949
    for rows, cols in [(5, 7), (40, 1024), (3, 2500)]:
        x = ref.generate_terms(rows, cols)
        for r in (0, 1, rows - 1):
            u = float(Fraction(r) * Fraction(phi))
            c0 = float(Fraction(float(Fraction(u) - math.floor(u))) * cols)
            w = float(Fraction(34 + r % 7) * Fraction(float(Fraction(1) / cols)))
            b = -float(Fraction(c0) * Fraction(w))
            for c in (0, 1, cols // 2, cols - 1):
                t = float(Fraction(c) * Fraction(w) + Fraction(b))        # fma: one rounding
                assert x[r, c] == -float(Fraction(t) * Fraction(t))
    x = ref.generate_terms(200, 1024)
    assert (x.max(axis=1) > -1.0).all() and x.max() <= 0.0                 # a term near 0 in every row
    assert x.min() < -745.0 and (x.min(axis=1) < -280.0).all()             # some rows reach past underflow
    # a few hundred terms per row carry the sum, the rest are negligible or underflow
    carrying = (x > x.max(axis=1, keepdims=True) - 36.0).sum(axis=1)
    assert 150 < carrying.min() and carrying.max() < 400      # fewer when the peak sits at an edge


def test_shard_partial_and_total_against_exact_sums(ref):
    rng = np.random.default_rng(3)
    v = rng.normal(size=(4, 1003)) * 10.0 ** rng.integers(-6, 6, size=(4, 1003))
    P, T = ref.vshard_total(v, 64)
    lo = [ref.shard_lo(1003, 64, s) for s in range(65)]
    assert lo[0] == 0 and lo[64] == 1003 and all(b - a in (15, 16) for a, b in zip(lo, lo[1:]))
    import math
    for c in range(4):
        for s in (0, 17, 63):
            exact = math.fsum(v[c, lo[s]:lo[s + 1]])
            assert abs(P[s, c] - exact) <= 4e-16 * np.abs(v[c, lo[s]:lo[s + 1]]).sum()
        acc = 0.0
        for s in range(64):
            acc += P[s, c]
        assert T[c] == acc                                                  # strictly left to right
    # integers: every order is exact, so the partition itself is what is checked
    ints = np.arange(1, 1004, dtype=np.float64)[None, :]
    _, t = ref.vshard_total(ints, 8)
    assert t[0] == 1003 * 1004 / 2


def test_spread_arguments_are_bit_assembled_as_documented(ref):
    for which, lo, hi, neg in (("exp_spread", 2.0 ** -6, 1024.0, True), ("log_spread", 2.0 ** -8, 256.0, False)):
        a = ref.spread_args(which, 3, 1, 1 << 14)
        assert ((a < 0) == neg).all()
        m = np.abs(a)
        assert m.min() >= lo and m.max() < hi
        e = np.floor(np.log2(m)).astype(int)
        counts = np.bincount(e - e.min(), minlength=16)
        assert len(counts) == 16 and counts.min() > 0.7 * len(a) / 16       # 16 octaves, roughly uniform
        assert (a.view(np.uint64) & np.uint64(0xFFFFFFFF) == 0).all()       # low word is zero
