"""ctypes view of oracle/libb9_groundwork_ref.so — the CPU checker.

Test infrastructure only (see the header of oracle/groundwork_ref.c): imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by
base_b200/.
"""
import ctypes as C

import numpy as np

_d, _i, _ll = C.c_double, C.c_int, C.c_longlong
_pd = C.POINTER(_d)


class Ref:
    def __init__(self, path):
        L = C.CDLL(str(path))
        L.b9ref_dfma_lane.restype, L.b9ref_dfma_lane.argtypes = _d, [_i, _i, _d, _d, _i]
        L.b9ref_trans_lane.restype, L.b9ref_trans_lane.argtypes = _d, [_i, _i, _i]
        L.b9ref_map.restype, L.b9ref_map.argtypes = None, [_i, _pd, _pd, _ll]
        L.b9ref_lse_rows.restype, L.b9ref_lse_rows.argtypes = None, [_pd, _ll, _ll, _i, _pd]
        L.b9ref_serial_sum.restype, L.b9ref_serial_sum.argtypes = _d, [_pd, _ll]
        L.b9ref_spread_args.restype, L.b9ref_spread_args.argtypes = None, [_i, _i, _i, _i, _pd]
        L.b9ref_gen_term.restype, L.b9ref_gen_term.argtypes = _d, [_ll, _ll, _ll]
        L.b9ref_generate_terms.restype, L.b9ref_generate_terms.argtypes = None, [_ll, _ll, _pd]
        L.b9ref_shard_lo.restype, L.b9ref_shard_lo.argtypes = _ll, [_ll, _i, _i]
        L.b9ref_shard_partial.restype, L.b9ref_shard_partial.argtypes = _d, [_pd, _ll, _ll]
        L.b9ref_vshard_total.restype, L.b9ref_vshard_total.argtypes = None, [_pd, _ll, _ll, _i, _pd, _pd]
        self.L = L

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(_pd)

    def dfma_lanes(self, a, b, iters, ilp=8):
        return np.array([self.L.b9ref_dfma_lane(l, ilp, a, b, iters) for l in range(32)])

    TRANS = {"exp": 0, "log": 1, "exp10": 2, "log10": 3, "exp_spread": 4, "log_spread": 5}
    MAP = {"exp": 0, "log": 1, "exp10": 2, "log10": 3, "pow10": 12}

    def trans_lanes(self, which, iters):
        return np.array([self.L.b9ref_trans_lane(l, self.TRANS[which], iters) for l in range(32)])

    def spread_args(self, which, lane, j, iters):
        a = np.empty(iters, dtype=np.float64)
        self.L.b9ref_spread_args(lane, j, self.TRANS[which], iters, self._p(a))
        return a

    def map(self, which, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        self.L.b9ref_map(self.MAP[which], self._p(x), self._p(y), x.size)
        return y

    def generate_terms(self, rows, cols):
        x = np.empty((rows, cols), dtype=np.float64)
        self.L.b9ref_generate_terms(rows, cols, self._p(x))
        return x

    def shard_lo(self, n, V, v):
        return self.L.b9ref_shard_lo(n, V, v)

    def shard_partial(self, row, lo, hi):
        row = np.ascontiguousarray(row, dtype=np.float64)
        return self.L.b9ref_shard_partial(self._p(row), lo, hi)

    def vshard_total(self, values, V):
        values = np.ascontiguousarray(values, dtype=np.float64)
        chains, n = values.shape
        partials = np.empty((V, chains), dtype=np.float64)
        total = np.empty(chains, dtype=np.float64)
        self.L.b9ref_vshard_total(self._p(values), chains, n, V, self._p(partials), self._p(total))
        return partials, total

    def lse_rows(self, x, warp_order):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty(x.shape[0], dtype=np.float64)
        self.L.b9ref_lse_rows(self._p(x), x.shape[0], x.shape[1], int(warp_order), self._p(out))
        return out

    def rows_total(self, row_values, V=64):
        """(P[V], total) of a vector of row values: the virtual-shard sum with one chain."""
        P, t = self.vshard_total(np.asarray(row_values, dtype=np.float64)[None, :], V)
        return P[:, 0], t[0]

    def serial_sum(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        return self.L.b9ref_serial_sum(self._p(v), v.size)


def load(path):
    return Ref(path)


def ulp_distance(a, b):
    """Distance in units in the last place between two float64 arrays (same sign assumed
    except around zero, which the monotone integer mapping handles)."""
    ia = np.asarray(a, dtype=np.float64).view(np.int64).copy()
    ib = np.asarray(b, dtype=np.float64).view(np.int64).copy()
    ia[ia < 0] = np.int64(-(2 ** 63)) - ia[ia < 0]
    ib[ib < 0] = np.int64(-(2 ** 63)) - ib[ib < 0]
    return np.abs(ia - ib)


def mixed_err(got, want):
    """max |got-want| / max(1, |want|).  A row log-likelihood can sit arbitrarily close to 0,
    where a pure relative error is ill-conditioned (5e-15 absolute is 3e-12 relative at
    |lse| = 4e-4); the quantity compared downstream is the SUM over rows, of large magnitude."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    if got.size == 0:
        return 0.0
    return float(np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))))
