"""ctypes binding of libb9_groundwork.so (include/b9_groundwork.h).

No fallback of any kind: if the library is missing this module raises on first
use, and on a box without a CUDA device every compute call raises
`GroundworkError` carrying the library's own message.  The CPU checker in
`oracle/` is never imported from here.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "libb9_groundwork.so"
ABI_VERSION = 5
DFMA_ILP, TRANS_ILP, THREADS = 8, 4, 256
LSE_STAGED_COLS, MAX_WORLD, MAX_VSHARDS, IPC_HANDLE_BYTES = 1024, 16, 128, 64
LSE_MAX_CHAINS = 65535
E_NODEVICE, E_CUDA, E_ARG, E_TIMEOUT, E_STATE = -1, -2, -3, -4, -5

# every symbol include/b9_groundwork.h declares: name -> (restype, argtypes)
_i, _ll, _d, _f = C.c_int, C.c_longlong, C.c_double, C.c_float
_pd, _pi, _pll, _pf = C.POINTER(_d), C.POINTER(_i), C.POINTER(_ll), C.POINTER(_f)
_vp, _ull = C.c_void_p, C.c_ulonglong
SYMBOLS = {
    "b9gw_abi_version": (_i, []),
    "b9gw_last_error": (C.c_char_p, []),
    "b9gw_device_count": (_i, []),
    "b9gw_device_info": (_i, [_i, _pi, _pi, _pll]),
    "b9gw_dfma_peak": (_i, [_i, _i, _i, _i, _i, _d, _d, _i, _i, _pd, _pll, _pf, _pd]),
    "b9gw_transcendental_rate": (_i, [_i, _i, _i, _i, _i, _i, _pd, _pll, _pf, _pd]),
    "b9gw_step_latency": (_i, [_i, _i, _i, _pf, _pf, _pf]),
    "b9gw_map": (_i, [_i, _i, _pd, _pd, _ll]),
    "b9gw_lse_rows": (_i, [_i, _pd, _ll, _ll, _i, _i, _i, _pd, _pd, _pd, _pf]),
    "b9gw_generate_terms": (_i, [_i, _ll, _ll, _pd]),
    "b9gw_lse_generated": (_i, [_i, _ll, _ll, _i, _i, _i, _pd, _pd, _pd, _pf]),
    "b9gw_lse_workspace_bytes": (_ll, [_ll, _i]),
    "b9gw_lse_generated_shards": (_i, [_i, _ll, _ll, _ll, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "b9gw_dev_malloc": (_i, [_i, _ll, C.POINTER(_vp)]),
    "b9gw_dev_free": (_i, [_i, _vp]),
    "b9gw_memcpy_h2d": (_i, [_i, _vp, _vp, _ll]),
    "b9gw_memcpy_d2h": (_i, [_i, _vp, _vp, _ll]),
    "b9gw_vshard_bounds": (_i, [_ll, _i, _i, _pll, _pll]),
    # *_dev arguments are raw device addresses (c_void_p), e.g. torch.Tensor.data_ptr()
    "b9gw_shard_partials": (_i, [_i, _vp, _ll, _ll, _ll, _i, _i, _i, _vp, _vp]),
    "b9gw_comm_create": (_i, [_i, _i, _i, _i, _ll, C.POINTER(_vp), _vp]),
    "b9gw_comm_connect": (_i, [_vp, _vp]),
    "b9gw_ordered_allreduce": (_i, [_vp, _vp, _vp, _ll, _vp]),
    "b9gw_comm_set_timeout_ms": (_i, [_vp, _i]),
    "b9gw_comm_status": (_i, [_vp, _pi, C.POINTER(_ull)]),
    "b9gw_allreduce_latency": (_i, [_vp, _ll, _i, _i, _pf, _pf]),
    "b9gw_lse_generated_step": (_i, [_vp, _ll, _ll, _ll, _vp, _vp, _vp, _vp, _vp]),
    "b9gw_sharded_step": (_i, [_vp, _ll, _ll, _ll, _i, _i, _pd, _pd, _pf, _pf, _pf]),
    "b9gw_comm_destroy": (_i, [_vp]),
    "b9gw_vshard_total": (_i, [_i, _pd, _ll, _ll, _i, _pd, _pd]),
}


class GroundworkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libb9_groundwork error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(
                f"{LIB_PATH} is not built; run `python -c 'import __graft_entry__ as g; g.build()'`"
                " (there is no CPU fallback)")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.b9gw_abi_version() != ABI_VERSION:
            raise RuntimeError("libb9_groundwork.so ABI version mismatch; rebuild")
        _lib = L
    return _lib


def _ck(rc: int) -> None:
    if rc != 0:
        raise GroundworkError(rc, lib().b9gw_last_error().decode())


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(_pd)


def device_count() -> int:
    return lib().b9gw_device_count()


def device_info(device: int = 0) -> dict:
    sm, mhz, l2 = _i(), _i(), _ll()
    _ck(lib().b9gw_device_info(device, C.byref(sm), C.byref(mhz), C.byref(l2)))
    return {"sm_count": sm.value, "sm_clock_mhz": mhz.value, "l2_bytes": l2.value}


def dfma_peak(device=0, ctas_per_sm=8, iters=1 << 16, a=1.0 - 2.0 ** -12, b=2.0 ** -12,
              warmup=3, reps=10, want_out=False, ilp=DFMA_ILP, int_per_fma=0) -> dict:
    n, ms, tf = _ll(), _f(), _d()
    out = None
    if want_out:
        out = np.empty(device_info(device)["sm_count"] * ctas_per_sm * THREADS, dtype=np.float64)
    _ck(lib().b9gw_dfma_peak(device, ctas_per_sm, ilp, int_per_fma, iters, a, b, warmup, reps, _ptr(out),
                             C.byref(n), C.byref(ms), C.byref(tf)))
    return {"n_threads": n.value, "ms_per_launch": ms.value, "tflops": tf.value, "out": out,
            "iters": iters, "ctas_per_sm": ctas_per_sm, "ilp": ilp, "int_per_fma": int_per_fma,
            "launches": warmup + reps}


TRANS_WHICH = {"exp": 0, "log": 1, "exp10": 2, "log10": 3, "exp_spread": 4, "log_spread": 5}
MAP_WHICH = {"exp": 0, "log": 1, "exp10": 2, "log10": 3, "exp_fast_path": 4}


def transcendental_rate(which: str, device=0, ctas_per_sm=8, iters=1 << 12, warmup=3, reps=10,
                        want_out=False) -> dict:
    w = TRANS_WHICH[which]
    n, ms, g = _ll(), _f(), _d()
    out = None
    if want_out:
        out = np.empty(device_info(device)["sm_count"] * ctas_per_sm * THREADS, dtype=np.float64)
    _ck(lib().b9gw_transcendental_rate(device, w, ctas_per_sm, iters, warmup, reps, _ptr(out),
                                       C.byref(n), C.byref(ms), C.byref(g)))
    return {"n_threads": n.value, "ms_per_launch": ms.value, "gevals_per_s": g.value, "out": out,
            "iters": iters, "ctas_per_sm": ctas_per_sm, "launches": warmup + reps}


def step_latency(device=0, warmup=50, reps=2000) -> dict:
    a, b, c = _f(), _f(), _f()
    _ck(lib().b9gw_step_latency(device, warmup, reps, C.byref(a), C.byref(b), C.byref(c)))
    return {"us_launch_sync": a.value, "us_launch_d2h_sync": b.value, "us_graph_d2h_sync": c.value,
            "launches": 3 * (warmup + reps)}


def device_map(which: str, x: np.ndarray, device=0) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    _ck(lib().b9gw_map(device, MAP_WHICH[which], _ptr(x), _ptr(y), x.size))
    return y


def lse_rows(x: np.ndarray, device=0, warmup=0, reps=1, n_vshards=64) -> dict:
    x = np.ascontiguousarray(x, dtype=np.float64)
    if x.ndim != 2:
        raise ValueError("x must be rows x cols")
    rows, cols = x.shape
    row_lse = np.empty(rows, dtype=np.float64)
    partials = np.empty(n_vshards, dtype=np.float64)
    total, ms = _d(), _f()
    _ck(lib().b9gw_lse_rows(device, _ptr(x), rows, cols, n_vshards, warmup, reps, _ptr(row_lse),
                            _ptr(partials), C.byref(total), C.byref(ms)))
    return {"row_lse": row_lse, "partials": partials, "total": total.value,
            "ms_per_launch": ms.value, "launches": warmup + reps}


def generate_terms(rows: int, cols: int, device=0) -> np.ndarray:
    x = np.empty((rows, cols), dtype=np.float64)
    _ck(lib().b9gw_generate_terms(device, rows, cols, _ptr(x)))
    return x


def lse_generated(rows: int, cols: int, device=0, warmup=0, reps=1, n_vshards=64) -> dict:
    row_lse = np.empty(rows, dtype=np.float64)
    partials = np.empty(n_vshards, dtype=np.float64)
    total, ms = _d(), _f()
    _ck(lib().b9gw_lse_generated(device, rows, cols, n_vshards, warmup, reps, _ptr(row_lse),
                                 _ptr(partials), C.byref(total), C.byref(ms)))
    return {"row_lse": row_lse, "partials": partials, "total": total.value,
            "ms_per_launch": ms.value, "launches": warmup + reps}


class DeviceBuffer:
    """b9gw_dev_malloc'd (zero-filled) device memory, for callers that hold no CUDA runtime."""

    def __init__(self, nbytes: int, device=0):
        self.device, self.nbytes, self.ptr = device, nbytes, _vp()
        _ck(lib().b9gw_dev_malloc(device, nbytes, C.byref(self.ptr)))

    def to_host(self, dtype, count: int) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        if out.nbytes:
            _ck(lib().b9gw_memcpy_d2h(self.device, out.ctypes.data_as(_vp), self.ptr, out.nbytes))
        return out

    def free(self) -> None:
        if self.ptr:
            lib().b9gw_dev_free(self.device, self.ptr)
            self.ptr = _vp()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.free()


def lse_generated_shards(n_stars_total: int, cols: int, chains: int, n_vshards: int,
                         first_shard: int, n_shards: int, device=0, launches=1) -> dict:
    """One rank's share of a star-sharded job (b9gw_lse_generated_shards), through device
    buffers of the library's own; `launches` > 1 re-runs it on the same workspace."""
    L = lib()
    lo = first_shard * n_stars_total // n_vshards
    hi = (first_shard + n_shards) * n_stars_total // n_vshards
    n_local = hi - lo
    ws = L.b9gw_lse_workspace_bytes(chains, n_shards)
    if ws < 0:
        _ck(int(ws))
    with DeviceBuffer(8 * chains * n_local, device) as rows, \
            DeviceBuffer(8 * n_shards * chains, device) as part, \
            DeviceBuffer(8 * chains, device) as tot, DeviceBuffer(ws, device) as work:
        for _ in range(launches):
            _ck(L.b9gw_lse_generated_shards(device, n_stars_total, cols, chains, n_vshards, first_shard,
                                            n_shards, rows.ptr, part.ptr, tot.ptr, work.ptr, None))
        out = {"row_lse": rows.to_host(np.float64, chains * n_local).reshape(chains, n_local),
               "partials": part.to_host(np.float64, n_shards * chains).reshape(n_shards, chains),
               "total": tot.to_host(np.float64, chains) if n_shards == n_vshards else None,
               "workspace": work.to_host(np.uint32, ws // 4), "launches": launches}
    return out


def vshard_bounds(n_stars: int, n_vshards: int, shard: int) -> tuple[int, int]:
    lo, hi = _ll(), _ll()
    _ck(lib().b9gw_vshard_bounds(n_stars, n_vshards, shard, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def vshard_total(values: np.ndarray, n_vshards: int, device=0) -> dict:
    """World = 1, host buffers: per-shard partials [V, chains] and total [chains]."""
    values = np.ascontiguousarray(values, dtype=np.float64)
    if values.ndim != 2:
        raise ValueError("values must be chains x stars")
    chains, n = values.shape
    partials = np.empty((n_vshards, chains), dtype=np.float64)
    total = np.empty(chains, dtype=np.float64)
    _ck(lib().b9gw_vshard_total(device, _ptr(values), chains, n, n_vshards, _ptr(partials),
                                _ptr(total)))
    return {"partials": partials, "total": total}
