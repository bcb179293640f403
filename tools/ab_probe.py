"""A/B of two builds of libb9_groundwork.so on one box: b9gw_lse_generated alone, alternating.

    python tools/ab_probe.py rows cols reps rounds build/libb9_head.so base_b200/libb9_groundwork.so ...

Plain ctypes on purpose: the two builds may differ in ABI version, but this one signature
has not changed since ABI 3.
"""
import ctypes as C
import sys

rows, cols, reps, rounds = (int(a) for a in sys.argv[1:5])
paths = sys.argv[5:]
libs = [C.CDLL(p) for p in paths]
for L in libs:
    L.b9gw_lse_generated.argtypes = [C.c_int, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float)]
for r in range(rounds):
    line = []
    for p, L in zip(paths, libs):
        t, ms = C.c_double(), C.c_float()
        rc = L.b9gw_lse_generated(0, rows, cols, 64, 5, reps, None, None, C.byref(t), C.byref(ms))
        assert rc == 0, (p, rc)
        line.append(f"{p.split('/')[-1]} {ms.value * 1e3:8.2f} us")
    print(f"{rows}x{cols} round {r}: " + " | ".join(line) + f" (total {t.value!r})")
