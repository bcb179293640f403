"""The C-ABI library loads and exports every symbol include/*.h declares (no GPU needed)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    names = set()
    for h in (ROOT / "include").glob("*.h"):
        names |= set(re.findall(r"\b(b9gw_\w+)\s*\(", h.read_text()))
    return names


def test_header_and_binding_declare_the_same_symbols():
    from base_b200 import groundwork as gw
    assert declared_symbols() == set(gw.SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    L = ctypes.CDLL(str(built.CUDA_LIB))
    for name in sorted(declared_symbols()):
        assert hasattr(L, name), name


def test_library_is_sm_100a_only(built):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", str(built.CUDA_LIB)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_abi_version_and_loud_failure_without_a_device(built):
    from base_b200 import groundwork as gw
    assert gw.lib().b9gw_abi_version() == gw.ABI_VERSION
    if gw.device_count() > 0:
        pytest.skip("a CUDA device is visible; the no-device path cannot be exercised")
    import numpy as np
    for call in (lambda: gw.dfma_peak(0), lambda: gw.transcendental_rate("exp"),
                 lambda: gw.transcendental_rate("exp_spread"),
                 lambda: gw.device_map("exp10", np.zeros(4)), lambda: gw.lse_rows(np.zeros((2, 2))),
                 lambda: gw.generate_terms(2, 2), lambda: gw.lse_generated(2, 2),
                 lambda: gw.vshard_total(np.zeros((2, 8)), 4)):
        with pytest.raises(gw.GroundworkError) as ei:
            call()
        assert ei.value.code == -1 and "no CPU fallback" in str(ei.value)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from base_b200 import groundwork as gw
    monkeypatch.setattr(gw, "_lib", None)
    monkeypatch.setattr(gw, "LIB_PATH", tmp_path / "libb9_groundwork.so")
    with pytest.raises(FileNotFoundError, match="no CPU fallback"):
        gw.lib()


def test_sizes_that_would_wrap_are_rejected_before_any_allocation(built):
    # rows*cols and n*8 used to be computed unchecked
    import ctypes as C
    from base_b200 import groundwork as gw
    L, t = gw.lib(), C.c_double()
    big = 1 << 62
    assert L.b9gw_lse_rows(0, None, big, 4, 64, 0, 1, None, None, C.byref(t), None) == gw.E_ARG
    assert L.b9gw_lse_generated(0, big, big, 64, 0, 1, None, None, C.byref(t), None) == gw.E_ARG
    assert L.b9gw_lse_generated(0, 8, 8, 48, 0, 1, None, None, C.byref(t), None) == gw.E_ARG
    assert L.b9gw_generate_terms(0, big, 8, None) == gw.E_ARG
    assert L.b9gw_map(0, 0, None, None, big) == gw.E_ARG
    assert L.b9gw_map(0, 7, None, None, 0) == gw.E_ARG
    assert b"overflow" in L.b9gw_last_error() or b"which" in L.b9gw_last_error()


def test_no_source_file_names_a_closed_batched_copy_call():
    # B200_PROFILING.md: gpurun refuses a repo whose sources name one of the four calls
    import re
    pat = re.compile("Memcpy" + "(3D)?" + "Batch" + "Async")
    for path in ROOT.rglob("*"):
        if path.is_file() and path.suffix in {".cu", ".cuh", ".h", ".c", ".cpp", ".py", ".sh"} \
                and ".git" not in path.parts and "gpurun_out" not in path.parts:
            assert not pat.search(path.read_text(errors="ignore")), path


def test_product_package_never_touches_the_checker():
    for py in (ROOT / "base_b200").rglob("*.py"):
        text = py.read_text()
        if py.name == "build.py":
            continue  # building the checker is not using it
        assert "groundwork_ref" not in text and "tests._ref" not in text and "from tests" not in text, py


@pytest.mark.parametrize("kernel", ["lse_staged_kernelILi0ELb0", "lse_staged_kernelILi1ELb0",
                                    "lse_staged_kernelILi1ELb1"])
def test_pass2_exp_chains_stay_interleaved_in_the_sass(built, kernel):
    """The four exps a lane evaluates per pass-2 iteration must be interleaved in the SASS (one
    dependent DFMA chain per warp cannot feed the FP64 pipe); ptxas once serialised them after an
    unrelated change, costing 5-9 % (profiles/r02b_groundwork.md).  No DFMA may be followed by more
    than one DFMA writing the same register, and all four chains' Horner steps must be present."""
    import sys
    sys.path.insert(0, str(ROOT / "tools"))
    from sass_interleave import fast_path_runs
    runs = fast_path_runs(built.CUDA_LIB, kernel)
    assert sum(n for _, n in runs) == 4 * 14, runs          # 14 DFMAs per exp (rounding fma .. last Horner step)
    assert max(n for _, n in runs) <= 2, runs


def _resource_usage(lib):
    """{mangled kernel name: {"REG": n, "STACK": n, "SHARED": n, "LOCAL": n}} from cuobjdump -res-usage."""
    import re
    import subprocess
    text = subprocess.run(["cuobjdump", "-res-usage", str(lib)], capture_output=True, text=True, check=True).stdout
    out, name = {}, None
    for line in text.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
        elif name and "REG:" in line:
            out[name] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line)}
            name = None
    return out


def test_kernels_fit_the_occupancy_their_launch_shapes_assume(built):
    """Register and stack budgets read from the built library on the CPU box, so a compiler or source
    change that spills, or that no longer fits the CTAs per SM the launch code counts on, fails here
    and not as a slower (or unlaunchable) kernel on the GPU.  65 536 registers per SM."""
    res = _resource_usage(built.CUDA_LIB)
    staged = {k: v for k, v in res.items() if "lse_staged_kernel" in k}
    assert len(staged) >= 3
    for name, r in staged.items():
        assert r["STACK"] == 0 and r["LOCAL"] == 0, (name, r)           # nothing spilled
        peer = "ELb1E" in name
        ctas = 6 if peer else 8                                         # __launch_bounds__(128, 8 | 6) in lse.cu
        assert r["REG"] * 128 * ctas <= 65536, (name, r)
    steps = {k: v for k, v in res.items() if "vshard_step_kernel" in k}
    assert len(steps) == 6                                              # V = 4 .. 128
    for name, r in steps.items():
        assert r["REG"] * 256 <= 65536, (name, r)                       # one 256-thread CTA must launch
        if "ILi32E" not in name and "ILi8E" not in name:                # V = 128 keeps 32 packets + 32 values per lane
            assert r["STACK"] == 0, (name, r)
    v64 = next(v for k, v in steps.items() if "ILi16E" in k)            # the default V = 64
    assert v64["STACK"] == 0 and v64["REG"] <= 168                      # >= 1 CTA of 256 threads with room to spare
    for name, r in res.items():
        if "dfma_peak_kernel" in name or "rate_kernel" in name:
            assert r["STACK"] == 0 and r["REG"] * 256 * 4 <= 65536, (name, r)   # >= 32 warps per SM: 0.99 of peak at ILP 8
