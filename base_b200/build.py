"""Build recipes: nvcc for the sm_100a library, gcc for the CPU checker.

Both outputs are in-tree (git-ignored, not gpurun-ignored) so they travel to the
GPU box with the snapshot.  nvcc cross-compiles here without a GPU.
"""
from __future__ import annotations

import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CUDA_SRCS = tuple(PKG / "csrc" / n for n in ("groundwork.cu", "lse.cu", "vshard.cu"))
CUDA_HDRS = (PKG / "csrc" / "common.cuh", ROOT / "include" / "b9_groundwork.h")
CUDA_LIB = PKG / "libb9_groundwork.so"
REF_SRC = ROOT / "oracle" / "groundwork_ref.c"
REF_LIB = ROOT / "oracle" / "libb9_groundwork_ref.so"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-std=c++17", "-shared", "-Xcompiler", "-fPIC"]
GCC_FLAGS = ["-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wextra"]


def _stale(out: Path, *srcs: Path) -> bool:
    return not out.exists() or any(s.stat().st_mtime > out.stat().st_mtime for s in srcs)


def _run(cmd):
    r = subprocess.run([str(c) for c in cmd], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"build failed: {' '.join(map(str, cmd))}\n{r.stdout}\n{r.stderr}")


def build_cuda(force: bool = False) -> Path:
    if force or _stale(CUDA_LIB, *CUDA_SRCS, *CUDA_HDRS):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        _run([nvcc, *NVCC_FLAGS, f"-I{ROOT / 'include'}", "-o", CUDA_LIB, *CUDA_SRCS])
    return CUDA_LIB


def build_ref(force: bool = False) -> Path:
    if force or _stale(REF_LIB, REF_SRC):
        _run(["gcc", *GCC_FLAGS, "-o", REF_LIB, REF_SRC, "-lm"])
    return REF_LIB


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_ref(force)


if __name__ == "__main__":
    build_all(force=True)
    print(CUDA_LIB, REF_LIB, sep="\n")
