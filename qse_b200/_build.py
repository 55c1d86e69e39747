"""Build recipe for libqsb (in-tree, sm_100a only): ``python -m qse_b200._build``.

nvcc cross-compiles without a GPU; the resulting ``qse_b200/libqsb.so`` is git-ignored but
travels with the working tree to the GPU box.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libqsb.so")
SOURCES = [os.path.join(CSRC, "qsb.cu")]
HEADERS = [os.path.join(CSRC, f) for f in ("common.cuh", "hpsi.cuh", "hpsi_ws.cuh", "small.cuh", "observe.cuh", "terms.cuh", "nccl_dl.h")] + [
    os.path.join(os.path.dirname(HERE), "include", "qsb.h")
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libqsb can only be built with the CUDA 12.9 toolkit")
    return nvcc


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > built for p in SOURCES + HEADERS)


def build_variant(out: str, defines: list[str]) -> str:
    """Developer A/B builds: same sources with extra -D macros into another .so (see tools/)."""
    cmd = [find_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", out, *SOURCES, "-ldl"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed")
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libqsb.so if missing or older than its sources; returns the library path."""
    if not force and not is_stale():
        return LIB
    cmd = [find_nvcc(), *NVCC_FLAGS, "-o", LIB, *SOURCES, "-ldl"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}): {' '.join(cmd)}")
    log = os.path.join(HERE, "libqsb.ptxas.log")
    with open(log, "w") as fh:
        fh.write(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
