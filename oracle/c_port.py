"""ctypes loader for the C restatement of the oracle (oracle/c/rydberg_ref.c).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this.  Builds the .so with gcc on first use if missing.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "c", "librydberg_ref.so")
_dp = C.POINTER(C.c_double)
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(HERE, "c", "rydberg_ref.c")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(HERE, "c"), "-s", "-B"], check=True)
    lib = C.CDLL(SO)
    lib.ref_threads.restype = C.c_int
    lib.ref_set_threads.argtypes = [C.c_int]
    lib.ref_build_diagonal.argtypes = [C.c_int, _dp, _dp]
    lib.ref_apply_h.argtypes = [C.c_int, _dp, C.c_double, C.c_double, _dp, _dp]
    lib.ref_evolve_schedule.restype = C.c_long
    lib.ref_evolve_schedule.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_long, _dp, C.c_double, C.c_double]
    lib.ref_spins.argtypes = [C.c_int, _dp, _dp]
    lib.ref_number.argtypes = [C.c_int, _dp, _dp]
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(_dp)


def threads() -> int:
    return load().ref_threads()


def set_threads(n: int):
    load().ref_set_threads(int(n))


def build_diagonal(v: np.ndarray, n: int) -> np.ndarray:
    v = np.ascontiguousarray(v, dtype=np.float64)
    e = np.empty(1 << n, dtype=np.float64)
    load().ref_build_diagonal(n, _p(v), _p(e))
    return e


def apply_h(psi: np.ndarray, e: np.ndarray, n: int, omega: float, delta: float) -> np.ndarray:
    x = np.ascontiguousarray(psi, dtype=np.complex128).reshape(-1)
    y = np.empty_like(x)
    load().ref_apply_h(n, _p(e), float(omega), float(delta), _p(x.view(np.float64)), _p(y.view(np.float64)))
    return y


def evolve_schedule(psi: np.ndarray, e: np.ndarray, n: int, omega, delta, dt, tol=1e-15, theta_max=3.0):
    """In place on a copy; returns (psi, n_hpsi)."""
    x = np.array(psi, dtype=np.complex128).reshape(-1)
    omega = np.ascontiguousarray(omega, dtype=np.float64)
    delta = np.ascontiguousarray(delta, dtype=np.float64)
    dt = np.ascontiguousarray(dt, dtype=np.float64)
    count = load().ref_evolve_schedule(n, _p(e), _p(omega), _p(delta), _p(dt), omega.size,
                                       _p(x.view(np.float64)), float(tol), float(theta_max))
    if count < 0:
        raise RuntimeError(f"C oracle: evolve failed ({count})")
    return x, int(count)


def spins(psi: np.ndarray, n: int) -> np.ndarray:
    x = np.ascontiguousarray(psi, dtype=np.complex128).reshape(-1)
    out = np.empty((n, 3), dtype=np.float64)
    load().ref_spins(n, _p(x.view(np.float64)), _p(out))
    return out


def number(psi: np.ndarray, n: int) -> np.ndarray:
    x = np.ascontiguousarray(psi, dtype=np.complex128).reshape(-1)
    out = np.empty(n, dtype=np.float64)
    load().ref_number(n, _p(x.view(np.float64)), _p(out))
    return out
