"""ctypes binding of libqsb (include/qsb.h) — the only way Python reaches the GPU here.

There is no CPU fallback: if the shared library is missing or no sm_100 device is present,
every entry point raises.  :class:`Engine` is a thin object wrapper over one ``qsb_ctx``.
"""

from __future__ import annotations

import ctypes as C
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QSB_LIB") or os.path.join(HERE, "libqsb.so")  # QSB_LIB: developer A/B builds
NCCL_UID_BYTES = 128

_c_double_p = C.POINTER(C.c_double)
_c_u64_p = C.POINTER(C.c_uint64)
_ctx_p = C.c_void_p


class QsbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libqsb error {code}: {message}")
        self.code = code


class QsbEop(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("omega", C.c_double), ("delta", C.c_double),
                ("term_begin", C.c_int64), ("term_end", C.c_int64)]


class QsbEops(C.Structure):
    _fields_ = [("n_ops", C.c_int32), ("reserved", C.c_int32), ("ops", C.POINTER(QsbEop)), ("n_terms", C.c_int64),
                ("coef", _c_double_p), ("xmask", _c_u64_p), ("ymask", _c_u64_p), ("zmask", _c_u64_p),
                ("nmask", _c_u64_p)]


class QsbMetrics(C.Structure):
    _fields_ = [("hpsi_calls", C.c_uint64), ("kernel_launches", C.c_uint64), ("segments", C.c_uint64),
                ("substeps", C.c_uint64), ("last_theta", C.c_double), ("last_terms", C.c_int32),
                ("hpsi_passes", C.c_int32), ("hpsi_bytes_per_amp", C.c_double), ("diag_min", C.c_double),
                ("diag_max", C.c_double)]


# name -> (restype, argtypes); mirrors include/qsb.h one to one (tests/test_abi.py checks it)
SIGNATURES = {
    "qsb_abi_version": (C.c_int, []),
    "qsb_build_info": (C.c_char_p, []),
    "qsb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "qsb_last_error": (C.c_char_p, [_ctx_p]),
    "qsb_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "qsb_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_ctx_p)]),
    "qsb_create_sharded": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(_ctx_p)]),
    "qsb_destroy": (C.c_int, [_ctx_p]),
    "qsb_host_alloc": (C.c_int, [C.c_uint64, C.POINTER(C.c_void_p)]),
    "qsb_host_free": (C.c_int, [C.c_void_p]),
    "qsb_local_dim": (C.c_int, [_ctx_p, _c_u64_p]),
    "qsb_set_option": (C.c_int, [_ctx_p, C.c_char_p, C.c_double]),
    "qsb_set_interaction": (C.c_int, [_ctx_p, _c_double_p]),
    "qsb_get_diagonal": (C.c_int, [_ctx_p, _c_double_p]),
    "qsb_set_site_weights": (C.c_int, [_ctx_p, _c_double_p, _c_double_p]),
    "qsb_set_operator_terms": (C.c_int, [_ctx_p, C.c_int64, _c_double_p, _c_u64_p, _c_u64_p, _c_u64_p, _c_u64_p,
                                         C.POINTER(C.c_int32)]),
    "qsb_reset_state": (C.c_int, [_ctx_p]),
    "qsb_set_state": (C.c_int, [_ctx_p, _c_double_p]),
    "qsb_get_state": (C.c_int, [_ctx_p, _c_double_p]),
    "qsb_apply_h": (C.c_int, [_ctx_p, C.c_double, C.c_double]),
    "qsb_get_work": (C.c_int, [_ctx_p, _c_double_p]),
    "qsb_apply_h_to": (C.c_int, [_ctx_p, C.c_double, C.c_double, _c_double_p, _c_double_p]),
    "qsb_evolve_segment": (C.c_int, [_ctx_p, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_int)]),
    "qsb_evolve_schedule": (C.c_int, [_ctx_p, _c_double_p, _c_double_p, _c_double_p, C.c_int64,
                                      C.POINTER(QsbEops), _c_double_p]),
    "qsb_evolve_schedule_sites": (C.c_int, [_ctx_p, _c_double_p, _c_double_p, _c_double_p, C.c_int64,
                                            C.POINTER(QsbEops), _c_double_p]),
    "qsb_norm": (C.c_int, [_ctx_p, _c_double_p]),
    "qsb_energy": (C.c_int, [_ctx_p, C.c_double, C.c_double, _c_double_p]),
    "qsb_expect_pauli": (C.c_int, [_ctx_p, C.c_int64, _c_double_p, _c_u64_p, _c_u64_p, _c_u64_p, _c_u64_p,
                                   _c_double_p]),
    "qsb_spins": (C.c_int, [_ctx_p, _c_double_p]),
    "qsb_sisj": (C.c_int, [_ctx_p, _c_double_p]),
    "qsb_number": (C.c_int, [_ctx_p, _c_double_p]),
    "qsb_probabilities": (C.c_int, [_ctx_p, _c_double_p]),
    "qsb_topk_probs": (C.c_int, [_ctx_p, C.c_int, _c_u64_p, _c_double_p]),
    "qsb_sample": (C.c_int, [_ctx_p, C.c_uint64, C.c_int64, _c_u64_p]),
    "qsb_describe_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "qsb_describe_ticket": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_int),
                                       _c_u64_p]),
    "qsb_get_metrics": (C.c_int, [_ctx_p, C.POINTER(QsbMetrics)]),
    "qsb_time_apply_h": (C.c_int, [_ctx_p, C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "qsb_time_schedule": (C.c_int, [_ctx_p, _c_double_p, _c_double_p, _c_double_p, C.c_int64,
                                    C.POINTER(C.c_float)]),
}

_lib = None


def _point_at_nccl():
    """libqsb binds NCCL with dlopen; prefer the copy bundled with torch (2.28.x)."""
    if os.environ.get("QSB_NCCL_LIB"):
        return
    for base in sys.path:
        hits = glob.glob(os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2"))
        if hits:
            os.environ["QSB_NCCL_LIB"] = hits[0]
            return


def load(path: str | None = None):
    """Load libqsb.so and declare every prototype.  Raises if the library is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise QsbError(-100, f"{path} is missing: build it with `python -m qse_b200._build` "
                             "(nvcc, sm_100a). There is no CPU fallback.")
    _point_at_nccl()
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.qsb_abi_version() != 1:
        raise QsbError(-101, f"ABI version {lib.qsb_abi_version()} != 1")
    _lib = lib
    return lib


def device_count() -> int:
    n = C.c_int(0)
    rc = load().qsb_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def is_available() -> tuple[bool, str]:
    """(usable, reason) — the availability gate the Calculator constructor applies."""
    try:
        lib = load()
    except (OSError, QsbError, AttributeError) as exc:
        return False, str(exc)
    n = C.c_int(0)
    if lib.qsb_device_count(C.byref(n)) != 0 or n.value < 1:
        return False, "no CUDA device visible: " + lib.qsb_last_error(None).decode()
    return True, ""


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(NCCL_UID_BYTES)
    rc = load().qsb_nccl_unique_id(buf)
    if rc != 0:
        raise QsbError(rc, load().qsb_last_error(None).decode())
    return buf.raw


def describe_plan(n_local: int, sb_bits: int = 0, tile_bits: int = 12) -> dict:
    """Launch plan of one H psi (host-only): tile / super-block bits and the (lb, hb, hb0) tile
    geometry of the inner sub-pass and of every plain launch."""
    out = (C.c_int * 32)()
    rc = load().qsb_describe_plan(int(n_local), int(sb_bits), int(tile_bits), out)
    if rc != 0:
        raise QsbError(rc, load().qsb_last_error(None).decode())
    geo = lambda i: {"lb": out[i], "hb": out[i + 1], "hb0": out[i + 2]}  # noqa: E731
    return {"tile_bits": out[0], "sb_bits": out[1], "inner": geo(3), "plain": [geo(6 + 3 * i) for i in range(out[2])]}


def pinned_empty(n: int, dtype=np.complex128) -> np.ndarray:
    """1-D array of ``n`` elements in page-locked host memory (freed when the array is collected)."""
    import weakref

    lib = load()
    nbytes = int(n) * np.dtype(dtype).itemsize
    ptr = C.c_void_p()
    rc = lib.qsb_host_alloc(nbytes, C.byref(ptr))
    if rc != 0:
        raise QsbError(rc, lib.qsb_last_error(None).decode())
    buf = (C.c_char * nbytes).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(n))
    weakref.finalize(buf, lib.qsb_host_free, ptr)
    return arr


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_c_double_p)


def _up(a: np.ndarray):
    return a.ctypes.data_as(_c_u64_p)


class Engine:
    """One shard of a Rydberg state vector on one GPU (one ``qsb_ctx``)."""

    def __init__(self, n_qubits: int, device: int = 0, rank: int = 0, nranks: int = 1, nccl_uid: bytes | None = None):
        self._lib = load()
        self._ctx = _ctx_p()
        self.n = int(n_qubits)
        self.rank, self.nranks = int(rank), int(nranks)
        if nranks > 1:
            if nccl_uid is None or len(nccl_uid) != NCCL_UID_BYTES:
                raise ValueError("sharded engine needs the 128-byte NCCL unique id of rank 0")
            uid = C.create_string_buffer(nccl_uid, NCCL_UID_BYTES)
            rc = self._lib.qsb_create_sharded(self.n, device, rank, nranks, uid, C.byref(self._ctx))
        else:
            rc = self._lib.qsb_create(self.n, device, C.byref(self._ctx))
        if rc != 0:
            self._ctx = _ctx_p()
            raise QsbError(rc, self._lib.qsb_last_error(None).decode())
        d = C.c_uint64(0)
        self._check(self._lib.qsb_local_dim(self._ctx, C.byref(d)))
        self.local_dim = int(d.value)

    # -- plumbing -------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise QsbError(rc, self._lib.qsb_last_error(self._ctx).decode())

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.qsb_destroy(self._ctx)
            self._ctx = _ctx_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_option(self, key: str, value: float):
        self._check(self._lib.qsb_set_option(self._ctx, key.encode(), float(value)))

    # -- Hamiltonian / state ---------------------------------------------------------------
    def set_interaction(self, v: np.ndarray):
        v = np.ascontiguousarray(v, dtype=np.float64)
        if v.shape != (self.n, self.n):
            raise ValueError(f"V must be ({self.n}, {self.n})")
        self._check(self._lib.qsb_set_interaction(self._ctx, _dp(v)))

    def set_site_weights(self, w_omega=None, w_delta=None):
        """Per-site weights of the drive / detuning (README.md:93-95): Omega_q = w_omega[q] * Omega(t), ..."""
        def arr(w):
            if w is None:
                return None
            w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1)
            if w.size != self.n:
                raise ValueError(f"site weights must hold {self.n} values")
            return w
        wo, wd = arr(w_omega), arr(w_delta)
        self._check(self._lib.qsb_set_site_weights(self._ctx, _dp(wo) if wo is not None else None,
                                                   _dp(wd) if wd is not None else None))

    def set_operator_terms(self, terms):
        """terms: iterable of (coef, xmask, ymask, zmask, nmask[, channel]); channel 0 static, 1 x amplitude(t),
        2 x detuning(t).  An empty list clears the operator part of H."""
        terms = list(terms)
        coef = np.array([t[0] for t in terms] + [0.0], dtype=np.float64)
        masks = [np.array([int(t[i]) for t in terms] + [0], dtype=np.uint64) for i in (1, 2, 3, 4)]
        chan = np.array([int(t[5]) if len(t) > 5 else 0 for t in terms] + [0], dtype=np.int32)
        self._check(self._lib.qsb_set_operator_terms(self._ctx, len(terms), _dp(coef), *[_up(m) for m in masks],
                                                     chan.ctypes.data_as(C.POINTER(C.c_int32))))

    def diagonal(self) -> np.ndarray:
        out = np.empty(self.local_dim, dtype=np.float64)
        self._check(self._lib.qsb_get_diagonal(self._ctx, _dp(out)))
        return out

    def reset_state(self):
        self._check(self._lib.qsb_reset_state(self._ctx))

    def set_state(self, psi: np.ndarray):
        psi = np.ascontiguousarray(psi, dtype=np.complex128).reshape(-1)
        if psi.size != self.local_dim:
            raise ValueError(f"state shard must hold {self.local_dim} amplitudes")
        self._check(self._lib.qsb_set_state(self._ctx, psi.view(np.float64).ctypes.data_as(_c_double_p)))

    def get_state(self, out: np.ndarray | None = None, pinned: bool = False) -> np.ndarray:
        if out is None:
            big = pinned and self.local_dim >= (1 << 18)
            out = pinned_empty(self.local_dim) if big else np.empty(self.local_dim, dtype=np.complex128)
        self._check(self._lib.qsb_get_state(self._ctx, out.view(np.float64).ctypes.data_as(_c_double_p)))
        return out

    # -- H psi / evolution -----------------------------------------------------------------
    def apply_h(self, omega: float, delta: float) -> np.ndarray:
        self._check(self._lib.qsb_apply_h(self._ctx, float(omega), float(delta)))
        out = np.empty(self.local_dim, dtype=np.complex128)
        self._check(self._lib.qsb_get_work(self._ctx, out.view(np.float64).ctypes.data_as(_c_double_p)))
        return out

    def apply_h_to(self, psi: np.ndarray, omega: float, delta: float) -> np.ndarray:
        """H(omega, delta) applied to a caller-supplied shard; the evolved state stays untouched."""
        psi = np.ascontiguousarray(psi, dtype=np.complex128).reshape(-1)
        if psi.size != self.local_dim:
            raise ValueError(f"state shard must hold {self.local_dim} amplitudes")
        out = np.empty(self.local_dim, dtype=np.complex128)
        self._check(self._lib.qsb_apply_h_to(self._ctx, float(omega), float(delta),
                                             psi.view(np.float64).ctypes.data_as(_c_double_p),
                                             out.view(np.float64).ctypes.data_as(_c_double_p)))
        return out

    def evolve_segment(self, omega: float, delta: float, dt_us: float) -> int:
        n = C.c_int(0)
        self._check(self._lib.qsb_evolve_segment(self._ctx, float(omega), float(delta), float(dt_us), C.byref(n)))
        return n.value

    def evolve_schedule(self, omega, delta, dt_us, eops: "PackedEops | None" = None) -> np.ndarray | None:
        omega = np.ascontiguousarray(omega, dtype=np.float64)
        delta = np.ascontiguousarray(delta, dtype=np.float64)
        dt_us = np.ascontiguousarray(dt_us, dtype=np.float64)
        if not (omega.shape == delta.shape == dt_us.shape) or omega.ndim != 1:
            raise ValueError("omega, delta, dt_us must be 1-D arrays of equal length")
        steps = omega.size
        if eops is None or eops.n_ops == 0:
            self._check(self._lib.qsb_evolve_schedule(self._ctx, _dp(omega), _dp(delta), _dp(dt_us), steps, None, None))
            return None
        out = np.zeros((steps, eops.n_ops), dtype=np.float64)
        self._check(self._lib.qsb_evolve_schedule(self._ctx, _dp(omega), _dp(delta), _dp(dt_us), steps,
                                                  C.byref(eops.struct), _dp(out)))
        return out

    def evolve_schedule_sites(self, omega_sites, delta_sites, dt_us, eops: "PackedEops | None" = None):
        """Per-site schedule: omega_sites / delta_sites are (n_steps, N) arrays (row s = the values of step s)."""
        omega_sites = np.ascontiguousarray(omega_sites, dtype=np.float64)
        delta_sites = np.ascontiguousarray(delta_sites, dtype=np.float64)
        dt_us = np.ascontiguousarray(dt_us, dtype=np.float64)
        steps = dt_us.size
        if omega_sites.shape != (steps, self.n) or delta_sites.shape != (steps, self.n):
            raise ValueError(f"omega_sites and delta_sites must be ({steps}, {self.n})")
        if eops is None or eops.n_ops == 0:
            self._check(self._lib.qsb_evolve_schedule_sites(self._ctx, _dp(omega_sites), _dp(delta_sites), _dp(dt_us), steps,
                                                            None, None))
            return None
        out = np.zeros((steps, eops.n_ops), dtype=np.float64)
        self._check(self._lib.qsb_evolve_schedule_sites(self._ctx, _dp(omega_sites), _dp(delta_sites), _dp(dt_us), steps,
                                                        C.byref(eops.struct), _dp(out)))
        return out

    # -- observables -----------------------------------------------------------------------
    def norm(self) -> float:
        v = C.c_double(0)
        self._check(self._lib.qsb_norm(self._ctx, C.byref(v)))
        return v.value

    def energy(self, omega: float, delta: float) -> float:
        v = C.c_double(0)
        self._check(self._lib.qsb_energy(self._ctx, float(omega), float(delta), C.byref(v)))
        return v.value

    def expect_pauli(self, coef, xmask, ymask, zmask, nmask) -> np.ndarray:
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        masks = [np.ascontiguousarray(m, dtype=np.uint64) for m in (xmask, ymask, zmask, nmask)]
        out = np.zeros(coef.size, dtype=np.complex128)
        self._check(self._lib.qsb_expect_pauli(self._ctx, coef.size, _dp(coef), *[_up(m) for m in masks],
                                               out.view(np.float64).ctypes.data_as(_c_double_p)))
        return out

    def spins(self) -> np.ndarray:
        out = np.empty((self.n, 3), dtype=np.float64)
        self._check(self._lib.qsb_spins(self._ctx, _dp(out)))
        return out

    def sisj(self) -> np.ndarray:
        out = np.empty((self.n, self.n), dtype=np.float64)
        self._check(self._lib.qsb_sisj(self._ctx, _dp(out)))
        return out

    def number(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.float64)
        self._check(self._lib.qsb_number(self._ctx, _dp(out)))
        return out

    def probabilities(self) -> np.ndarray:
        out = np.empty(self.local_dim, dtype=np.float64)
        self._check(self._lib.qsb_probabilities(self._ctx, _dp(out)))
        return out

    def topk(self, k: int) -> tuple[np.ndarray, np.ndarray]:
        idx = np.empty(k, dtype=np.uint64)
        p = np.empty(k, dtype=np.float64)
        self._check(self._lib.qsb_topk_probs(self._ctx, int(k), _up(idx), _dp(p)))
        return idx, p

    def sample(self, shots: int, seed: int = 0) -> np.ndarray:
        """`shots` basis indices drawn from |psi|^2 (uint64, in draw order; same on every rank)."""
        out = np.empty(int(shots), dtype=np.uint64)
        self._check(self._lib.qsb_sample(self._ctx, int(seed) & (2**64 - 1), int(shots), _up(out)))
        return out

    # -- measurement -----------------------------------------------------------------------
    def metrics(self) -> dict:
        m = QsbMetrics()
        self._check(self._lib.qsb_get_metrics(self._ctx, C.byref(m)))
        return {name: getattr(m, name) for name, _ in QsbMetrics._fields_}

    def time_apply_h(self, omega: float, delta: float, iters: int, flush_l2: bool = False) -> np.ndarray:
        ms = (C.c_float * iters)()
        self._check(self._lib.qsb_time_apply_h(self._ctx, float(omega), float(delta), iters, int(flush_l2), ms))
        return np.array(ms[:], dtype=np.float64)

    def time_schedule(self, omega, delta, dt_us) -> np.ndarray:
        omega = np.ascontiguousarray(omega, dtype=np.float64)
        delta = np.ascontiguousarray(delta, dtype=np.float64)
        dt_us = np.ascontiguousarray(dt_us, dtype=np.float64)
        ms = (C.c_float * omega.size)()
        self._check(self._lib.qsb_time_schedule(self._ctx, _dp(omega), _dp(delta), _dp(dt_us), omega.size, ms))
        return np.array(ms[:], dtype=np.float64)


class PackedEops:
    """Flat arrays + struct for qsb_evolve_schedule's e_ops (keeps the numpy buffers alive)."""

    def __init__(self, ops):
        """ops: list of ("energy", omega, delta) or ("pauli", [(coef, xmask, ymask, zmask, nmask), ...])."""
        self.n_ops = len(ops)
        coef, xm, ym, zm, nm = [], [], [], [], []
        self._ops = (QsbEop * max(1, self.n_ops))()
        for i, op in enumerate(ops):
            if op[0] == "energy":
                self._ops[i] = QsbEop(0, 0, float(op[1]), float(op[2]), 0, 0)
            elif op[0] == "pauli":
                begin = len(coef)
                for c, x, y, z, n in op[1]:
                    coef.append(float(c)); xm.append(int(x)); ym.append(int(y)); zm.append(int(z)); nm.append(int(n))
                self._ops[i] = QsbEop(1, 0, 0.0, 0.0, begin, len(coef))
            else:
                raise ValueError(f"unknown e_op kind {op[0]!r}")
        self._coef = np.array(coef + [0.0], dtype=np.float64)
        self._masks = [np.array(m + [0], dtype=np.uint64) for m in (xm, ym, zm, nm)]
        self.struct = QsbEops(self.n_ops, 0, self._ops, len(coef), _dp(self._coef), *[_up(m) for m in self._masks])
