"""``qse_b200.B200`` — the B200-native state-vector Calculator (drop-in beside ``qse.calc.Qutip``).

Mirrors the reference plugin protocol (ICHEC/qse 1.1.21):

* constructor ``(qbits, amplitude, detuning)`` in Qutip's positional order, bare ``Signal``
  wrapped into ``Signals([signal])``, ``_check_pulses`` rules and messages
  (qse/calc/qutip.py:13-32, 75-88); ``qbits`` may be ``None`` and attached later as with
  ``Pulser`` (qse/calc/pulser.py:95-135); availability gate through
  ``Calculator.__init__(qbits, label, is_calculator_available, installation_message)``
  (qse/calc/calculator.py:86-101): a missing library / GPU raises ``Exception(message)``.
* ``calculate(e_ops=None)`` runs the loop of qse/calc/qutip.py:34-53 on the GPU and BOTH returns
  Qutip's dict ``{"state": (2^N, 1) complex128[, "expectations": (n_steps, n_ops)]}`` AND sets
  ``results`` / ``statevector`` / ``spins`` the way Pulser and Myqlm do
  (qse/calc/pulser.py:194-204) so the base-class getters work.
* ``get_spins`` / ``get_sij`` / ``structure_factor_from_sij`` override the base-class versions
  (qse/calc/calculator.py:125-187), which call the O(4^N) Python loops of qse/magnetic.py.
* ``get_hamiltonian(amp, det)`` returns a handle usable as an ``e_op`` (the notebook idiom of
  docs/use-cases/adiabatic.ipynb cell 6) instead of a QuTiP ``Qobj``.

Sharded (multi-GPU) use is SPMD: one process per GPU under ``torchrun``; every rank builds
the same calculator and calls the same methods; ``torch.distributed`` only carries the
bootstrap (NCCL id) and optional state gathers — the exchange of sharded qubits lives in
libqsb (peer memory over NVLink, NCCL for the scalar reductions).
"""

from __future__ import annotations

import time

import numpy as np

from . import host, lib

try:  # a real install of the reference: be a true subclass of its plugin base class
    from qse.calc.calculator import Calculator as _ReferenceCalculator  # type: ignore
except Exception:  # the GPU boxes: no qse / QuTiP / matplotlib
    _ReferenceCalculator = None

C6 = host.C6

INSTALLATION_MESSAGE = (
    "The B200 calculator needs libqsb.so (build: python -m qse_b200._build) and an sm_100 "
    "CUDA device; there is no CPU fallback."
)


class _MirrorCalculator:
    """Stand-in for qse.calc.calculator.Calculator (qse/calc/calculator.py:70-187)."""

    def __init__(self, qbits, label: str, is_calculator_available: bool, installation_message: str):
        if not is_calculator_available:
            raise Exception(installation_message)
        self._qbits = qbits
        self._label = label
        self.spins = None
        self.sij = None

    @property
    def label(self):
        return self._label

    @label.setter
    def label(self, label):
        self._label = label

    @property
    def qbits(self):
        return self._qbits

    @qbits.setter
    def qbits(self, qbits):
        self._qbits = qbits

    def calculate(self):
        raise NotImplementedError


_Base = _ReferenceCalculator or _MirrorCalculator


class Hamiltonian:
    """Handle for H(amplitude, detuning) of one calculator — what ``get_hamiltonian`` returns.

    Usable in ``calculate(e_ops=[...])`` (expectation of H after every step); ``diagonal()`` and
    ``apply(psi)`` give access to the matrix-free operator on the device.
    """

    def __init__(self, calc: "B200", amplitude: float, detuning: float):
        self.calc = calc
        self.amplitude = float(amplitude)
        self.detuning = float(detuning)

    def diagonal(self) -> np.ndarray:
        """diag(H) of this rank's shard: E[k] - detuning * sum_q w_q b_q(k) (E holds the interaction and any static
        diagonal operator strings; w_q are the per-site detuning weights, 1 by default)."""
        eng = self.calc._engine_ready()
        n = self.calc._nqbits()
        k = np.arange(eng.local_dim, dtype=np.uint64) + np.uint64(eng.rank * eng.local_dim)
        w = self.calc._site_weights()[1]
        pop = np.zeros(eng.local_dim, dtype=np.float64)
        for q in range(n):
            pop += w[q] * ((k >> np.uint64(n - 1 - q)) & np.uint64(1)).astype(np.float64)
        return self.detuning * (-1.0 * pop) + eng.diagonal()

    def apply(self, psi) -> np.ndarray:
        """H psi for a caller-supplied state (this rank's shard); the calculator's evolved state is untouched."""
        eng = self.calc._engine_ready()
        return eng.apply_h_to(np.asarray(psi, dtype=np.complex128).reshape(-1), self.amplitude, self.detuning)

    def __repr__(self):
        return f"Hamiltonian(amplitude={self.amplitude}, detuning={self.detuning})"


def _as_signals(pulse):
    if pulse is None:
        return None
    if host._is_signals(pulse):
        return pulse
    return host.Signals([pulse])


def _check_pulses(amplitude, detuning):
    """qse/calc/qutip.py:75-88 — same rules, same messages."""
    if amplitude.duration != detuning.duration:
        raise Exception("The amplitude and detuning must have the same duration.")
    if len(amplitude) != len(detuning):
        raise Exception("The amplitude and detuning must contain the same number of signals.")
    for a, d in zip(amplitude, detuning):
        if len(a) != len(d):
            raise Exception("The amplitude and detuning must have the same amount of values.")


def _nano_to_micro(t):
    return t / 1000


def interaction_matrix(positions, c6: float = C6, tol: float = 1e-8) -> np.ndarray:
    """V[i, j] (i<j) = c6 / |R_i - R_j|**6, kept iff |V| > tol.

    The host keeps the reference's exact expression — ``lambda d: self.c6 / d**6``
    (qse/calc/qutip.py:61-63) through the pair loop and cut-off of qse/qbits.py:1296-1304 with
    the distance of qse/qbits.py:1008 — so the coefficients are bit-identical; only the
    expansion to the 2^N diagonal happens on the GPU (qsb_set_interaction).
    """
    positions = np.asarray(positions, dtype=float)
    n = positions.shape[0]
    v = np.zeros((n, n), dtype=np.float64)
    for i in range(n - 1):
        for j in range(i + 1, n):
            coef = c6 / np.linalg.norm(positions[i] - positions[j]) ** 6
            if np.abs(coef) > tol:
                v[i, j] = coef
    return v


def flatten_schedule(amplitude, detuning):
    """(omega[], delta[], dt_us[]) per signal value, in the order of qse/calc/qutip.py:40-42.

    dt is taken from the AMPLITUDE segment only (``amp_s.time_per_value()``, :41).
    """
    om, de, dt = [], [], []
    for amp_s, det_s in zip(amplitude, detuning):
        step = _nano_to_micro(amp_s.time_per_value())
        for a, d in zip(amp_s.values, det_s.values):
            om.append(float(a))
            de.append(float(d))
            dt.append(step)
    return (np.array(om, dtype=np.float64), np.array(de, dtype=np.float64), np.array(dt, dtype=np.float64))


def _pauli_masks(operator, indices, nqbits):
    """(xmask, ymask, zmask, nmask) of one Pauli string; qubit q <-> bit N-1-q."""
    masks = {"X": 0, "Y": 0, "Z": 0, "N": 0}
    seen = 0
    for op, q in zip(operator, indices):
        if op not in masks:
            raise Exception("An operator must be one of 'X', 'Y', 'Z' or 'N'.")
        bit = 1 << (nqbits - 1 - int(q))
        if seen & bit:
            raise Exception("A qubit may appear only once in an operator string.")
        seen |= bit
        masks[op] |= bit
    return masks["X"], masks["Y"], masks["Z"], masks["N"]


def operator_terms(operators, nqbits: int, channel: int = 0):
    """``qse.Operators`` / ``qse.Operator`` -> [(coef, xmask, ymask, zmask, nmask, channel), ...] for
    ``Engine.set_operator_terms`` (qubit q <-> index bit N-1-q, Operator.to_qutip order:
    qse/operator/operator.py:80-92)."""
    ops = operators.operator_list if hasattr(operators, "operator_list") else [operators]
    out = []
    for t in ops:
        if int(t.nqbits) != int(nqbits):
            raise Exception("The operators must act on the calculator's number of qubits.")
        out.append((float(t.coef), *_pauli_masks(t.operator, t.indices, nqbits), int(channel)))
    return out


def evolve_operators(hamiltonian, t_final: float, psi0=None, nqbits: int | None = None, device: int = 0, tol: float = 1e-14):
    """psi(t_final) = exp(-i H t_final) psi0 for a static operator-sum Hamiltonian -- the
    ``qutip.sesolve(hamiltonian.to_qutip(), psi0, tlist=[0, t_final])`` idiom of docs/use-cases/time_evo.ipynb,
    on one GPU.  ``psi0`` defaults to |0...0>; returns the flat (2^N,) complex128 state."""
    n = int(nqbits if nqbits is not None else hamiltonian.nqbits)
    with lib.Engine(n, device=device) as eng:
        eng.set_option("tol", tol)
        eng.set_interaction(np.zeros((n, n)))
        eng.set_operator_terms(operator_terms(hamiltonian, n))
        if psi0 is None:
            eng.reset_state()
        else:
            eng.set_state(np.asarray(psi0, dtype=np.complex128).reshape(-1))
        eng.evolve_segment(0.0, 0.0, float(t_final))
        return eng.get_state()


class B200(_Base):
    """B200-native state-vector calculator for the Rydberg Hamiltonian

        H = (Omega/2) sum_i X_i - delta sum_i n_i + sum_{i<j} C6/|R_i-R_j|^6 n_i n_j

    and its generalisations: per-site drive and detuning ``Omega_i = w_i Omega(t)``,
    ``delta_i = v_i delta(t)`` (README.md:93-95) and extra operator terms from ``qse.Operators`` (static, or scaled
    by the amplitude / detuning signal: the TFIM of docs/use-cases/time_evo.ipynb, XY couplings, ...).

    Parameters
    ----------
    qbits : qse.Qbits | qse_b200.Qbits | None
        Anything with ``positions`` (N x d) and ``nqbits``; may be attached later.
    amplitude, detuning : Signal | Signals
        Piecewise-constant pulses (ns; values in rad/us).
    device : int, optional
        CUDA device (default: ``LOCAL_RANK`` under torchrun, else 0).
    method : {"taylor", "rk4"}
        "taylor": norm-terminated Taylor propagator of each segment (exact to ``tol``);
        "rk4": its 4-term truncation (classical RK4, one H psi chain per segment).
    tol : float
        Termination threshold on the norm of a Taylor term (default 1e-14).
    site_amplitude, site_detuning : array_like (N,), optional
        Per-site weights of the amplitude / detuning signal (default: 1 for every site, the global pulse).
    hamiltonian : qse.Operators, optional
        Operator terms added to H: ``hamiltonian`` itself is static; ``hamiltonian_amplitude`` /
        ``hamiltonian_detuning`` are multiplied by the amplitude / detuning signal value of the step (so
        ``amp * H_amp + det * H_det + H_static`` generalises qse/calc/qutip.py:55-58).
    interaction : bool
        ``False`` drops the C6/r^6 term (a pure operator Hamiltonian on the register).
    distributed : bool | None
        Shard over ``torch.distributed`` ranks; ``None`` = when a process group is up.
    return_state : bool | "shard"
        ``False`` keeps the final state on the device (needed from ~28 qubits up); ``"shard"`` (sharded
        runs) returns only this rank's shard, ``(2^(N-p), 1)``, instead of gathering the whole register
        on every rank.
    wtimes : bool
        Print wall-clock timings like the Pulser / Myqlm calculators do.
    """

    c6 = C6  # qse/calc/qutip.py:7, 11

    def __init__(self, qbits=None, amplitude=None, detuning=None, *, device=None, method="taylor", tol=1e-14,
                 distributed=None, return_state=True, compute_spins=True, wtimes=False, label="b200",
                 site_amplitude=None, site_detuning=None, hamiltonian=None, hamiltonian_amplitude=None,
                 hamiltonian_detuning=None, interaction=True):
        ok, why = lib.is_available()
        super().__init__(qbits=qbits, label=label, is_calculator_available=ok,
                         installation_message=INSTALLATION_MESSAGE + (" (" + why + ")" if why else ""))
        amplitude = _as_signals(amplitude)
        detuning = _as_signals(detuning)
        if amplitude is not None and detuning is not None:
            _check_pulses(amplitude, detuning)
        self.amplitude = amplitude
        self.detuning = detuning
        if method not in ("taylor", "rk4"):
            raise ValueError("method must be 'taylor' or 'rk4'")
        self.method = method
        self.tol = float(tol)
        self.device = device
        self.distributed = distributed
        self.return_state = return_state
        self.compute_spins = compute_spins
        self.wtimes = wtimes
        self.site_amplitude = site_amplitude
        self.site_detuning = site_detuning
        self.hamiltonian = hamiltonian
        self.hamiltonian_amplitude = hamiltonian_amplitude
        self.hamiltonian_detuning = hamiltonian_detuning
        self.interaction = bool(interaction)
        self.results = None
        self.metrics = None
        self._statevector = None
        self._engine = None
        self._engine_key = None

    # ---- attachment --------------------------------------------------------------------
    def set_qbits(self, qbits):
        """Hook called by ``Qbits(calculator=...)`` / ``qbits.calc = calc`` (qse/qbits.py:169-173)."""
        self._qbits = qbits

    def _nqbits(self) -> int:
        if self.qbits is None:
            raise Exception("No qbits attached to the calculator.")
        return int(self.qbits.nqbits)

    # ---- engine life cycle ---------------------------------------------------------------
    def _comm(self):
        """(rank, world, device) of this process."""
        rank, world = 0, 1
        use = self.distributed
        import os
        import sys

        # importing torch costs seconds: only look for a process group when one can exist
        maybe_distributed = "torch" in sys.modules or "WORLD_SIZE" in os.environ or "RANK" in os.environ
        if use or (use is None and maybe_distributed):
            try:
                import torch.distributed as dist

                if dist.is_available() and dist.is_initialized():
                    rank, world = dist.get_rank(), dist.get_world_size()
            except ImportError:
                pass
        if use and world == 1:
            raise Exception("distributed=True needs an initialised torch.distributed process group")
        if use is False:
            rank, world = 0, 1
        device = self.device
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) if world > 1 else 0
        return rank, world, device

    def _site_weights(self):
        n = self._nqbits()
        out = []
        for w in (self.site_amplitude, self.site_detuning):
            w = np.ones(n) if w is None else np.asarray(w, dtype=float).reshape(-1)
            if w.size != n:
                raise Exception("site_amplitude / site_detuning must hold one weight per qubit.")
            out.append(w)
        return out

    def _operator_terms(self):
        n = self._nqbits()
        terms = []
        for ops, chan in ((self.hamiltonian, 0), (self.hamiltonian_amplitude, 1), (self.hamiltonian_detuning, 2)):
            if ops is not None:
                terms += operator_terms(ops, n, chan)
        return terms

    def _engine_ready(self) -> lib.Engine:
        """(Re)build the device context when the register or the Hamiltonian changed (the reference goes stale)."""
        n = self._nqbits()
        positions = np.array(self.qbits.positions, dtype=float)
        w_amp, w_det = self._site_weights()
        terms = self._operator_terms()
        key = (n, positions.tobytes(), float(self.c6), w_amp.tobytes(), w_det.tobytes(), repr(terms), self.interaction)
        if self._engine is not None and self._engine_key == key:
            return self._engine
        if self._engine is not None:
            self._engine.close()
            self._engine = None
        rank, world, device = self._comm()
        uid = None
        if world > 1:
            from . import distributed as qdist

            uid = qdist.broadcast_unique_id()
        eng = lib.Engine(n, device=device, rank=rank, nranks=world, nccl_uid=uid)
        eng.set_interaction(interaction_matrix(positions, self.c6) if self.interaction else np.zeros((n, n)))
        if self.site_amplitude is not None or self.site_detuning is not None:
            eng.set_site_weights(w_amp, w_det)
        if terms:
            eng.set_operator_terms(terms)
        self._engine = eng
        self._engine_key = key
        self._statevector = None
        self.results = None
        return eng

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None
            self._engine_key = None

    # ---- the hot path ---------------------------------------------------------------------
    def get_hamiltonian(self, amplitude, detuning) -> Hamiltonian:
        """H(amplitude, detuning) as an e_op handle (reference: a Qobj, qse/calc/qutip.py:55-58)."""
        return Hamiltonian(self, amplitude, detuning)

    def _pack_eops(self, e_ops):
        n = self._nqbits()
        packed = []
        for op in e_ops:
            if isinstance(op, Hamiltonian):
                packed.append(("energy", op.amplitude, op.detuning))
            elif isinstance(op, tuple) and len(op) == 3 and op[0] in ("H", "energy"):
                packed.append(("energy", float(op[1]), float(op[2])))
            elif hasattr(op, "operator_list"):  # qse.Operators
                packed.append(("pauli", [(float(t.coef), *_pauli_masks(t.operator, t.indices, n))
                                         for t in op.operator_list]))
            elif hasattr(op, "operator") and hasattr(op, "indices"):  # qse.Operator
                packed.append(("pauli", [(float(op.coef), *_pauli_masks(op.operator, op.indices, n))]))
            else:
                raise TypeError("e_ops entries must be calc.get_hamiltonian(amp, det), ('H', amp, det), "
                                "or qse.Operator / qse.Operators with X/Y/Z/N strings")
        return lib.PackedEops(packed)

    # ---- checkpoint / resume (long sweeps at 30+ qubits; the reference has none: SURVEY.md section 5) ----
    def save_checkpoint(self, prefix: str, next_step: int):
        """Write this rank's shard of the current state to ``{prefix}.rank{r}.npy`` plus a JSON header.

        ``next_step`` is the index of the first schedule value that has NOT been applied yet; pass the
        same prefix to ``calculate(resume_from=prefix)`` to continue from there (same register, same
        number of ranks).  SPMD: every rank calls it."""
        import json

        eng = self._engine_ready()
        np.save(f"{prefix}.rank{eng.rank}.npy", eng.get_state())
        if eng.rank == 0:
            with open(f"{prefix}.json", "w") as fh:
                json.dump({"format": "qse_b200-checkpoint-1", "nqbits": self._nqbits(), "nranks": eng.nranks,
                           "next_step": int(next_step), "c6": float(self.c6),
                           "positions": np.asarray(self.qbits.positions, dtype=float).tolist()}, fh)

    def _load_checkpoint(self, prefix: str, eng) -> int:
        import json

        with open(f"{prefix}.json") as fh:
            head = json.load(fh)
        if head.get("format") != "qse_b200-checkpoint-1":
            raise Exception(f"{prefix}.json is not a qse_b200 checkpoint")
        if head["nqbits"] != self._nqbits() or head["nranks"] != eng.nranks:
            raise Exception("checkpoint was written for a different register size or number of ranks")
        if not np.array_equal(np.asarray(head["positions"], dtype=float), np.asarray(self.qbits.positions, dtype=float)):
            raise Exception("checkpoint was written for different qubit positions")
        eng.set_state(np.load(f"{prefix}.rank{eng.rank}.npy"))
        return int(head["next_step"])

    def calculate(self, e_ops=None, resume_from=None, stop_after=None):
        """Time-evolve |0...0> through the pulse schedule (qse/calc/qutip.py:34-53).

        ``resume_from`` (a ``save_checkpoint`` prefix) starts from the stored state at its
        ``next_step`` instead; ``stop_after=k`` applies only the schedule values before index k (then
        ``save_checkpoint(prefix, k)``).  Expectations are returned for the steps actually run."""
        if self.amplitude is None or self.detuning is None:
            raise Exception("The amplitude and detuning must be set before calculate().")
        _check_pulses(self.amplitude, self.detuning)
        t0 = time.time()
        eng = self._engine_ready()
        eng.set_option("tol", self.tol)
        eng.set_option("fixed_terms", 4 if self.method == "rk4" else 0)
        omega, delta, dt = flatten_schedule(self.amplitude, self.detuning)
        first = 0
        if resume_from is not None:
            first = self._load_checkpoint(resume_from, eng)
        else:
            eng.reset_state()
        last = omega.size if stop_after is None else int(stop_after)
        if not 0 <= first <= last <= omega.size:
            raise Exception(f"cannot run schedule values [{first}, {last}) of {omega.size}")
        omega, delta, dt = omega[first:last], delta[first:last], dt[first:last]
        before = eng.metrics()
        t1 = time.time()
        exps = eng.evolve_schedule(omega, delta, dt, self._pack_eops(e_ops) if e_ops is not None else None)
        t2 = time.time()
        after = eng.metrics()
        self._statevector = None
        self.results = None  # drop our own references to the previous state buffer
        result = {}
        if self.return_state == "shard":
            result["state"] = self.local_statevector.reshape(-1, 1)
        elif self.return_state:
            result["state"] = self.statevector.reshape(-1, 1)
        if e_ops is not None:
            result["expectations"] = exps if exps is not None else np.zeros((omega.size, 0))
        self.results = result
        self.sij = None
        self.spins = self.get_spins() if self.compute_spins else None
        self.metrics = {
            "steps": int(omega.size),
            "hpsi_calls": int(after["hpsi_calls"] - before["hpsi_calls"]),
            "kernel_launches": int(after["kernel_launches"] - before["kernel_launches"]),
            "substeps": int(after["substeps"] - before["substeps"]),
            "hpsi_passes": int(after["hpsi_passes"]),
            "evolve_s": t2 - t1,
            "total_s": time.time() - t0,
        }
        if self.wtimes:
            print(f"time in compute and simulation = {self.metrics['evolve_s']} s.")
            print(f"total time = {self.metrics['total_s']} s.")
        return result

    @property
    def statevector(self):
        """Flat (2^N,) complex128 final state (gathered over ranks when sharded), fetched lazily."""
        if self._statevector is None:
            if self._engine is None:
                return None
            local = self._engine.get_state(out=self._state_buffer(self._engine.local_dim))
            if self._engine.nranks > 1:
                from . import distributed as qdist

                local = qdist.gather_shards(local)
            self._statevector = local
        return self._statevector

    @property
    def local_statevector(self):
        """This rank's shard of the final state, flat (2^(N-p),) complex128 (the whole state on one GPU)."""
        if self._engine is None:
            return None
        return self._engine.get_state(out=self._state_buffer(self._engine.local_dim))

    def _state_buffer(self, n: int):
        """Page-locked host buffer for the state read-back (PCIe speed instead of a staged copy).

        Locking memory is slow (~0.3 ms/MiB), so the buffer is kept and reused by the next
        ``calculate()`` — but only if nobody else still references it (a result the caller kept
        is never overwritten; a new buffer is locked instead)."""
        import sys

        if n < (1 << 18):
            return None
        buf = getattr(self, "_pinned", None)
        # 3 = the attribute, the local name and getrefcount's own argument: nobody else holds it
        if buf is None or buf.size != n or sys.getrefcount(buf) > 3:
            buf = lib.pinned_empty(n)
            self._pinned = buf
        return buf

    @statevector.setter
    def statevector(self, value):
        self._statevector = value

    # ---- observables (override qse/calc/calculator.py:125-187) ------------------------------
    def get_spins(self):
        """(N, 3) <X_i>, <Y_i>, <Z_i> — qse.magnetic.get_spins on the device."""
        if self.results is None:
            self.calculate()
        return self._engine.spins()

    def get_sij(self):
        """(N, N) spin correlation — qse.magnetic.get_sisj on the device; caches ``self.sij``."""
        if self.results is None:
            self.calculate()
        self.sij = self._engine.sisj()
        return self.sij

    def get_number_operator(self):
        """(N,) <n_i> — qse.magnetic.get_number_operator on the device."""
        if self.results is None:
            self.calculate()
        return self._engine.number()

    def structure_factor_from_sij(self, L1: int, L2: int, L3: int):
        """Structure factor S[q] on the L1 x L2 x L3 reciprocal grid (host numpy, N <= 40).

        The reference function raises TypeError (qse/magnetic.py:284: tuple @ Cell); this is the
        DFT its docstring (qse/magnetic.py:251-256) defines — PARITY UNPINNED.
        """
        if self.sij is None:
            self.get_sij()
        from .magnetic import structure_factor_from_sij

        return structure_factor_from_sij(L1, L2, L3, self.qbits, self.sij)

    def sample_final_state(self, shots: int = 1000, seed: int = 0):
        """Measurement read-out: ``Counter({bitstring: count})`` of ``shots`` projective measurements
        of the final state (what Pulser's ``results.sample_final_state`` returns; qubit 0 first)."""
        from collections import Counter

        if self.results is None:
            self.calculate()
        n = self._nqbits()
        idx, counts = np.unique(self._engine.sample(shots, seed), return_counts=True)
        return Counter({np.binary_repr(int(i), n): int(k) for i, k in zip(idx, counts)})

    def bitstring_probabilities(self, k: int = 16):
        """Top-k basis states as ``[(bitstring, probability), ...]`` (adiabatic.ipynb cell 8)."""
        if self.results is None:
            self.calculate()
        n = self._nqbits()
        idx, p = self._engine.topk(k)
        return [(np.binary_repr(int(i), n), float(q)) for i, q in zip(idx, p) if int(i) < (1 << n)]
