"""The Python plugin surface of B200 exercised on CPU against a FAKE engine (oracle-backed, tests
only): constructor / attachment rules, calculate() return value and attributes, e_ops packing,
lazy getters, checkpoint arguments, bit-string helpers.  The real engine is covered by -m gpu."""
import numpy as np
import pytest

import qse_b200 as qb
from oracle import rydberg_oracle as orc
from qse_b200 import calculator as calc_mod
from qse_b200 import lib


class FakeEngine:
    """Same interface as lib.Engine, arithmetic by the CPU oracle (single rank)."""

    def __init__(self, n_qubits, device=0, rank=0, nranks=1, nccl_uid=None):
        self.n, self.rank, self.nranks = n_qubits, rank, nranks
        self.local_dim = 1 << n_qubits
        self.psi = np.zeros(self.local_dim, dtype=complex)
        self.psi[0] = 1.0
        self.e = None
        self.options = {}
        self.closed = False
        self.hpsi = 0

    def set_option(self, key, value):
        self.options[key] = value

    def set_interaction(self, v):
        self.e = orc.diagonal_energies(np.asarray(v), self.n)

    def diagonal(self):
        return self.e.copy()

    def reset_state(self):
        self.psi[:] = 0
        self.psi[0] = 1.0

    def set_state(self, psi):
        self.psi = np.array(psi, dtype=complex).reshape(-1)

    def get_state(self, out=None, pinned=False):
        if out is None:
            return self.psi.copy()
        out[:] = self.psi
        return out

    def evolve_schedule(self, omega, delta, dt, eops=None):
        rows = []
        for a, d, t in zip(omega, delta, dt):
            self.psi = orc.propagate_segment(self.psi, self.e, self.n, a, d, t, "eigh")
            self.hpsi += 1
            if eops is not None and eops.n_ops:
                rows.append([self._eop(eops, o) for o in range(eops.n_ops)])
        return np.array(rows) if rows else None

    def _eop(self, eops, o):
        op = eops.struct.ops[o]
        if op.kind == 0:
            return orc.energy(self.psi, self.e, self.n, op.omega, op.delta)
        total = 0.0
        for t in range(op.term_begin, op.term_end):
            masks = [int(m[t]) for m in eops._masks]
            terms = []
            for name, m in zip("XYZN", masks):
                terms += [(name, self.n - 1 - b) for b in range(self.n) if (m >> b) & 1]
            total += orc.pauli_expectation(self.psi, self.n, [(eops._coef[t], terms)]).real
        return total

    def metrics(self):
        return {"hpsi_calls": self.hpsi, "kernel_launches": self.hpsi, "substeps": self.hpsi, "hpsi_passes": 1}

    def spins(self):
        return orc.get_spins(self.psi, self.n)

    def sisj(self):
        return orc.get_sisj(self.psi, self.n)

    def number(self):
        return orc.get_number_operator(self.psi, self.n)

    def topk(self, k):
        p = np.abs(self.psi) ** 2
        order = np.lexsort((np.arange(p.size), -p))[:k]
        return order.astype(np.uint64), p[order]

    def sample(self, shots, seed=0):
        rng = np.random.default_rng(seed)
        return rng.choice(self.local_dim, size=shots, p=np.abs(self.psi) ** 2).astype(np.uint64)

    def close(self):
        self.closed = True


@pytest.fixture()
def fake(monkeypatch):
    monkeypatch.setattr(lib, "is_available", lambda: (True, ""))
    monkeypatch.setattr(lib, "Engine", FakeEngine)
    monkeypatch.setattr(lib, "pinned_empty", lambda n, dtype=np.complex128: np.empty(n, dtype=dtype))
    return FakeEngine


class _Op:
    def __init__(self, operator, indices, nqbits, coef=1.0):
        self.operator = [operator] * len(indices) if isinstance(operator, str) else operator
        self.indices, self.nqbits, self.coef = list(indices), nqbits, coef


class _Ops:
    def __init__(self, operator_list):
        self.operator_list = operator_list


def _setup(t=6):
    qbits = qb.lattices.square(0.9 * qb.blockade_radius(2 * np.pi), 2, 2)
    amp = qb.Signal(np.linspace(0, 2 * np.pi, t), 10 * t)
    det = qb.Signal(np.linspace(-5, 5, t), 10 * t)
    return qbits, amp, det


def test_calculate_returns_qutip_dict_and_sets_pulser_attributes(fake):
    qbits, amp, det = _setup()
    calc = qb.B200(qbits, amp, det)
    assert calc.label == "b200" and calc.results is None and calc.c6 == 5420158.53
    assert isinstance(calc.amplitude, qb.Signals) and len(calc.amplitude) == 1  # bare Signal wrapped
    out = calc.calculate()
    ref = orc.calculate(qbits.positions, [(amp.values, amp.duration)], [(det.values, det.duration)], method="eigh")
    assert set(out) == {"state"} and out["state"].shape == (16, 1) and out["state"].dtype == np.complex128
    assert np.max(np.abs(out["state"] - ref["state"])) < 1e-12
    assert calc.results is out and calc.statevector.shape == (16,)
    assert calc.spins.shape == (4, 3) and np.allclose(calc.spins, orc.get_spins(ref["state"], 4))
    assert calc.metrics["steps"] == 6
    assert calc._engine.options == {"tol": 1e-14, "fixed_terms": 0}


def test_e_ops_vocabulary(fake):
    qbits, amp, det = _setup()
    calc = qb.B200(qbits, amp, det, method="rk4")
    h = calc.get_hamiltonian(1.0, 0.5)
    zz = _Ops([_Op("Z", [0], 4, 0.5), _Op(["X", "Y"], [1, 3], 4, 2.0)])
    out = calc.calculate(e_ops=[h, ("H", 2.0, -1.0), zz, _Op("N", [2], 4)])
    assert out["expectations"].shape == (6, 4)
    assert calc._engine.options["fixed_terms"] == 4
    psi = out["state"].reshape(-1)
    e = orc.diagonal_energies(orc.interaction_matrix(qbits.positions), 4)
    assert abs(out["expectations"][-1, 0] - orc.energy(psi, e, 4, 1.0, 0.5)) < 1e-12
    want = orc.pauli_expectation(psi, 4, [(0.5, [("Z", 0)]), (2.0, [("X", 1), ("Y", 3)])]).real
    assert abs(out["expectations"][-1, 2] - want) < 1e-12
    assert abs(out["expectations"][-1, 3] - orc.get_number_operator(psi, 4)[2]) < 1e-12
    with pytest.raises(TypeError, match="e_ops entries"):
        calc.calculate(e_ops=[np.eye(16)])


def test_lazy_getters_and_reattachment(fake):
    qbits, amp, det = _setup()
    calc = qb.B200(None, amp, det, compute_spins=False)
    with pytest.raises(Exception, match="No qbits"):
        calc.get_spins()
    calc.qbits = qbits
    assert calc.get_spins().shape == (4, 3)  # triggered calculate()
    assert calc.results is not None and calc.spins is None
    sij = calc.get_sij()
    assert calc.sij is sij and np.allclose(np.diag(sij), 1.0)
    sf = calc.structure_factor_from_sij(2, 2, 1)
    assert sf.shape == (2, 2, 1)
    first_engine = calc._engine
    calc.qbits = qb.lattices.chain(7.0, 3)  # new register: the device context is rebuilt, H is not stale
    assert calc.calculate()["state"].shape == (8, 1)
    assert first_engine.closed and calc._engine is not first_engine
    host = qb.Qbits(np.array([[0.0, 0.0], [9.0, 0.0]]), calculator=calc)  # qse/qbits.py:169-173 hook
    assert calc.qbits is host


def test_results_are_not_overwritten_while_referenced(fake, monkeypatch):
    big = 1 << 18
    monkeypatch.setattr(calc_mod.B200, "_nqbits", lambda self: 4)
    qbits, amp, det = _setup()
    calc = qb.B200(qbits, amp, det)
    first = id(calc._state_buffer(big))
    assert id(calc._state_buffer(big)) == first   # nobody else holds it: reused
    keep = calc._state_buffer(big)                 # a second reference, like a result the caller kept
    assert id(keep) == first
    assert calc._state_buffer(big) is not keep     # ... is never overwritten: a new buffer is locked
    assert calc._state_buffer(100) is None         # small states use ordinary memory


def test_checkpoint_arguments(fake, tmp_path):
    qbits, amp, det = _setup(8)
    calc = qb.B200(qbits, amp, det)
    ref = calc.calculate()["state"].copy()
    head = calc.calculate(stop_after=3)
    calc.save_checkpoint(str(tmp_path / "c"), next_step=3)
    tail = calc.calculate(resume_from=str(tmp_path / "c"))
    assert np.max(np.abs(tail["state"] - ref)) < 1e-12 and head["state"].shape == (16, 1)
    with pytest.raises(Exception, match="cannot run schedule values"):
        calc.calculate(stop_after=99)
    other = qb.B200(qb.lattices.chain(5.0, 4), amp, det)
    with pytest.raises(Exception, match="different qubit positions"):
        other.calculate(resume_from=str(tmp_path / "c"))


def test_bitstring_helpers(fake):
    qbits, amp, det = _setup()
    calc = qb.B200(qbits, amp, det)
    top = calc.bitstring_probabilities(3)
    assert len(top) == 3 and all(len(b) == 4 for b, _ in top) and top[0][1] >= top[1][1] >= top[2][1]
    counts = calc.sample_final_state(500, seed=3)
    assert sum(counts.values()) == 500 and all(len(b) == 4 for b in counts)
    calc.close()
    assert calc._engine is None


# ---- generalised Hamiltonians through the plugin surface, fake engine ----------------
class GeneralFakeEngine(FakeEngine):
    """FakeEngine + the per-site / operator-term entry points, arithmetic by the oracle's general propagator."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.w_amp = np.ones(self.n)
        self.w_det = np.ones(self.n)
        self.terms = []

    def set_site_weights(self, w_omega=None, w_delta=None):
        self.w_amp = np.ones(self.n) if w_omega is None else np.asarray(w_omega, float)
        self.w_det = np.ones(self.n) if w_delta is None else np.asarray(w_delta, float)

    def set_operator_terms(self, terms):
        self.terms = list(terms)

    def _h(self, a, d):
        chan = {0: 1.0, 1: a, 2: d}
        terms = []
        for coef, x, y, z, nm, ch in self.terms:
            ops = []
            for name, m in zip("XYZN", (x, y, z, nm)):
                ops += [(name, self.n - 1 - b) for b in range(self.n) if (m >> b) & 1]
            terms.append((coef * chan[ch], ops))
        return dict(e_int=self.e, hw_sites=0.5 * a * self.w_amp, d_sites=d * self.w_det, terms=terms)

    def evolve_schedule(self, omega, delta, dt, eops=None):
        for a, d, t in zip(omega, delta, dt):
            self.psi = orc.general_propagate(self.psi, self.n, t, **self._h(a, d))
            self.hpsi += 1
        return None

    def apply_h_to(self, psi, omega, delta):
        return orc.general_apply(np.asarray(psi, complex).reshape(-1), self.n, **self._h(omega, delta))


@pytest.fixture()
def general_fake(monkeypatch):
    monkeypatch.setattr(lib, "is_available", lambda: (True, ""))
    monkeypatch.setattr(lib, "Engine", GeneralFakeEngine)
    monkeypatch.setattr(lib, "pinned_empty", lambda n, dtype=np.complex128: np.empty(n, dtype=dtype))
    return GeneralFakeEngine


def test_site_weights_and_operator_hamiltonians_reach_the_engine(general_fake):
    qbits, amp, det = _setup(4)
    n = qbits.nqbits
    w, v = np.array([1.0, 0.5, 2.0, 1.5]), np.array([0.0, 1.0, 1.0, -1.0])
    h_static = qb.Operators([qb.Operator(["Z", "Z"], [0, 1], n, coef=0.3), qb.Operator("Y", 2, n, coef=0.1)])
    h_amp = qb.Operators([qb.Operator(["X", "X"], [1, 3], n, coef=0.2)])
    calc = qb.B200(qbits, amp, det, site_amplitude=w, site_detuning=v, hamiltonian=h_static, hamiltonian_amplitude=h_amp)
    out = calc.calculate()
    eng = calc._engine
    assert np.array_equal(eng.w_amp, w) and np.array_equal(eng.w_det, v)
    assert [t[-1] for t in eng.terms] == [0, 0, 1]  # channels: static, static, x amplitude
    assert eng.terms[0][:5] == (0.3, 0, 0, 0b1100, 0) and eng.terms[1][:5] == (0.1, 0, 0b0010, 0, 0)
    assert eng.terms[2][:5] == (0.2, 0b0101, 0, 0, 0)
    # independent evaluation of the same schedule
    e = orc.diagonal_energies(orc.interaction_matrix(qbits.positions), n)
    psi = np.zeros(1 << n, complex)
    psi[0] = 1.0
    for a, d in zip(amp.values, det.values):
        terms = orc.operators_as_terms(h_static) + orc.operators_as_terms(h_amp * float(a))
        psi = orc.general_propagate(psi, n, 0.01, e_int=e, hw_sites=0.5 * a * w, d_sites=d * v, terms=terms)
    assert np.max(np.abs(out["state"].reshape(-1) - psi)) < 1e-12
    # Hamiltonian.apply goes through apply_h_to and leaves the evolved state alone
    before = calc.statevector.copy()
    y = calc.get_hamiltonian(2.0, 1.0).apply(np.eye(1 << n)[3])
    assert y.shape == (1 << n,) and np.array_equal(eng.psi, before)
    # a changed Hamiltonian rebuilds the context (the reference goes stale)
    calc.hamiltonian = None
    calc.calculate()
    assert calc._engine is not eng and eng.closed and len(calc._engine.terms) == 1
    calc.close()
    with pytest.raises(Exception, match="one weight per qubit"):
        qb.B200(qbits, amp, det, site_amplitude=[1.0, 2.0]).calculate()
    with pytest.raises(Exception, match="number of qubits"):
        qb.B200(qbits, amp, det, hamiltonian=qb.Operators([qb.Operator("X", 0, n + 1)])).calculate()


def test_interaction_false_and_shard_return(general_fake):
    qbits, amp, det = _setup(3)
    calc = qb.B200(qbits, amp, det, interaction=False, return_state="shard")
    out = calc.calculate()
    assert np.all(calc._engine.e == 0.0)
    assert out["state"].shape == (16, 1)  # one rank: the shard is the whole state
    assert np.array_equal(out["state"].reshape(-1), calc.local_statevector)
    calc.close()
