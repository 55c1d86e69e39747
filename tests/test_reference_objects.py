"""Container-only: the calculator's host side fed with the REFERENCE's own objects (qse.Qbits from
qse.lattices, qse.Signal / qse.Signals, qse.Operator / qse.Operators), imported unmodified from
/root/reference through oracle/refshim.py.  Skipped where the reference tree is absent (GPU box).
No device work: this pins the duck-typed boundary (what B200 reads from those objects)."""
import numpy as np
import pytest

from oracle import refshim
from oracle import rydberg_oracle as orc

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="reference tree only exists in the build container")


@pytest.fixture(scope="module")
def qse():
    return refshim.import_reference()


def test_interaction_matrix_from_reference_qbits(qse):
    import qse_b200 as qb

    for q in (qse.lattices.square(0.9 * qse.calc.blockade_radius(2 * np.pi), 3, 3), qse.lattices.kagome(6.0, 2, 2),
              qse.lattices.ring(5.0, 8), qse.lattices.chain(60.0, 6)):
        v = qb.interaction_matrix(q.positions)
        ops = q.compute_interaction_hamiltonian(lambda d: qb.C6 / d**6, "N")  # qse/calc/qutip.py:61-63
        assert int(np.count_nonzero(v)) == len(ops.operator_list)
        for op in ops.operator_list:
            i, j = op.indices
            assert v[i, j] == op.coef  # bit-identical coefficient, same pairs kept by the 1e-8 cut-off


def test_schedule_from_reference_signals(qse):
    import qse_b200 as qb
    from qse_b200 import calculator as calc_mod

    amp = qse.Signals([qse.Signal(np.linspace(0, 1, 5), 20), qse.Signal([3.0], 400)])
    det = qse.Signals([qse.Signal(np.linspace(-1, 1, 5), 10), qse.Signal([0.5], 410)])
    calc_mod._check_pulses(amp, det)
    om, de, dt = qb.flatten_schedule(amp, det)
    # the loop of Qutip.calculate, restated literally (qse/calc/qutip.py:40-42)
    want = [(a, d, a_s.time_per_value() / 1000) for a_s, d_s in zip(amp, det) for a, d in zip(a_s.values, d_s.values)]
    assert [(a, d, t) for a, d, t in zip(om, de, dt)] == want
    wrapped = calc_mod._as_signals(qse.Signal([1.0, 2.0], 4))  # bare Signal -> Signals([signal]) (:21-25)
    assert len(wrapped) == 1 and wrapped.duration == 4
    assert calc_mod._as_signals(amp) is amp
    with pytest.raises(Exception, match="same duration"):
        calc_mod._check_pulses(qse.Signals([qse.Signal([1.0], 10)]), qse.Signals([qse.Signal([1.0], 20)]))


def test_pauli_masks_from_reference_operators(qse):
    from qse_b200 import calculator as calc_mod

    n = 5
    rng = np.random.default_rng(2)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi /= np.linalg.norm(psi)
    ops = qse.Operators([qse.Operator("Z", 1, n, 0.7), qse.Operator(["X", "Y"], [0, 3], n, -1.2),
                         qse.Operator("N", [2, 4], n, 2.0)])
    total = 0.0
    for t in ops.operator_list:
        x, y, z, nm = calc_mod._pauli_masks(t.operator, t.indices, n)
        # evaluate the masks with the kernel's formula: sum_k conj(psi[k^f]) phase(k) psi[k]
        k = np.arange(1 << n)
        f = x | y
        sel = (k & nm) == nm
        sign = 1 - 2 * (np.array([bin(int(v)).count("1") for v in (k & (y | z))]) & 1)
        val = (1j) ** bin(y).count("1") * np.sum(np.conj(psi[k ^ f]) * sign * psi * sel)
        total += t.coef * val
    want = orc.pauli_expectation(psi, n, [(t.coef, list(zip(t.operator, t.indices))) for t in ops.operator_list])
    assert abs(total - want) < 1e-13


def test_b200_is_a_reference_calculator_subclass(qse):
    """With the reference importable, B200 derives from qse.calc.calculator.Calculator itself."""
    import importlib
    import sys

    for name in [m for m in sys.modules if m.startswith("qse_b200")]:
        del sys.modules[name]
    qb = importlib.import_module("qse_b200")
    assert issubclass(qb.B200, qse.calc.Calculator)
    for name in ("get_spins", "get_sij", "structure_factor_from_sij", "calculate"):
        assert name in qb.B200.__dict__  # overridden, not inherited (base versions are O(4^N) Python)
