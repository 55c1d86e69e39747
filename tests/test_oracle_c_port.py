"""The C restatement of the oracle (used at sizes numpy/scipy cannot reach, and as the CPU
baseline) against the numpy/scipy oracle that is pinned to the reference's golden vectors."""
import numpy as np
import pytest

from oracle import c_port
from oracle import rydberg_oracle as orc


def _case(n, seed):
    from qse_b200.host import lattices

    rng = np.random.default_rng(seed)
    pos = lattices.random_register(n, 6.0, seed=seed).positions  # min distance 3.6 um: |H| stays sane
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return pos, psi / np.linalg.norm(psi)


@pytest.mark.parametrize("n", [1, 3, 8, 11, 12, 14, 17])
def test_diagonal_and_apply_h(n):
    pos, psi = _case(n, n)
    v = orc.interaction_matrix(pos)
    e = orc.diagonal_energies(v, n)
    assert np.array_equal(c_port.build_diagonal(v, n), e)  # bit-exact
    want = orc.apply_hamiltonian(psi, e, n, 3.3, -1.7)
    got = c_port.apply_h(psi, e, n, 3.3, -1.7)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


@pytest.mark.parametrize("n,method", [(4, "eigh"), (9, "eigh"), (13, "expm_multiply")])
def test_sweep_matches_exact_propagator(n, method):
    pos, _ = _case(n, 40 + n)
    t = 10
    amp = [(np.linspace(0, 6.0, t), 5 * t)]
    det = [(np.linspace(-12, 12, t), 5 * t)]
    ref = orc.calculate(pos, amp, det, method=method)["state"].reshape(-1)
    om, de, dt = orc.flatten_schedule(amp, det)
    psi0 = np.zeros(1 << n, dtype=complex)
    psi0[0] = 1.0
    e = orc.diagonal_energies(orc.interaction_matrix(pos), n)
    got, count = c_port.evolve_schedule(psi0, e, n, om, de, dt)
    assert count >= t
    assert orc.infidelity(got, ref) <= 1e-12
    assert np.max(np.abs(got - ref)) <= 1e-11
    assert np.max(np.abs(c_port.spins(got, n) - orc.get_spins(ref, n))) <= 1e-10
    assert np.max(np.abs(c_port.number(got, n) - orc.get_number_operator(ref, n))) <= 1e-10


def test_threads_reported():
    assert c_port.threads() >= 1
