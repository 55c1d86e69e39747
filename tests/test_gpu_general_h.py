"""Parity of the generalised Hamiltonians (-m gpu): per-site drive / detuning
(README.md:93-95) and operator-sum Hamiltonians from qse.Operators (docs/use-cases/time_evo.ipynb TFIM, XY
couplings, mixed strings), through the C ABI against the CPU oracle.  Same gates as the Rydberg path:
infidelity <= 1e-10, state <= 1e-9, H psi <= 1e-13 relative."""
import numpy as np
import pytest

import qse_b200 as qb
from oracle import rydberg_oracle as orc
from qse_b200 import lib

pytestmark = pytest.mark.gpu

OMEGA_MAX = 2 * np.pi


def rand_state(n, seed):
    rng = np.random.default_rng(seed)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return psi / np.linalg.norm(psi)


@pytest.mark.parametrize("n,path", [(6, 0), (12, 0), (14, 0), (14, 2), (17, 1), (22, 0), (23, 0)])
def test_per_site_hpsi(n, path):
    """H psi with per-site Omega_i, delta_i on every kernel: small / untiled / one launch per bit group / the
    warp-specialised fused launch (14..21 qubits) and fused + plain launches (22+)."""
    reg = qb.lattices.random_register(n, 6.0, seed=300 + n)
    rng = np.random.default_rng(n)
    w_amp, w_det = rng.uniform(0.2, 1.8, n), rng.uniform(-1.0, 2.0, n)
    e = orc.diagonal_energies(orc.interaction_matrix(reg.positions), n)
    psi = rand_state(n, 10 + n)
    omega, delta = 5.1, -3.3
    with lib.Engine(n) as eng:
        eng.set_option("hpsi_path", path)
        eng.set_interaction(qb.interaction_matrix(reg.positions))
        eng.set_site_weights(w_amp, w_det)
        got = eng.apply_h_to(psi, omega, delta)
        eng.set_state(psi)
        assert np.array_equal(eng.apply_h(omega, delta), got)
        en = eng.energy(omega, delta)
    want = orc.general_apply(psi, n, e_int=e, hw_sites=0.5 * omega * w_amp, d_sites=delta * w_det)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))
    assert abs(en - np.real(np.vdot(psi, want))) <= 1e-10 * max(1.0, abs(en))


@pytest.mark.parametrize("n", [9, 13])
def test_per_site_sweep(n):
    """Adiabatic ramp with per-site weights through the calculator (9 qubits: the single-launch schedule kernel;
    13: the tiled kernels) against the exact propagator of every step."""
    reg = qb.lattices.random_register(n, 0.9 * qb.blockade_radius(OMEGA_MAX), seed=n)
    rng = np.random.default_rng(100 + n)
    w_amp, w_det = rng.uniform(0.5, 1.5, n), rng.uniform(0.5, 1.5, n)
    t = 6
    amp = qb.Signal(np.linspace(0.5, OMEGA_MAX, t), 5 * t)
    det = qb.Signal(np.linspace(-OMEGA_MAX, OMEGA_MAX, t), 5 * t)
    calc = qb.B200(reg, amp, det, site_amplitude=w_amp, site_detuning=w_det)
    got = calc.calculate()["state"].reshape(-1)
    diag = calc.get_hamiltonian(1.0, 2.0).diagonal()
    calc.close()
    e = orc.diagonal_energies(orc.interaction_matrix(reg.positions), n)
    k = np.arange(1 << n)
    want_diag = e - 2.0 * sum(w_det[q] * ((k >> (n - 1 - q)) & 1) for q in range(n))
    assert np.max(np.abs(diag - want_diag)) <= 1e-12 * np.max(np.abs(want_diag))
    psi = np.zeros(1 << n, complex)
    psi[0] = 1.0
    for a, d in zip(amp.values, det.values):
        psi = orc.general_propagate(psi, n, 0.005, e_int=e, hw_sites=0.5 * a * w_amp, d_sites=d * w_det)
    assert orc.infidelity(got, psi) <= 1e-10
    assert np.max(np.abs(got - psi)) <= 1e-9


def tfim(n, j=1.0, h=0.5):
    qbits = qb.lattices.chain(1.0, n)
    ham = qbits.compute_interaction_hamiltonian(lambda d: -j * np.isclose(d, 1.0), "Z")
    for i in range(n):
        ham += qb.Operator("X", i, n, coef=-h)
    return ham


@pytest.mark.parametrize("n", [6, 13, 15])
def test_tfim_time_evolution(n):
    """docs/use-cases/time_evo.ipynb: exp(-i H t)|0...0> for the transverse-field Ising chain.  ZZ strings fold into
    the (on-the-fly) diagonal, single-X strings into the per-site drive: the whole model runs in the tiled kernels."""
    ham = tfim(n)
    got = qb.evolve_operators(ham, 1.0)
    psi0 = np.zeros(1 << n, complex)
    psi0[0] = 1.0
    want = orc.general_propagate(psi0, n, 1.0, terms=orc.operators_as_terms(ham))
    assert orc.infidelity(got, want) <= 1e-10
    assert np.max(np.abs(got - want)) <= 1e-9


@pytest.mark.parametrize("n", [5, 13, 16])
def test_generic_strings_hpsi(n):
    """XX + YY couplings (the XY channel of qse/calc/pulser.py:174-177), Y and mixed strings, a 3-local diagonal
    string (no 2-local model: the stored-E path), on top of the Rydberg interaction."""
    reg = qb.lattices.random_register(n, 6.0, seed=40 + n)
    ham = reg.compute_interaction_hamiltonian(lambda d: 3.0 / d**3, ["X", "X"])
    ham += reg.compute_interaction_hamiltonian(lambda d: 3.0 / d**3, ["Y", "Y"])
    ham += qb.Operator("Y", n - 1, n, coef=0.7)
    ham += qb.Operator(["X", "Z"], [0, n - 2], n, coef=-0.4)
    ham += qb.Operator(["Z", "Z", "N"], [0, 1, n - 1], n, coef=1.1)
    ham += qb.Operator(["Z", "N"], [1, 2], n, coef=-0.6)
    ham += qb.Operator("Z", 0, n, coef=0.3)
    e = orc.diagonal_energies(orc.interaction_matrix(reg.positions), n)
    psi = rand_state(n, 77 + n)
    omega, delta = 2.2, 0.9
    with lib.Engine(n) as eng:
        eng.set_interaction(qb.interaction_matrix(reg.positions))
        eng.set_operator_terms(qb.operator_terms(ham, n))
        got = eng.apply_h_to(psi, omega, delta)
    want = orc.general_apply(psi, n, e_int=e, hw_sites=np.full(n, 0.5 * omega), d_sites=np.full(n, delta),
                             terms=orc.operators_as_terms(ham))
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


def test_time_scaled_operator_channels():
    """H(t) = amp(t) H_amp + det(t) H_det + H_static with operator-valued H_amp / H_det: the general form of
    qse/calc/qutip.py:55-58 (there H_amp = sum X / 2, H_det = -sum N)."""
    n = 8
    reg = qb.lattices.chain(5.5, n)
    h_amp = qb.Operators([qb.Operator(["X", "X"], [i, i + 1], n, coef=0.1) for i in range(n - 1)])
    h_det = qb.Operators([qb.Operator("Z", i, n, coef=0.05 * (i + 1)) for i in range(n)])
    h_static = qb.Operators([qb.Operator(["Y", "Y"], [0, n - 1], n, coef=0.2)])
    t = 5
    amp = qb.Signal(np.linspace(1.0, 4.0, t), 10 * t)
    det = qb.Signal(np.linspace(-2.0, 3.0, t), 10 * t)
    calc = qb.B200(reg, amp, det, hamiltonian=h_static, hamiltonian_amplitude=h_amp, hamiltonian_detuning=h_det)
    got = calc.calculate()["state"].reshape(-1)
    calc.close()
    e = orc.diagonal_energies(orc.interaction_matrix(reg.positions), n)
    psi = np.zeros(1 << n, complex)
    psi[0] = 1.0
    for a, d in zip(amp.values, det.values):
        terms = orc.operators_as_terms(h_static) + orc.operators_as_terms(h_amp * float(a)) + orc.operators_as_terms(h_det * float(d))
        psi = orc.general_propagate(psi, n, 0.01, e_int=e, hw_sites=np.full(n, 0.5 * a), d_sites=np.full(n, d), terms=terms)
    assert orc.infidelity(got, psi) <= 1e-10
    assert np.max(np.abs(got - psi)) <= 1e-9


def test_hamiltonian_apply_leaves_the_evolved_state_alone():
    """Hamiltonian.apply(psi) must not overwrite the sweep result on the device."""
    n = 10
    reg = qb.lattices.random_register(n, 6.0, seed=1)
    calc = qb.B200(reg, qb.Signal(np.linspace(1, 5, 4), 40), qb.Signal(np.linspace(-3, 3, 4), 40))
    calc.calculate()
    before = calc.get_spins()
    other = rand_state(n, 3)
    y = calc.get_hamiltonian(2.0, 1.0).apply(other)
    e = orc.diagonal_energies(orc.interaction_matrix(reg.positions), n)
    assert np.max(np.abs(y - orc.apply_hamiltonian(other, e, n, 2.0, 1.0))) <= 1e-12
    assert np.array_equal(calc.get_spins(), before)
    calc.close()


def test_topk_on_sparse_states():
    """top-k of |0...0> (all ties at probability zero) and of a state with fewer than k non-zeros."""
    n = 18
    with lib.Engine(n) as eng:
        eng.set_interaction(np.zeros((n, n)))
        eng.reset_state()
        idx, p = eng.topk(16)
        assert list(idx) == list(range(16)) and p[0] == 1.0 and np.all(p[1:] == 0.0)
        psi = np.zeros(1 << n, complex)
        psi[[5, 70000, 200001]] = [0.6, 0.0 + 0.64j, 0.48]
        eng.set_state(psi)
        idx, p = eng.topk(8)
        assert list(idx[:3]) == [70000, 5, 200001] and list(idx[3:]) == [0, 1, 2, 3, 4]
        assert np.allclose(p[:3], [0.4096, 0.36, 0.2304], rtol=1e-15) and np.all(p[3:] == 0.0)


@pytest.mark.parametrize("n", [12, 13, 17, 21, 22, 24])
def test_tiled_spins_match_the_per_qubit_path_and_the_oracle(n):
    """<X_q>, <Y_q>, <Z_q> of all qubits from one tile pass per group of index bits (3 reads of the state at 25
    qubits) against the per-qubit Pauli-string path (1 + 2N reads) and, where it is affordable, the oracle
    (qse/magnetic.py:119-176)."""
    psi = rand_state(n, 500 + n)
    with lib.Engine(n) as eng:
        eng.set_interaction(np.zeros((n, n)))
        eng.set_state(psi)
        tiled = eng.spins()
        launches = eng.metrics()["kernel_launches"]
        eng.set_option("tiled_observables", 0)
        plain = eng.spins()
        launches_plain = eng.metrics()["kernel_launches"] - launches
        number = eng.number()
    assert np.max(np.abs(tiled - plain)) <= 1e-13
    assert np.max(np.abs(tiled[:, 2] - (1.0 - 2.0 * number))) <= 1e-13
    assert launches_plain >= 4 * n  # what the tile passes replace
    if n <= 17:
        assert np.max(np.abs(tiled - orc.get_spins(psi, n))) <= 1e-12


@pytest.mark.parametrize("n", [8, 13])
def test_per_site_schedule_rows(n):
    """qsb_evolve_schedule_sites: every step carries its own Omega_q, delta_q per site (n_steps x N rows) -- the fully
    general form of the README Hamiltonian's time dependence.  8 qubits: single-launch schedule kernel; 13: tiled."""
    reg = qb.lattices.random_register(n, 0.9 * qb.blockade_radius(OMEGA_MAX), seed=60 + n)
    rng = np.random.default_rng(n)
    steps = 5
    om = rng.uniform(0.5, OMEGA_MAX, size=(steps, n))
    de = rng.uniform(-OMEGA_MAX, OMEGA_MAX, size=(steps, n))
    dt = np.full(steps, 0.004)
    e = orc.diagonal_energies(orc.interaction_matrix(reg.positions), n)
    with lib.Engine(n) as eng:
        eng.set_interaction(qb.interaction_matrix(reg.positions))
        eng.reset_state()
        eng.evolve_schedule_sites(om, de, dt)
        got = eng.get_state()
    psi = np.zeros(1 << n, complex)
    psi[0] = 1.0
    for s in range(steps):
        psi = orc.general_propagate(psi, n, dt[s], e_int=e, hw_sites=0.5 * om[s], d_sites=de[s])
    assert orc.infidelity(got, psi) <= 1e-10
    assert np.max(np.abs(got - psi)) <= 1e-9
