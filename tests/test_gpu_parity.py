"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle
on the same seeded inputs, against the committed golden vectors of the reference, and — at
BASELINE.json's full sizes — through size-independent properties.

Tolerances (BASELINE.json north_star): final-state infidelity <= 1e-10, observables <= 1e-9
absolute (complex128), diagonal / basis indexing bit-exact."""
import os

import numpy as np
import pytest

import qse_b200 as qb
from oracle import rydberg_oracle as orc
from qse_b200 import lib

pytestmark = pytest.mark.gpu

OMEGA_MAX = 2 * np.pi
INFID_TOL = 1e-10
OBS_TOL = 1e-9


def random_state(n, seed):
    rng = np.random.default_rng(seed)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return psi / np.linalg.norm(psi)


def random_positions(n, seed, spacing=6.0):
    return qb.lattices.random_register(n, spacing, seed=seed).positions


# ---- K1: diagonal -------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 5, 9, 12, 15])
def test_diagonal_bit_exact(n):
    pos = random_positions(n, 100 + n)
    v = qb.interaction_matrix(pos)
    with lib.Engine(n) as eng:
        eng.set_interaction(v)
        got = eng.diagonal()
        m = eng.metrics()
    want = orc.diagonal_energies(orc.interaction_matrix(pos), n)
    assert np.array_equal(got, want)  # bit-exact, including the index -> qubit mapping
    assert m["diag_min"] == want.min() and m["diag_max"] == want.max()


def test_diagonal_with_cutoff_pairs(golden):
    g = golden("interaction.npz")
    pos = g["3_positions"]  # long chain: far pairs dropped by the 1e-8 cut-off
    n = pos.shape[0]
    with lib.Engine(n) as eng:
        eng.set_interaction(qb.interaction_matrix(pos))
        assert np.array_equal(eng.diagonal(), orc.diagonal_energies(orc.interaction_matrix(pos), n))


# ---- K2/K3: H psi --------------------------------------------------------------------------
@pytest.mark.parametrize("n,path", [(1, 0), (3, 0), (9, 0), (11, 0), (12, 0), (13, 0), (15, 0), (16, 1),
                                    (18, 0), (21, 0), (22, 0)])
def test_apply_h_matches_oracle(n, path):
    pos = random_positions(n, 7 * n)
    v = orc.interaction_matrix(pos)
    e = orc.diagonal_energies(v, n)
    psi = random_state(n, n)
    omega, delta = 1.9 * np.pi, -3.7
    with lib.Engine(n) as eng:
        eng.set_option("hpsi_path", path)
        eng.set_interaction(qb.interaction_matrix(pos))
        eng.set_state(psi)
        got = eng.apply_h(omega, delta)
        assert np.array_equal(eng.get_state(), psi)  # input untouched
    want = orc.apply_hamiltonian(psi, e, n, omega, delta)
    scale = np.max(np.abs(want))
    assert np.max(np.abs(got - want)) <= 1e-13 * scale


@pytest.mark.parametrize("n", [12, 14, 17, 22, 23])
def test_independent_kernels_agree(n):
    """Three independent kernels on one input: warp-specialised super-block path (0), one
    launch per tile-bit group (2), one thread per amplitude (1)."""
    pos = random_positions(n, 3 * n + 1)
    psi = random_state(n, 50 + n)
    out = {}
    for path in (0, 2, 1):
        with lib.Engine(n) as eng:
            eng.set_option("hpsi_path", path)
            eng.set_interaction(qb.interaction_matrix(pos))
            eng.set_state(psi)
            out[path] = eng.apply_h(4.0, 2.5)
            passes = eng.metrics()["hpsi_passes"]
            if path == 2:
                assert passes == 1 + (n - 12 + 8) // 9
            elif path == 1:
                assert passes == 1
            else:
                assert passes == (1 if n <= 21 else 2)
    assert np.max(np.abs(out[0] - out[1])) <= 1e-13 * np.max(np.abs(out[1]))
    assert np.max(np.abs(out[2] - out[1])) <= 1e-13 * np.max(np.abs(out[1]))


@pytest.mark.parametrize("n,sb", [(13, 13), (15, 13), (17, 13), (17, 15), (19, 16), (22, 14), (22, 19), (23, 21)])
def test_super_block_schedules(n, sb):
    """Many small super-blocks stress the in-kernel A -> B dependency counters; every schedule
    must give the same H psi (checked against the untiled kernel), twice in a row (epochs)."""
    pos = random_positions(n, 5 * n + sb)
    psi = random_state(n, 70 + n)
    with lib.Engine(n) as eng:
        eng.set_interaction(qb.interaction_matrix(pos))
        eng.set_state(psi)
        eng.set_option("hpsi_path", 1)
        want = eng.apply_h(4.0, 2.5)
        eng.set_option("hpsi_path", 0)
        eng.set_option("sb_bits", sb)
        for _ in range(3):
            got = eng.apply_h(4.0, 2.5)
            assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))
        eng.set_option("sb_lag", -1)  # no lead: B(s) handed out right after A(s) (the 30-qubit schedule)
        got = eng.apply_h(4.0, 2.5)
        assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))
        eng.set_option("sb_lag", 0)
        # changing the block size mid-life restarts the completion counters
        eng.set_option("sb_bits", 13 if sb != 13 else 14)
        got = eng.apply_h(4.0, 2.5)
        assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))
        assert eng.time_apply_h(4.0, 2.5, 4).shape == (4,)


# ---- time evolution --------------------------------------------------------------------------
def sweep(n_values, duration_per_value=1):
    t = n_values
    amp = qb.Signal(np.linspace(0, OMEGA_MAX, t), t * duration_per_value)
    det = qb.Signal(np.linspace(-2 * OMEGA_MAX, 2 * OMEGA_MAX, t), t * duration_per_value)
    return amp, det


def oracle_run(qbits, amp, det, **kw):
    return orc.calculate(qbits.positions, [(amp.values, amp.duration)], [(det.values, det.duration)], **kw)


def test_config1_square_3x3_sweep():
    """BASELINE config 1: square 3x3 at 0.9 r_b, adiabatic Omega/delta ramp, spins + <s_i s_j>."""
    qbits = qb.lattices.square(0.9 * qb.blockade_radius(OMEGA_MAX), 3, 3)
    amp, det = sweep(120, 10)  # 120 values of 10 ns
    calc = qb.B200(qbits, amp, det)
    out = calc.calculate()
    ref = oracle_run(qbits, amp, det, method="eigh")["state"]
    assert out["state"].shape == (512, 1) and out["state"].dtype == np.complex128
    assert orc.infidelity(out["state"], ref) <= INFID_TOL
    assert np.max(np.abs(out["state"] - ref)) <= OBS_TOL  # also phase-exact
    assert np.max(np.abs(calc.get_spins() - orc.get_spins(ref, 9))) <= OBS_TOL
    assert np.max(np.abs(calc.get_sij() - orc.get_sisj(ref, 9))) <= OBS_TOL
    assert np.max(np.abs(calc.get_number_operator() - orc.get_number_operator(ref, 9))) <= OBS_TOL
    assert calc.spins.shape == (9, 3) and calc.sij.shape == (9, 9)
    assert abs(np.vdot(calc.statevector, calc.statevector).real - 1) < 1e-12
    sf = calc.structure_factor_from_sij(3, 3, 1)
    assert np.max(np.abs(sf - orc.structure_factor_from_sij(3, 3, 1, orc.get_sisj(ref, 9)))) <= OBS_TOL
    calc.close()


def test_single_qubit_closed_form(golden):
    """tests/qse/calc_test.py:32-53 through the GPU calculator (N = 1, no interaction)."""
    g = golden("exact_n1.npz")
    dur = int(g["duration"])
    for omega, delta, ref_state, closed in zip(g["omega"], g["delta"], g["state"], g["closed_form"]):
        calc = qb.B200(qb.Qbits(np.zeros((1, 2))), qb.Signal([omega], dur), qb.Signal([delta], dur))
        psi = calc.calculate()["state"].reshape(-1)
        assert orc.infidelity(psi, ref_state) < INFID_TOL
        assert orc.infidelity(psi, closed) < INFID_TOL
        assert np.max(np.abs(psi - np.exp(1j * delta * dur / 2000) * closed)) < OBS_TOL
        calc.close()


def test_stored_reference_states(golden):
    """tests/local_tests/results_{myqlm,pulser}.npy: chain(4.0, 4), one 400 ns constant pulse —
    a single segment with |H| dt ~ 1e3, exercising the sub-stepping."""
    g = golden("local_tests.npz")
    dur = int(g["duration"])
    calc = qb.B200(qb.Qbits(g["positions"]), qb.Signal([float(g["omega"])], dur), qb.Signal([float(g["delta"])], dur))
    psi = calc.calculate()["state"].reshape(-1)
    assert calc.metrics["substeps"] > 1
    assert orc.infidelity(psi, g["myqlm_state"]) < 1e-7
    assert np.max(np.abs(calc.get_spins() - g["myqlm_spins"])) < 5e-5
    assert orc.infidelity(psi, g["pulser_state"]) < 1e-4
    ref = orc.calculate(g["positions"], [([float(g["omega"])], dur)], [([float(g["delta"])], dur)], method="eigh")
    assert orc.infidelity(psi, ref["state"]) <= INFID_TOL
    assert np.max(np.abs(psi - ref["state"].reshape(-1))) <= OBS_TOL
    calc.close()


def test_pulser_test_scenario_2x2():
    """tests/qse_pulser/pulser_test.py:83-106: 2x2 square at 0.9 r_b, 10.01 / 0.12, 400 ns."""
    qbits = qb.lattices.square(0.9 * qb.blockade_radius(10.01), 2, 2)
    amp, det = qb.Signal([10.01], 400), qb.Signal([0.12], 400)
    calc = qb.B200(qbits, amp, det)
    out = calc.calculate()
    ref = oracle_run(qbits, amp, det, method="eigh")["state"]
    assert orc.infidelity(out["state"], ref) <= INFID_TOL
    calc.close()


@pytest.mark.parametrize("name,args", [("triangular", (3, 4)), ("kagome", (2, 2)), ("ring", (10,)), ("chain", (11,))])
def test_lattice_families_against_exact_propagator(name, args):
    a = 0.9 * qb.blockade_radius(OMEGA_MAX)
    qbits = getattr(qb.lattices, name)(a, *args)
    n = qbits.nqbits
    amp, det = sweep(25, 4)
    calc = qb.B200(qbits, amp, det)
    out = calc.calculate()
    ref = oracle_run(qbits, amp, det, method="expm_multiply")["state"]
    assert orc.infidelity(out["state"], ref) <= INFID_TOL
    assert np.max(np.abs(calc.get_spins() - orc.get_spins(ref, n))) <= OBS_TOL
    calc.close()


@pytest.mark.parametrize("n", [13, 16])
def test_tiled_path_sweep(n):
    """Sizes that run the tiled multi-pass kernels with the fused Taylor epilogue."""
    qbits = qb.lattices.random_register(n, 0.9 * qb.blockade_radius(OMEGA_MAX), seed=n)
    amp, det = sweep(6, 2)
    calc = qb.B200(qbits, amp, det)
    out = calc.calculate()
    assert calc.metrics["hpsi_passes"] == 1  # one fused launch (two sub-passes through L2)
    ref = oracle_run(qbits, amp, det, method="expm_multiply")["state"]
    assert orc.infidelity(out["state"], ref) <= INFID_TOL
    assert np.max(np.abs(out["state"] - ref)) <= OBS_TOL
    calc.close()


def test_multi_segment_signals_and_tight_lattice():
    """Signals with several segments of different dt; 0.5 r_b spacing (max |H| dt >> 1)."""
    qbits = qb.lattices.square(0.5 * qb.blockade_radius(OMEGA_MAX), 3, 2)
    amp = qb.Signals([qb.Signal(np.linspace(0, OMEGA_MAX, 8), 80), qb.Signal([OMEGA_MAX], 300)])
    det = qb.Signals([qb.Signal(np.linspace(-5, 5, 8), 80), qb.Signal([5.0], 300)])
    calc = qb.B200(qbits, amp, det)
    out = calc.calculate()
    ref = orc.calculate(qbits.positions, [(s.values, s.duration) for s in amp], [(s.values, s.duration) for s in det],
                        method="eigh")["state"]
    assert orc.infidelity(out["state"], ref) <= INFID_TOL
    assert np.max(np.abs(out["state"] - ref)) <= OBS_TOL
    calc.close()


def test_rk4_method_is_fourth_order():
    qbits = qb.lattices.square(0.9 * qb.blockade_radius(OMEGA_MAX), 2, 3)
    amp, det = sweep(50, 1)
    ref = oracle_run(qbits, amp, det, method="eigh")["state"]
    calc = qb.B200(qbits, amp, det, method="rk4")
    out = calc.calculate()
    assert calc.metrics["hpsi_calls"] == 4 * 50
    assert orc.infidelity(out["state"], ref) < 1e-9  # RK4 is NOT at the 1e-10 gate in general
    calc.close()


# ---- e_ops ------------------------------------------------------------------------------------
class _Op:  # duck-typed qse.Operator
    def __init__(self, operator, indices, nqbits, coef=1.0):
        self.operator = [operator] * len(indices) if isinstance(operator, str) else operator
        self.indices, self.nqbits, self.coef = list(indices), nqbits, coef


class _Ops:  # duck-typed qse.Operators
    def __init__(self, operator_list):
        self.operator_list = operator_list


def test_expectations_per_step():
    qbits = qb.lattices.chain(0.9 * qb.blockade_radius(OMEGA_MAX), 5)
    n = 5
    amp, det = sweep(12, 5)
    calc = qb.B200(qbits, amp, det)
    h_i, h_f = calc.get_hamiltonian(0.0, -2 * OMEGA_MAX), calc.get_hamiltonian(OMEGA_MAX, 2 * OMEGA_MAX)
    stagger = _Ops([_Op("Z", [q], n, (-1.0) ** q) for q in range(n)])
    yz = _Op(["Y", "Z"], [1, 3], n, 0.5)
    nn = _Op("N", [0, 2], n, 2.0)
    out = calc.calculate(e_ops=[h_i, h_f, stagger, yz, nn, ("H", 1.0, 0.5)])
    assert out["expectations"].shape == (12, 6)
    v = orc.interaction_matrix(qbits.positions)
    e = orc.diagonal_energies(v, n)
    fns = [lambda p: orc.energy(p, e, n, 0.0, -2 * OMEGA_MAX),
           lambda p: orc.energy(p, e, n, OMEGA_MAX, 2 * OMEGA_MAX),
           lambda p: orc.pauli_expectation(p, n, [((-1.0) ** q, [("Z", q)]) for q in range(n)]).real,
           lambda p: orc.pauli_expectation(p, n, [(0.5, [("Y", 1), ("Z", 3)])]).real,
           lambda p: orc.pauli_expectation(p, n, [(2.0, [("N", 0), ("N", 2)])]).real,
           lambda p: orc.energy(p, e, n, 1.0, 0.5)]
    ref = oracle_run(qbits, amp, det, method="eigh", e_ops=fns)
    scale = max(1.0, np.max(np.abs(ref["expectations"])))
    assert np.max(np.abs(out["expectations"] - ref["expectations"])) <= OBS_TOL * scale
    assert orc.infidelity(out["state"], ref["state"]) <= INFID_TOL
    calc.close()


# ---- K6/K7: observables -----------------------------------------------------------------------
def test_observables_against_reference_golden(golden):
    """Outputs of the reference's own qse.magnetic functions (tests/golden/magnetic.npz)."""
    g = golden("magnetic.npz")
    for n in g["sizes"]:
        n = int(n)
        psi = g[f"n{n}_state"]
        assert np.max(np.abs(qb.magnetic.get_spins(psi, n) - g[f"n{n}_spins"])) <= OBS_TOL
        assert np.max(np.abs(qb.magnetic.get_sisj(psi, n) - g[f"n{n}_sisj"])) <= OBS_TOL
        assert np.max(np.abs(qb.magnetic.get_number_operator(psi, n) - g[f"n{n}_number"])) <= OBS_TOL
    for bits in ("1", "00", "001", "101011"):  # tests/qse/magnetic_test.py:46-53 — MSB-first indexing
        n = len(bits)
        psi = np.zeros(1 << n, dtype=complex)
        psi[int(bits, 2)] = 1.0
        assert np.array_equal(qb.magnetic.get_number_operator(psi, n), np.array([int(c) for c in bits], dtype=float))


@pytest.mark.parametrize("n", [10, 14])
def test_observables_against_oracle(n):
    psi = random_state(n, 1000 + n)
    with lib.Engine(n) as eng:
        eng.set_state(psi)
        assert abs(eng.norm() - 1.0) < 1e-12
        assert np.max(np.abs(eng.spins() - orc.get_spins(psi, n))) <= 1e-12
        assert np.max(np.abs(eng.number() - orc.get_number_operator(psi, n))) <= 1e-12
        if n <= 10:
            assert np.max(np.abs(eng.sisj() - orc.get_sisj(psi, n))) <= 1e-12
        assert np.max(np.abs(eng.probabilities() - np.abs(psi) ** 2)) <= 1e-16


def test_topk_bitstrings():
    n = 14
    psi = random_state(n, 4)
    psi[5] = psi[77]  # an exact tie: lower index first
    p = np.abs(psi) ** 2
    with lib.Engine(n) as eng:
        eng.set_state(psi)
        idx, prob = eng.topk(10)
    order = np.lexsort((np.arange(p.size), -p))[:10]
    assert np.array_equal(idx, order.astype(np.uint64))
    assert np.allclose(prob, p[order], rtol=4e-16, atol=0)  # |c|^2 contracted with FMA on the device


# ---- errors ------------------------------------------------------------------------------------
def test_error_behaviour():
    with lib.Engine(4) as eng:
        with pytest.raises(lib.QsbError, match="set_interaction"):
            eng.apply_h(1.0, 0.0)
        with pytest.raises(lib.QsbError, match="unknown option"):
            eng.set_option("nope", 1)
        with pytest.raises(lib.QsbError, match="overlap"):
            eng.expect_pauli([1.0], [1], [1], [0], [0])
    with pytest.raises(Exception, match="same duration"):
        qb.B200(qb.lattices.chain(5.0, 3), qb.Signal([1.0], 10), qb.Signal([0.0], 20))
    calc = qb.B200(None, qb.Signal([1.0], 10), qb.Signal([0.0], 10))
    with pytest.raises(Exception, match="No qbits"):
        calc.calculate()
    calc.qbits = qb.lattices.chain(5.0, 3)  # attach later, as README.md:50-52
    assert calc.calculate()["state"].shape == (8, 1)
    first = calc.statevector.copy()
    calc.qbits = qb.lattices.chain(9.0, 4)  # re-attach: H must be rebuilt (reference goes stale)
    assert calc.calculate()["state"].shape == (16, 1)
    qbits = qb.Qbits(np.array([[0.0, 0.0], [7.0, 0.0]]), calculator=calc)  # Qbits(calculator=...) hook
    assert calc.qbits is qbits
    assert first.shape == (8,)
    calc.close()


# ---- full-size properties (BASELINE config 2: 25 qubits) ----------------------------------------
def test_full_size_25_qubits_properties():
    n = 25
    qbits = qb.lattices.square(0.9 * qb.blockade_radius(OMEGA_MAX), 5, 5)
    with lib.Engine(n) as eng:
        eng.set_interaction(qb.interaction_matrix(qbits.positions))
        assert eng.metrics()["hpsi_passes"] == 2  # fused super-block launch + one plain launch
        # (1) |0..0> is an eigenvector of the diagonal part: H|0> = (omega/2) sum_q |1_q>
        eng.reset_state()
        y = eng.apply_h(2.0, 5.0)
        nz = np.flatnonzero(y)
        assert np.array_equal(nz, np.sort(1 << np.arange(n))) and np.all(y[nz] == 1.0)
        # (2) exact round trip: evolve forward then backward in time returns to the start
        om, de = np.array([3.0, 6.0]), np.array([-4.0, 2.0])
        eng.evolve_schedule(om, de, np.array([0.002, 0.002]))
        assert abs(eng.norm() - 1.0) < 1e-12
        mid = eng.number()
        assert np.all(mid > 0) and np.all(mid < 1)
        eng.evolve_schedule(om[::-1], de[::-1], np.array([-0.002, -0.002]))
        back = eng.get_state()
        assert 1.0 - abs(back[0]) ** 2 <= INFID_TOL
        # (3) energy is real and equals <psi|H|psi> from spins-like reductions on a product state
        eng.reset_state()
        assert abs(eng.energy(2.0, 5.0)) < 1e-15
    # (4) the two independent kernels agree on a random full-size state
    rng = np.random.default_rng(25)
    psi = (rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)) / np.sqrt(2.0 ** (n + 1))
    outs = []
    for path in (0, 2):
        with lib.Engine(n) as eng:
            eng.set_option("hpsi_path", path)
            eng.set_interaction(qb.interaction_matrix(qbits.positions))
            eng.set_state(psi)
            outs.append(eng.apply_h(OMEGA_MAX, -1.5))
    assert np.max(np.abs(outs[0] - outs[1])) <= 1e-13 * np.max(np.abs(outs[1]))


def test_bench_workload_25_qubits_vs_c_port():
    """The exact workload bench.py times (BASELINE config 2: square 5x5, 25 qubits, adiabatic ramp at
    1 ns per value) against the oracle's C port at full size: H psi on a random state, then three
    mid-ramp sweep steps from that state -- infidelity <= 1e-10, max |delta psi| <= 1e-9 (phase
    included), spins <= 1e-9.  qse/calc/qutip.py:34-53."""
    import bench
    from oracle import c_port

    n = 25
    qbits, amp, det = bench.workload(n, 23)
    om, de, dt = qb.flatten_schedule(qb.Signals([amp]), qb.Signals([det]))
    v = orc.interaction_matrix(qbits.positions)
    e = c_port.build_diagonal(v, n)
    rng = np.random.default_rng(2025)
    psi = (rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)) / np.sqrt(2.0 ** (n + 1))
    psi /= np.linalg.norm(psi)
    c_port.set_threads(os.cpu_count() or 1)
    sel = slice(10, 13)  # mid-ramp: Omega ~ 2.9..3.4, delta ~ -1.1..+1.1
    want_h = c_port.apply_h(psi, e, n, om[12], de[12])
    want, n_hpsi = c_port.evolve_schedule(psi, e, n, om[sel], de[sel], dt[sel])
    with lib.Engine(n) as eng:
        eng.set_interaction(qb.interaction_matrix(qbits.positions))
        assert np.array_equal(eng.diagonal(), e)  # the 256 MiB diagonal, bit for bit
        eng.set_state(psi)
        got_h = eng.apply_h(om[12], de[12])
        assert np.max(np.abs(got_h - want_h)) <= 1e-13 * np.max(np.abs(want_h))
        eng.evolve_schedule(om[sel], de[sel], dt[sel])
        got = eng.get_state()
        spins = eng.spins()
    assert orc.infidelity(got, want) <= INFID_TOL
    assert np.max(np.abs(got - want)) <= OBS_TOL
    assert np.max(np.abs(spins - c_port.spins(want, n))) <= OBS_TOL
    assert n_hpsi >= 3


# ---- BASELINE configs 3 and 5 at full size: size-independent properties --------------------------
def test_config5_random_28_qubits_bitstrings():
    """BASELINE config 5: 28 random-position qubits (unit-disk register), detuning sweep, bit-string
    probabilities read out on the device (the host never sees the 4 GiB state)."""
    n = 28
    reg = qb.lattices.random_register(n, 0.9 * qb.blockade_radius(OMEGA_MAX), seed=20261019)
    t = 4
    amp = qb.Signal(np.full(t, OMEGA_MAX), 4 * t)
    det = qb.Signal(np.linspace(-OMEGA_MAX, 2 * OMEGA_MAX, t), 4 * t)
    calc = qb.B200(reg, amp, det, return_state=False)
    out = calc.calculate()
    assert "state" not in out
    eng = calc._engine
    assert abs(eng.norm() - 1.0) < 1e-11
    top = calc.bitstring_probabilities(8)
    probs = [p for _, p in top]
    assert all(len(b) == n for b, _ in top)
    assert probs == sorted(probs, reverse=True) and 0 < sum(probs) <= 1.0 + 1e-12
    assert top[0][0] == "0" * n  # 16 ns is far too short to leave |0...0> behind
    number = calc.get_number_operator()
    assert np.all(number > 0) and np.all(number < 0.5)
    assert np.max(np.abs(calc.spins[:, 2] - (1 - 2 * number))) < 1e-12  # <Z> = 1 - 2<n>
    # time reversal on the device returns to |0...0>
    om, de, dt = qb.flatten_schedule(calc.amplitude, calc.detuning)
    eng.evolve_schedule(om[::-1], de[::-1], -dt[::-1])
    idx, p = eng.topk(1)
    assert int(idx[0]) == 0 and 1.0 - p[0] <= INFID_TOL
    calc.close()


def test_config3_kagome_30_qubits_quench():
    """BASELINE config 3: kagome 5x2 (30 qubits) quench at constant Omega, delta = 0 on ONE GPU
    (16 GiB state): norm, symmetry of the lattice in <n_i>, and an exact time-reversal round trip."""
    n = 30
    qbits = qb.lattices.kagome(0.9 * qb.blockade_radius(OMEGA_MAX), 5, 2)
    assert qbits.nqbits == n
    with lib.Engine(n) as eng:
        eng.set_interaction(qb.interaction_matrix(qbits.positions))
        m = eng.metrics()
        assert m["hpsi_passes"] == 2 and m["hpsi_bytes_per_amp"] == 88.0
        om, de, dt = np.full(2, OMEGA_MAX), np.zeros(2), np.full(2, 0.002)
        eng.evolve_schedule(om, de, dt)
        assert abs(eng.norm() - 1.0) < 1e-11
        number = eng.number()
        assert np.all(number > 0) and np.all(number < 0.1)
        # sites related by the inversion symmetry of the finite kagome patch evolve identically
        pos = qbits.positions
        centre = pos.mean(axis=0)
        for i in range(n):
            j = int(np.argmin(np.linalg.norm(pos - (2 * centre - pos[i]), axis=1)))
            if np.linalg.norm(pos[j] - (2 * centre - pos[i])) < 1e-9:
                assert abs(number[i] - number[j]) < 1e-12
        eng.evolve_schedule(om, de, -dt)
        idx, p = eng.topk(1)
        assert int(idx[0]) == 0 and 1.0 - p[0] <= INFID_TOL


# ---- launch-bound sizes: whole-schedule kernel vs the multi-kernel stepper ------------------------
@pytest.mark.parametrize("n", [1, 5, 9, 12])
def test_small_schedule_kernel_matches_general_path(n):
    reg = qb.lattices.random_register(n, 0.8 * qb.blockade_radius(OMEGA_MAX), seed=40 + n) if n > 1 \
        else qb.Qbits(np.zeros((1, 2)))
    t = 30
    om = np.linspace(0.0, OMEGA_MAX, t)
    de = np.linspace(-2 * OMEGA_MAX, 2 * OMEGA_MAX, t)
    dt = np.full(t, 0.004)
    dt[7] = 1.0  # one long segment: sub-stepping inside the kernel
    states, exps, hpsi = [], [], []
    for small in (1, 0):
        with lib.Engine(n) as eng:
            eng.set_option("small_kernel", small)
            eng.set_interaction(qb.interaction_matrix(reg.positions))
            eops = lib.PackedEops([("energy", 0.0, -2 * OMEGA_MAX), ("energy", OMEGA_MAX, 2 * OMEGA_MAX)])
            m0 = eng.metrics()
            exps.append(eng.evolve_schedule(om, de, dt, eops))
            m1 = eng.metrics()
            states.append(eng.get_state())
            hpsi.append(m1["hpsi_calls"] - m0["hpsi_calls"])
            launches = m1["kernel_launches"] - m0["kernel_launches"]
            assert (launches == 1) == bool(small)
            assert m1["substeps"] - m0["substeps"] > t
    assert np.max(np.abs(states[0] - states[1])) <= 1e-12
    assert np.max(np.abs(exps[0] - exps[1])) <= 1e-10 * max(1.0, np.max(np.abs(exps[1])))
    # the general path spends 2 H psi per step on the e_ops and may overshoot a term (it only reads
    # norms back from term m_prev - 1 on); the in-kernel stepper tests every term
    assert 0.5 * (hpsi[1] - 2 * t) <= hpsi[0] <= hpsi[1] - 2 * t


def test_small_schedule_kernel_reports_failure():
    with lib.Engine(3) as eng:
        eng.set_interaction(np.zeros((3, 3)))
        eng.set_option("max_terms", 2)
        with pytest.raises(lib.QsbError, match="did not reach tol"):
            eng.evolve_schedule(np.array([5.0]), np.array([1.0]), np.array([0.5]))


@pytest.mark.parametrize("opts", [{"tile_bits": 11}, {"tensor_map": 0}, {"tile_bits": 11, "tensor_map": 0, "sb_bits": 16, "sb_lag": 2},
                                  {"sb_bits": 21, "sb_lag": 1}])
def test_kernel_options_do_not_change_results(opts):
    """Every tuning knob of the H psi kernel (tile size, TMA flavour, super-block size / lead) is a
    pure scheduling choice: same result as the untiled kernel, for H psi and for a short sweep."""
    n = 22
    pos = random_positions(n, 9)
    psi = random_state(n, 9)
    with lib.Engine(n) as eng:
        eng.set_interaction(qb.interaction_matrix(pos))
        eng.set_state(psi)
        eng.set_option("hpsi_path", 1)
        want = eng.apply_h(3.0, -1.0)
        eng.evolve_schedule(np.array([2.0, 4.0]), np.array([-3.0, 3.0]), np.array([0.002, 0.002]))
        ref_state = eng.get_state()
        eng.set_option("hpsi_path", 0)
        for key, value in opts.items():
            eng.set_option(key, value)
        eng.set_state(psi)
        got = eng.apply_h(3.0, -1.0)
        assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))
        eng.evolve_schedule(np.array([2.0, 4.0]), np.array([-3.0, 3.0]), np.array([0.002, 0.002]))
        assert np.max(np.abs(eng.get_state() - ref_state)) <= 1e-13


def test_sampling_follows_the_born_rule():
    n = 10
    psi = random_state(n, 321)
    p = np.abs(psi) ** 2
    with lib.Engine(n) as eng:
        eng.set_state(psi)
        shots = 400_000
        a = eng.sample(shots, seed=7)
        assert np.array_equal(a, eng.sample(shots, seed=7))  # deterministic in the seed
        assert not np.array_equal(a, eng.sample(shots, seed=8))
        freq = np.bincount(a.astype(np.int64), minlength=1 << n) / shots
        # binomial standard error per bin: sqrt(p (1 - p) / shots) <= 5e-5 * ...; allow 6 sigma
        assert np.all(np.abs(freq - p) <= 6 * np.sqrt(p * (1 - p) / shots) + 1e-6)
        basis = np.zeros(1 << n, dtype=complex)
        basis[0b1011001110] = 1.0
        eng.set_state(basis)
        assert np.all(eng.sample(100, seed=1) == 0b1011001110)
    n = 14  # several chunks of 4096 amplitudes
    psi = random_state(n, 5)
    with lib.Engine(n) as eng:
        eng.set_state(psi)
        a = eng.sample(200_000, seed=3)
        chunk_freq = np.bincount(a.astype(np.int64) >> 12, minlength=4) / a.size
        chunk_p = (np.abs(psi) ** 2).reshape(4, -1).sum(axis=1)
        assert np.max(np.abs(chunk_freq - chunk_p)) < 6 * np.sqrt(0.25 / a.size)


def test_calculator_sample_final_state():
    qbits = qb.lattices.chain(0.9 * qb.blockade_radius(OMEGA_MAX), 4)
    calc = qb.B200(qbits, qb.Signal([OMEGA_MAX], 200), qb.Signal([0.0], 200))
    counts = calc.sample_final_state(shots=5000, seed=11)
    assert sum(counts.values()) == 5000 and all(len(b) == 4 for b in counts)
    p = np.abs(calc.statevector) ** 2
    top = max(counts, key=counts.get)
    assert int(top, 2) == int(np.argmax(p))
    calc.close()


# ---- edge cases ---------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [3, 13])
def test_edge_schedules(n):
    reg = qb.lattices.random_register(n, 6.0, seed=n) if n > 1 else qb.Qbits(np.zeros((1, 2)))
    v = orc.interaction_matrix(reg.positions)
    e = orc.diagonal_energies(v, n)
    psi = random_state(n, 77)
    with lib.Engine(n) as eng:
        eng.set_interaction(qb.interaction_matrix(reg.positions))
        eng.set_state(psi)
        # empty schedule and zero-length segments leave the state untouched (bit for bit)
        assert eng.evolve_schedule(np.zeros(0), np.zeros(0), np.zeros(0)) is None
        eng.evolve_schedule(np.array([3.0, 1.0]), np.array([1.0, -2.0]), np.zeros(2))
        assert np.array_equal(eng.get_state(), psi)
        # omega = 0: H is diagonal, the propagator is a pure phase per basis state
        eng.evolve_schedule(np.array([0.0]), np.array([2.5]), np.array([0.003]))
        diag = e - 2.5 * orc.popcounts(n)
        assert np.max(np.abs(eng.get_state() - np.exp(-1j * diag * 0.003) * psi)) <= 1e-12
        # e_ops given but empty: (steps, 0) expectations
        out = eng.evolve_schedule(np.array([1.0, 1.0]), np.zeros(2), np.full(2, 0.001), lib.PackedEops([]))
        assert out is None
    # no interaction at all (far-apart atoms): product of single-qubit rotations
    with lib.Engine(n) as eng:
        eng.set_interaction(np.zeros((n, n)))
        eng.evolve_schedule(np.array([4.0]), np.array([0.0]), np.array([0.1]))
        one = orc.exact_single_qubit(4.0, 0.0, 0.1)
        want = one
        for _ in range(n - 1):
            want = np.kron(want, one)
        assert np.max(np.abs(eng.get_state() - want)) <= 1e-12


def test_calculate_with_empty_e_ops_and_reuse():
    qbits = qb.lattices.chain(0.9 * qb.blockade_radius(OMEGA_MAX), 3)
    calc = qb.B200(qbits, qb.Signal([1.0, 2.0, 3.0], 30), qb.Signal([0.0, 0.5, 1.0], 30))
    out = calc.calculate(e_ops=[])
    assert out["expectations"].shape == (3, 0)
    first = out["state"].copy()
    again = calc.calculate()  # same schedule from |0...0> again: identical result, fresh dict
    assert np.array_equal(again["state"], first) and again is not out
    calc.close()


def test_reference_notebook_workload():
    """docs/use-cases/adiabatic.ipynb (ring of 8, Omega 0->2, delta -2->0, 1 ns values, e_ops H_i, H_f),
    shortened to 400 values, through examples/adiabatic_ring.py against the exact propagator."""
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "adiabatic_ring.py")
    spec = importlib.util.spec_from_file_location("adiabatic_ring", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    t = 400
    out = mod.main(t)
    qbits = qb.lattices.ring(qb.blockade_radius(1.0), 8)
    e = orc.diagonal_energies(orc.interaction_matrix(qbits.positions), 8)
    ref = orc.calculate(qbits.positions, [(np.linspace(0.0, 2.0, t), t)], [(np.linspace(-2.0, 0.0, t), t)], method="eigh",
                        e_ops=[lambda s: orc.energy(s, e, 8, 0.0, -2.0), lambda s: orc.energy(s, e, 8, 2.0, 0.0)])
    assert orc.infidelity(out["state"], ref["state"]) <= INFID_TOL
    assert np.max(np.abs(out["expectations"] - ref["expectations"])) <= OBS_TOL


def test_checkpoint_resume(tmp_path):
    """A sweep interrupted after k values and resumed from the checkpoint equals the uninterrupted one."""
    qbits = qb.lattices.triangular(0.9 * qb.blockade_radius(OMEGA_MAX), 3, 3)
    amp, det = sweep(24, 2)
    whole = qb.B200(qbits, amp, det)
    ref = whole.calculate(e_ops=[whole.get_hamiltonian(OMEGA_MAX, 0.0)])
    whole.close()
    prefix = str(tmp_path / "ckpt")
    part = qb.B200(qbits, amp, det)
    head = part.calculate(e_ops=[part.get_hamiltonian(OMEGA_MAX, 0.0)], stop_after=10)
    assert head["expectations"].shape == (10, 1)
    part.save_checkpoint(prefix, next_step=10)
    part.close()
    resumed = qb.B200(qbits, amp, det)
    tail = resumed.calculate(e_ops=[resumed.get_hamiltonian(OMEGA_MAX, 0.0)], resume_from=prefix)
    assert tail["expectations"].shape == (14, 1)
    assert np.max(np.abs(tail["state"] - ref["state"])) <= 1e-13
    assert np.max(np.abs(np.vstack([head["expectations"], tail["expectations"]]) - ref["expectations"])) <= 1e-11
    other = qb.B200(qb.lattices.chain(5.0, 9), amp, det)
    with pytest.raises(Exception, match="different qubit positions"):
        other.calculate(resume_from=prefix)
    other.close()
    resumed.close()
