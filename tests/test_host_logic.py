"""Host-side logic of the calculator that needs no GPU: schedule flattening, pulse checks,
interaction matrix, Pauli mask packing, the structure-factor DFT, the shard plan."""
import numpy as np
import pytest

import qse_b200 as qb
from oracle import rydberg_oracle as orc
from qse_b200 import calculator as calc_mod
from qse_b200.distributed import ShardPlan


def test_interaction_matrix_bit_identical_to_reference_terms(golden):
    g = golden("interaction.npz")
    for c in range(int(g["n_cases"])):
        v = qb.interaction_matrix(g[f"{c}_positions"])
        pairs = np.argwhere(v != 0.0)
        assert np.array_equal(pairs, g[f"{c}_pairs"])
        assert np.array_equal(v[pairs[:, 0], pairs[:, 1]], g[f"{c}_coefs"])
        assert np.array_equal(v, orc.interaction_matrix(g[f"{c}_positions"]))


def test_flatten_schedule_takes_dt_from_amplitude_only():
    amp = qb.Signals([qb.Signal(np.linspace(0, 1, 5), 20), qb.Signal([3.0], 400)])
    det = qb.Signals([qb.Signal(np.linspace(-1, 1, 5), 10), qb.Signal([0.5], 410)])
    calc_mod._check_pulses(amp, det)  # total durations equal (420), per-segment ones are not checked
    om, de, dt = qb.flatten_schedule(amp, det)
    assert np.array_equal(om, [0, 0.25, 0.5, 0.75, 1.0, 3.0])
    assert np.array_equal(dt, [0.004] * 5 + [0.4])
    o2, d2, t2 = orc.flatten_schedule([(s.values, s.duration) for s in amp], [(s.values, s.duration) for s in det])
    assert np.array_equal(om, o2) and np.array_equal(de, d2) and np.array_equal(dt, t2)


def test_check_pulses_messages():
    s = qb.Signal
    with pytest.raises(Exception, match="must have the same duration"):
        calc_mod._check_pulses(qb.Signals([s([1.0], 10)]), qb.Signals([s([1.0], 20)]))
    with pytest.raises(Exception, match="same number of signals"):
        calc_mod._check_pulses(qb.Signals([s([1.0], 10), s([1.0], 10)]), qb.Signals([s([1.0], 20)]))
    with pytest.raises(Exception, match="same amount of values"):
        calc_mod._check_pulses(qb.Signals([s([1.0, 2.0], 10)]), qb.Signals([s([1.0], 10)]))


def test_pauli_masks_msb_first():
    assert calc_mod._pauli_masks(["X"], [0], 3) == (0b100, 0, 0, 0)
    assert calc_mod._pauli_masks(["Y", "Z"], [2, 4], 5) == (0, 0b00100, 0b00001, 0)
    assert calc_mod._pauli_masks(["N", "N"], [0, 5], 6) == (0, 0, 0, 0b100001)
    with pytest.raises(Exception, match="must be one of"):
        calc_mod._pauli_masks(["Q"], [0], 2)
    with pytest.raises(Exception, match="only once"):
        calc_mod._pauli_masks(["X", "Z"], [1, 1], 2)


def test_structure_factor_matches_oracle_dft():
    rng = np.random.default_rng(5)
    l1, l2, l3 = 3, 2, 2
    n = l1 * l2 * l3
    s = rng.normal(size=(n, n))
    s = s + s.T
    q = qb.Qbits(np.zeros((n, 3)))
    got = qb.magnetic.structure_factor_from_sij(l1, l2, l3, q, s)
    assert got.shape == (l1, l2, l3)
    assert np.max(np.abs(got - orc.structure_factor_from_sij(l1, l2, l3, s))) < 1e-13


def test_shard_plan():
    plan = ShardPlan(5, 4)
    assert (plan.p, plan.nl, plan.local_dim) == (2, 3, 8)
    assert plan.owner(0b10110) == 0b10 and plan.local_index(0b10110) == 0b110
    assert plan.partner(0b10, 0) == 0b00 and plan.partner(0b10, 1) == 0b11
    assert plan.shard_slice(3) == slice(24, 32)
    v = np.zeros((5, 5))
    v[0, 1] = 7.0
    assert plan.rank_diagonal_offset(0b11, v) == 7.0 and plan.rank_diagonal_offset(0b10, v) == 0.0
    with pytest.raises(ValueError):
        ShardPlan(5, 3)


def test_sharded_algorithm_reproduces_full_hpsi():
    """The shard decomposition libqsb implements, restated with numpy shards against the oracle:
    local H psi with rank-dependent diagonal + (omega/2) * partner shard per sharded qubit."""
    rng = np.random.default_rng(11)
    n, world = 6, 4
    plan = ShardPlan(n, world)
    pos = rng.uniform(0, 20, size=(n, 2))
    v = orc.interaction_matrix(pos)
    e = orc.diagonal_energies(v, n)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    omega, delta = 2.3, -1.1
    full = orc.apply_hamiltonian(psi, e, n, omega, delta)
    for r in range(world):
        sl = plan.shard_slice(r)
        local = psi[sl]
        k = np.arange(plan.local_dim)
        pop = np.array([bin(int(x)).count("1") for x in k]) + plan.rank_popcount(r)
        y = (e[sl] - delta * pop) * local
        for q in range(plan.nl):
            y = y + 0.5 * omega * local[k ^ (1 << q)]
        for g in plan.sharded_qubits():
            y = y + 0.5 * omega * psi[plan.shard_slice(plan.partner(r, g))]
        assert np.max(np.abs(y - full[sl])) < 1e-12


@pytest.mark.parametrize("tile_bits", [12, 11])
@pytest.mark.parametrize("sb_bits", [0, 13, 18, 21])
def test_hpsi_launch_plan_covers_every_qubit_once(tile_bits, sb_bits):
    """The planner of the H psi launches (host-only entry point, no GPU): over all sub-passes every
    local index bit is an ACTIVE tile bit exactly once, tiles hold 2^tile_bits amplitudes, rows are
    at least 128 bytes, and at most two launches touch HBM up to 30 local qubits (auto plan)."""
    from qse_b200 import lib

    for nl in range(13, 34):
        plan = lib.describe_plan(nl, sb_bits, tile_bits)
        tb, sb = plan["tile_bits"], plan["sb_bits"]
        assert tb == tile_bits and tb < sb <= min(nl, 21)
        active = list(range(tb))  # FIRST sub-pass: the contiguous tile
        for geo in [plan["inner"]] + plan["plain"]:
            assert geo["lb"] + geo["hb"] == tb and geo["lb"] >= 3 and geo["hb"] >= 1  # >= 128-byte rows
            active += list(range(geo["hb0"], geo["hb0"] + geo["hb"]))
        assert sorted(active) == list(range(nl)), (nl, plan)
        assert plan["inner"]["hb0"] == tb and plan["inner"]["hb"] == sb - tb
        if sb_bits == 0 and tile_bits == 12:
            assert len(plan["plain"]) == (0 if nl <= 21 else 1 if nl <= 30 else 2)


@pytest.mark.parametrize("n_sb,n_a,lag,sb_xor", [(1, 4, 1, 0), (2, 8, 1, 0), (8, 4, 3, 0), (8, 4, 8, 5), (16, 2, 2, 12), (32, 1, 3, 7),
                                                 (8, 4, 0, 0), (4, 2, 0, 3), (1, 4, 0, 0)])
def test_fused_launch_schedule_invariants(n_sb, n_a, lag, sb_xor):
    """The ticket order the producer warps follow (evaluated on the host through the C ABI): every
    (sub-pass, tile) is handed out exactly once, and every INNER tile of a super-block comes after ALL
    FIRST tiles of that block, by at least lag - 1 whole phases — what makes the in-kernel dependency
    wait deadlock-free (a CTA only waits for earlier tickets) and short."""
    import ctypes

    from qse_b200 import lib

    dll = lib.load()
    seen, last_first, first_inner = set(), {}, {}
    kind, tile = ctypes.c_int(), ctypes.c_uint64()
    for ticket in range(2 * n_a * n_sb):
        assert dll.qsb_describe_ticket(ticket, n_a, n_sb, lag, sb_xor, ctypes.byref(kind), ctypes.byref(tile)) == 0
        key = (kind.value, tile.value)
        assert key not in seen and 0 <= tile.value < n_a * n_sb and kind.value in (0, 1)
        seen.add(key)
        sb = tile.value // n_a
        if kind.value == 0:
            last_first[sb] = ticket
        else:
            first_inner.setdefault(sb, ticket)
    assert len(seen) == 2 * n_a * n_sb
    for sb in range(n_sb):
        assert first_inner[sb] > last_first[sb]
        gap = first_inner[sb] - last_first[sb] - 1
        assert gap >= max(0, (min(lag, n_sb) - 1) * n_a - (n_a if sb >= n_sb - lag else 0)) or n_sb <= lag
    assert dll.qsb_describe_ticket(2 * n_a * n_sb, n_a, n_sb, lag, sb_xor, ctypes.byref(kind), ctypes.byref(tile)) == -1
