"""Host-side helpers next to the hot path: unit-disk MIS checks for BASELINE
config 5, the position-general structure factor, and a11's grid structure factor against the LITERAL reference
loop (qse/magnetic.py:279-291) with its `tuple @ Cell` defect repaired."""
import numpy as np
import pytest

import qse_b200 as qb
from qse_b200 import magnetic, mis


def test_unit_disk_graph_and_independent_sets():
    reg = qb.lattices.square(1.0, 3, 3)
    adj = mis.unit_disk_graph(reg, 1.01)  # nearest neighbours only
    assert adj.sum() == 2 * 12 and not adj.diagonal().any() and np.array_equal(adj, adj.T)
    checker = "101010101"  # the 5-site checkerboard of the 3x3 grid: the maximum independent set
    assert mis.is_independent_set(checker, adj) and mis.is_maximal_independent_set(checker, adj)
    assert mis.violations(checker, adj) == 0
    assert not mis.is_independent_set("110000000", adj) and mis.violations("110100000", adj) == 2
    assert mis.is_independent_set("100000000", adj) and not mis.is_maximal_independent_set("100000000", adj)
    assert mis.is_independent_set("000000000", adj)
    g = mis.greedy_independent_set(adj)
    assert mis.is_maximal_independent_set(g, adj) and g.sum() >= 4
    assert np.array_equal(mis.index_to_bits(0b101010101, 9), mis.bits_of(checker))
    score = mis.score_readout([(checker, 0.5), ("110000000", 0.2), ("010101010", 0.1)], adj)
    assert score["p_independent"] == pytest.approx(0.6) and score["largest_independent_set"] == 5
    assert score["p_largest"] == pytest.approx(0.5)
    with pytest.raises(ValueError):
        mis.is_independent_set("101", adj)


def test_config5_register_graph():
    """BASELINE config 5: 28 random-position qubits; the blockade graph at r_b is what the sweep solves."""
    rb = qb.blockade_radius(2 * np.pi)
    reg = qb.lattices.random_register(28, 0.9 * rb, seed=20261019)
    adj = mis.unit_disk_graph(reg.positions, rb)
    assert adj.shape == (28, 28) and adj.any()
    g = mis.greedy_independent_set(adj)
    assert mis.is_maximal_independent_set(g, adj)
    assert mis.is_independent_set("0" * 28, adj)


def reference_loop_repaired(L1, L2, L3, cell, s_ij):
    """qse/magnetic.py:279-291 line by line, with `Ri @ qbits.cell` / `Qi @ qbits.cell.reciprocal()` evaluated on
    the cell's lattice vectors (the reference passes the Cell object to `@`, which raises TypeError at :284)."""
    n = L1 * L2 * L3
    struc_fac = np.empty((L1, L2, L3), dtype=complex)
    Qi = tuple((i1, i2, i3) for i1 in range(L1) for i2 in range(L2) for i3 in range(L3))
    Ri = Qi
    A = np.asarray(cell, dtype=float)
    Rvecs = np.array(Ri) @ A
    Qvecs = np.array(Qi) @ (2 * np.pi * np.linalg.inv(A.T)) / (L1, L2, L3)
    for q, qvec in zip(Qi, Qvecs):
        exri = np.exp(1j * Rvecs @ qvec)
        exrj = exri.conj()
        struc_fac[q] = np.einsum("ij,i,j->", s_ij, exri, exrj)
    struc_fac /= n**2
    return struc_fac.real.copy()


@pytest.mark.parametrize("shape,cell", [((3, 3, 1), np.diag([1.0, 1.0, 1.0])), ((2, 4, 3), np.diag([0.7, 1.3, 2.1])),
                                        ((4, 1, 2), np.diag([5.0, 2.0, 3.0]))])
def test_grid_structure_factor_against_the_literal_reference_loop(shape, cell):
    L1, L2, L3 = shape
    n = L1 * L2 * L3
    rng = np.random.default_rng(n)
    s = rng.normal(size=(n, n))
    s = 0.5 * (s + s.T)
    np.fill_diagonal(s, 1.0)
    sites = np.array([(a, b, c) for a in range(L1) for b in range(L2) for c in range(L3)], dtype=float)
    qbits = qb.Qbits(sites @ cell, cell=cell)
    want = reference_loop_repaired(L1, L2, L3, cell, s)
    got = magnetic.structure_factor_from_sij(L1, L2, L3, qbits, s)
    assert got.shape == (L1, L2, L3) and np.max(np.abs(got - want)) <= 1e-12
    # the position-general form on the same grid of wave vectors and the true positions agrees too
    qv = magnetic.reciprocal_grid(cell, shape)
    assert np.max(np.abs(magnetic.structure_factor(qbits.positions, s, qv).reshape(shape) - want)) <= 1e-12


def test_position_general_structure_factor_properties():
    reg = qb.lattices.triangular(1.0, 4, 3)
    n = reg.nqbits
    s = np.ones((n, n))
    # uniform correlations: all weight at q = 0, S(0) = 1
    q = np.array([[0.0, 0.0], [2 * np.pi / 4, 0.0], [0.3, -1.1]])
    out = magnetic.structure_factor(reg, s, q)
    assert out[0] == pytest.approx(1.0) and out.shape == (3,)
    # S(q) = |sum_i v_i exp(i q.x_i)|^2 / N^2 for a rank-one s_ij = v_i v_j: non-negative, and even in q
    v = np.random.default_rng(0).normal(size=n)
    s1 = np.outer(v, v)
    direct = np.abs(np.exp(1j * q @ reg.positions.T) @ v) ** 2 / n**2
    assert np.max(np.abs(magnetic.structure_factor(reg.positions, s1, q) - direct)) <= 1e-13
    assert np.max(np.abs(magnetic.structure_factor(reg.positions, s1, -q) - direct)) <= 1e-13
    with pytest.raises(ValueError):
        magnetic.structure_factor(reg.positions, s1, np.zeros((2, 3)))


def test_heatmap_forwards_or_explains():
    reg = qb.lattices.chain(1.0, 3)
    with pytest.raises(Exception, match="length of scalar"):
        magnetic.heatmap(reg, [1.0, 2.0])
    try:
        import qse.vis  # noqa: F401
    except Exception:
        with pytest.raises(Exception, match="qbits_heatmap"):
            magnetic.heatmap(reg, [0.1, 0.2, 0.3])
