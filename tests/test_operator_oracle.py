"""The oracle's restatement of operator-sum Hamiltonians (qse.Operators) against an
independent construction -- Kronecker products in the order of Operator.to_qutip
(qse/operator/operator.py:80-92: qubit 0 first) and a dense matrix exponential -- and the host-side mirror of
qse.Operator / qse.Operators / Qbits.compute_interaction_hamiltonian against the reference's own classes."""
import numpy as np
import pytest

import qse_b200 as qb
from oracle import rydberg_oracle as orc

X = np.array([[0, 1], [1, 0]], complex)
Y = np.array([[0, -1j], [1j, 0]], complex)
Z = np.diag([1.0, -1.0]).astype(complex)
N = np.diag([0.0, 1.0]).astype(complex)  # qp.num(2), qse/operator/operator.py:141-142
PAULI = {"X": X, "Y": Y, "Z": Z, "N": N}


def kron_matrix(terms, n):
    h = np.zeros((1 << n, 1 << n), complex)
    for coef, ops in terms:
        mats = [np.eye(2, dtype=complex)] * n
        for op, q in ops:
            mats[q] = PAULI[op]
        m = mats[0]
        for a in mats[1:]:
            m = np.kron(m, a)
        h += coef * m
    return h


def tfim(n, j=1.0, h=0.5):
    """docs/use-cases/time_evo.ipynb cell 3: -J sum Z_i Z_{i+1} - h sum X_i on a chain."""
    qbits = qb.lattices.chain(1.0, n)
    ham = qbits.compute_interaction_hamiltonian(lambda d: -j * np.isclose(d, 1.0), "Z")
    for i in range(n):
        ham += qb.Operator("X", i, n, coef=-h)
    return ham


def test_pauli_apply_matches_kron():
    n = 5
    rng = np.random.default_rng(5)
    psi = rng.normal(size=32) + 1j * rng.normal(size=32)
    terms = [(0.7, [("X", 0)]), (-1.3, [("Y", 2), ("Z", 4)]), (0.4, [("N", 1), ("N", 3)]), (2.0, [("X", 1), ("Y", 0), ("Z", 3)]),
             (0.9, [("Y", 1), ("Y", 4)])]
    assert np.max(np.abs(orc.pauli_apply(psi, n, terms) - kron_matrix(terms, n) @ psi)) < 1e-13


def test_general_propagate_tfim_notebook():
    from scipy.linalg import expm

    n = 6
    terms = orc.operators_as_terms(tfim(n))
    psi0 = np.zeros(1 << n, complex)
    psi0[0] = 1.0
    want = expm(-1j * kron_matrix(terms, n) * 1.0) @ psi0
    assert np.max(np.abs(orc.general_propagate(psi0, n, 1.0, terms=terms) - want)) < 1e-12


def test_general_apply_reduces_to_the_rydberg_form():
    n = 7
    pos = qb.lattices.random_register(n, 6.0, seed=3).positions
    e = orc.diagonal_energies(orc.interaction_matrix(pos), n)
    rng = np.random.default_rng(1)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    want = orc.apply_hamiltonian(psi, e, n, 3.0, -1.5)
    got = orc.general_apply(psi, n, e_int=e, hw_sites=np.full(n, 1.5), d_sites=np.full(n, -1.5))
    assert np.max(np.abs(got - want)) < 1e-12


def test_operator_mirror_matches_reference_classes():
    """qse_b200.Operator / Operators / compute_interaction_hamiltonian expose what qse's classes do."""
    try:
        from oracle import refshim

        qse = refshim.import_reference()
    except Exception as exc:  # the GPU boxes have no /root/reference
        pytest.skip(f"reference not importable: {exc}")
    ref_q = qse.lattices.square(1.0, 2, 3)
    our_q = qb.lattices.square(1.0, 2, 3)
    f = lambda d: 1.7 / d**3  # noqa: E731
    ref_h = ref_q.compute_interaction_hamiltonian(f, ["X", "Y"])
    our_h = our_q.compute_interaction_hamiltonian(f, ["X", "Y"])
    ref_h += qse.Operator("Z", 2, 6, coef=0.25)
    our_h += qb.Operator("Z", 2, 6, coef=0.25)
    assert [(t.operator, list(t.indices), t.nqbits, t.coef) for t in ref_h.operator_list] == \
           [(t.operator, list(t.indices), t.nqbits, t.coef) for t in our_h.operator_list]
    assert qb.operator_terms(our_h, 6) == qb.operator_terms(ref_h, 6)  # the calculator reads either
    with pytest.raises(Exception, match="must be one of"):
        qb.Operator("Q", 0, 3)
    with pytest.raises(Exception, match="between 0 and nqbits-1"):
        qb.Operator("X", 3, 3)
