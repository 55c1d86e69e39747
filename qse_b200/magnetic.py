"""Drop-ins for the ``qse.magnetic`` observables (qse/magnetic.py), evaluated on the GPU.

Same call signatures as the reference functions for the (statevector, nqbits) part; the
``ibasis`` argument of the reference (a materialised 2^N x N bit table, qse/magnetic.py:14-52)
is never needed: bit q of index k is ``(k >> (N-1-q)) & 1``.
"""

from __future__ import annotations

import numpy as np

from . import lib


def _engine_with(statevector, nqbits: int, device: int = 0) -> lib.Engine:
    psi = np.asarray(statevector, dtype=np.complex128).reshape(-1)
    if psi.size != 1 << nqbits:
        raise ValueError("statevector must hold 2**nqbits amplitudes")
    eng = lib.Engine(nqbits, device=device)
    eng.set_state(psi)
    return eng


def get_spins(statevector, nqbits: int, device: int = 0) -> np.ndarray:
    """(N, 3) <X_i>, <Y_i>, <Z_i>  (qse/magnetic.py:119-176)."""
    with _engine_with(statevector, nqbits, device) as eng:
        return eng.spins()


def get_sisj(statevector, nqbits: int, device: int = 0) -> np.ndarray:
    """(N, N) <Z_iZ_j> + (1/3)<X_iX_j + Y_iY_j>, 1 on the diagonal  (qse/magnetic.py:179-243)."""
    with _engine_with(statevector, nqbits, device) as eng:
        return eng.sisj()


def get_number_operator(statevector, nqbits: int, device: int = 0) -> np.ndarray:
    """(N,) <n_i>  (qse/magnetic.py:295-326)."""
    with _engine_with(statevector, nqbits, device) as eng:
        return eng.number()


def structure_factor_from_sij(L1: int, L2: int, L3: int, qbits, s_ij) -> np.ndarray:
    """S[q] = N^-2 sum_ij s_ij exp(i q.(x_i - x_j)) on the L1 x L2 x L3 reciprocal grid.

    Host numpy: N^2 * N_q flops with N <= 40 (kernel K8 of SURVEY.md needs no GPU).  The
    reference implementation (qse/magnetic.py:246-292) raises TypeError at :284; this follows
    its docstring: site i sits at row-major (n1, n2, n3), R_i = n . A, q = m . B / L with
    B = 2 pi inv(A^T), hence q.R = 2 pi sum_c m_c n_c / L_c independent of the cell.
    PARITY UNPINNED (no working reference to compare against).
    """
    n = L1 * L2 * L3
    assert n == qbits.nqbits
    s_ij = np.asarray(s_ij, dtype=float)
    dims = np.array([L1, L2, L3], dtype=float)
    sites = np.array([(a, b, c) for a in range(L1) for b in range(L2) for c in range(L3)], dtype=float)
    phases = np.exp(2j * np.pi * (sites / dims) @ sites.T)  # [q, i] = exp(i q.R_i)
    s = np.einsum("qi,ij,qj->q", phases, s_ij, phases.conj()).real / n**2
    return s.reshape(L1, L2, L3).copy()


def structure_factor(positions, s_ij, qvecs) -> np.ndarray:
    """Position-general structure factor: S(q) = N^-2 sum_ij s_ij exp(i q.(x_i - x_j)) with
    the TRUE site positions x_i (any register: triangular, kagome, random) and an arbitrary list of wave vectors
    ``qvecs`` of shape (n_q, dim) -- the definition of docs/theory/spatial-correlation.md:32-34 and of the
    docstring of qse/magnetic.py:251-256, without the one-site-per-grid-point restriction of :279.  Host numpy."""
    pos = np.asarray(getattr(positions, "positions", positions), dtype=float)
    q = np.atleast_2d(np.asarray(qvecs, dtype=float))
    if q.shape[1] != pos.shape[1]:
        raise ValueError("qvecs must have the dimension of the positions")
    s_ij = np.asarray(s_ij, dtype=float)
    if s_ij.shape != (pos.shape[0], pos.shape[0]):
        raise ValueError("s_ij must be (nqbits, nqbits)")
    phases = np.exp(1j * q @ pos.T)  # [q, i]
    return np.einsum("qi,ij,qj->q", phases, s_ij, phases.conj()).real / pos.shape[0] ** 2


def reciprocal_grid(cell, shape) -> np.ndarray:
    """The wave vectors q = m . B / L (m_c = 0..L_c-1, B = 2 pi inv(A^T): qse/cell.py:58-68) the reference's
    structure factor samples (qse/magnetic.py:285), as an (L1*...*Ld, dim) list in row-major order of m."""
    a = np.asarray(getattr(cell, "lattice_vectors", cell), dtype=float)
    b = 2 * np.pi * np.linalg.inv(a.T)
    shape = tuple(int(v) for v in shape)
    m = np.array(np.meshgrid(*[np.arange(v) for v in shape], indexing="ij")).reshape(len(shape), -1).T
    return (m / np.array(shape, dtype=float)) @ b


def heatmap(qbits, scalar, **kwargs):
    """Hook-up to the reference's ``qse.vis.qbits_heatmap`` (qse/vis/scalar.py:6-78) for per-site results of the
    calculator -- ``heatmap(qbits, calc.get_number_operator())``, spins[:, 2], a row of s_ij ...  Needs the reference
    package and matplotlib; raises with a clear message otherwise (visualisation is out of this package's scope)."""
    scalar = np.asarray(scalar, dtype=float).reshape(-1)
    if scalar.size != qbits.nqbits:
        raise Exception("The length of scalar must equal the number of Qubits.")
    try:
        from qse.vis import qbits_heatmap  # type: ignore
    except Exception as exc:
        raise Exception("heatmap() forwards to qse.vis.qbits_heatmap: install ICHEC/qse with matplotlib") from exc
    return qbits_heatmap(qbits, scalar, **kwargs)
