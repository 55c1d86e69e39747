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
