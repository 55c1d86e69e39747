"""Unit-disk maximum-independent-set helpers for the MIS read-out of BASELINE config 5.

The reference has no MIS encoder: its nearest code is the QUBO <-> distance mapping of qse/qubo.py:131-158
(``Q2rij`` turns couplings into target distances, ``cost_function`` scores a layout against them).  A Rydberg
register encodes an MIS instance geometrically instead: two atoms closer than the blockade radius cannot both be
excited, so the register IS the unit-disk graph and the detuning sweep of config 5 prepares large independent
sets.  These helpers go from positions to that graph and check / score the bit strings ``B200`` reads out
(``bitstring_probabilities``, ``sample_final_state``).  Host-side numpy only; no device work.
"""

from __future__ import annotations

import numpy as np


def unit_disk_graph(positions, radius: float) -> np.ndarray:
    """Boolean adjacency (N, N) of the unit-disk graph: i ~ j iff 0 < |R_i - R_j| <= radius
    (``radius`` = the blockade radius, qse.calc.blockade_radius)."""
    pos = np.asarray(getattr(positions, "positions", positions), dtype=float)
    d = np.linalg.norm(pos[:, None, :] - pos[None, :, :], axis=-1)
    adj = d <= radius
    np.fill_diagonal(adj, False)
    return adj


def bits_of(bitstring) -> np.ndarray:
    """'0110' / iterable of 0-1 / basis index is NOT accepted here: qubit 0 first, as ``np.binary_repr(k, N)``
    prints a basis index (docs/use-cases/adiabatic.ipynb cell 8)."""
    if isinstance(bitstring, str):
        return np.array([c == "1" for c in bitstring], dtype=bool)
    return np.asarray(bitstring).astype(bool)


def index_to_bits(k: int, nqbits: int) -> np.ndarray:
    """Basis index -> occupation bits, qubit q <-> bit N-1-q (qse/magnetic.py:51)."""
    return np.array([(int(k) >> (nqbits - 1 - q)) & 1 for q in range(nqbits)], dtype=bool)


def is_independent_set(bitstring, adjacency) -> bool:
    """No two excited qubits are adjacent."""
    b = bits_of(bitstring)
    adj = np.asarray(adjacency, dtype=bool)
    if b.size != adj.shape[0]:
        raise ValueError("bit string length must equal the number of qubits")
    return not bool(np.any(adj[np.ix_(b, b)]))


def is_maximal_independent_set(bitstring, adjacency) -> bool:
    """Independent, and every unexcited qubit has an excited neighbour (nothing can be added)."""
    b = bits_of(bitstring)
    adj = np.asarray(adjacency, dtype=bool)
    if not is_independent_set(b, adj):
        return False
    return bool(np.all(b | np.any(adj[:, b], axis=1)))


def violations(bitstring, adjacency) -> int:
    """Number of adjacent excited pairs (0 for an independent set): the blockade-violation count."""
    b = bits_of(bitstring)
    adj = np.asarray(adjacency, dtype=bool)
    return int(np.count_nonzero(np.triu(adj[np.ix_(b, b)], 1)))


def greedy_independent_set(adjacency, order=None) -> np.ndarray:
    """A maximal independent set by the minimum-degree greedy rule (a classical baseline to score a read-out
    against); returns the occupation bits."""
    adj = np.asarray(adjacency, dtype=bool).copy()
    n = adj.shape[0]
    alive = np.ones(n, dtype=bool)
    chosen = np.zeros(n, dtype=bool)
    while alive.any():
        deg = np.where(alive, (adj & alive[None, :]).sum(axis=1), n + 1)
        if order is not None:
            cand = [q for q in order if alive[q]]
            v = cand[0]
        else:
            v = int(np.argmin(deg))
        chosen[v] = True
        alive[v] = False
        alive[adj[v]] = False
    return chosen


def score_readout(bitstrings_with_probability, adjacency) -> dict:
    """Summary of a read-out ``[(bitstring, probability), ...]`` (``B200.bitstring_probabilities``): probability
    mass on independent sets, the largest independent set seen and its probability."""
    adj = np.asarray(adjacency, dtype=bool)
    p_ind, best, p_best = 0.0, 0, 0.0
    for bits, p in bitstrings_with_probability:
        if is_independent_set(bits, adj):
            p_ind += p
            size = int(bits_of(bits).sum())
            if size > best:
                best, p_best = size, p
            elif size == best:
                p_best += p
    return {"p_independent": p_ind, "largest_independent_set": best, "p_largest": p_best,
            "greedy_size": int(greedy_independent_set(adj).sum())}
