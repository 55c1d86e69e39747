"""Host-side mirror of the slice of the reference interface the calculator consumes.

The calculator of :mod:`qse_b200.calculator` attaches to the *reference's own* objects
(``qse.Qbits``, ``qse.Signal``/``qse.Signals``, anything ``qse.lattices.*`` returns) by duck
typing: it only reads ``qbits.positions`` / ``qbits.nqbits`` and, per signal segment,
``.values`` / ``.duration`` / ``time_per_value()``.  On a machine without the reference
package (the GPU boxes: no ``qse``, no QuTiP, no matplotlib) the same workloads must still
be constructible, so this module provides small stand-ins with the same names, argument
meaning and error behaviour — geometry containers and pulse containers only, none of the
reference's geometry-editing, visualisation or converter methods.

Site ordering, cell conventions and schedule arithmetic are pinned against the reference's
generators by ``tests/test_host_mirror.py`` (golden positions in ``tests/golden``).

Reference interfaces mirrored (file:line in ICHEC/qse 1.1.21):
  Cell                      qse/cell.py:4-77
  Qbits (hot subset)        qse/qbits.py:144, 175-205, 487-534, 575-584, 992-1042
  Signal                    qse/signal/signal.py:58-60, 239-274
  Signals                   qse/signal/signals.py:39-77
  lattices.*                qse/lattices/lattices.py:15-372
  blockade_radius           qse/calc/blockade_radius.py:1-30
"""

from __future__ import annotations

import itertools
import math
from types import SimpleNamespace

import numpy as np

__all__ = ["Cell", "Qbits", "Signal", "Signals", "lattices", "blockade_radius", "C6"]

C6 = 5420158.53  # (rad/us)(um)^6, qse/calc/qutip.py:7


def blockade_radius(rabi_frequency: float, c6: float = C6) -> float:
    """r_b = (c6 / Omega)^(1/6)  [um]  (qse/calc/blockade_radius.py:30)."""
    return (c6 / rabi_frequency) ** (1 / 6)


class Cell:
    """Lattice cell; rows of ``lattice_vectors`` are the lattice vectors (qse/cell.py)."""

    def __init__(self, lattice_vectors):
        vecs = np.array(lattice_vectors, dtype=float)
        if vecs.shape not in ((1, 1), (2, 2), (3, 3)):
            raise Exception("The lattice vectors must be a 1x1, 2x2 or 3x3 matrix.")
        self.lattice_vectors = vecs

    @property
    def dim(self) -> int:
        return self.lattice_vectors.shape[0]

    def volume(self) -> float:
        return abs(float(np.linalg.det(self.lattice_vectors)))

    def rank(self) -> int:
        return int(np.linalg.matrix_rank(self.lattice_vectors))

    def reciprocal(self) -> np.ndarray:
        """2 pi inv(A^T)  (qse/cell.py:58-68)."""
        return 2 * np.pi * np.linalg.inv(self.lattice_vectors.T)

    def __repr__(self):
        return "\n".join(" ".join(f"{x:10.5f}" for x in row) for row in self.lattice_vectors)


class Qbits:
    """Positions (+ optional cell) of a qubit register: the part a calculator reads."""

    def __init__(self, positions=None, cell=None, labels=None, calculator=None):
        if positions is None:
            raise ValueError("positions are required")
        pos = np.array(positions, dtype=float)
        if pos.ndim != 2 or pos.shape[1] not in (1, 2, 3):
            raise ValueError("positions must have shape (nqbits, dim) with dim in 1..3")
        self.positions = pos
        self.cell = None if cell is None else (cell if isinstance(cell, Cell) else Cell(cell))
        self.labels = [str(i) for i in range(len(pos))] if labels is None else list(labels)
        self.shape = (1,) * self.dim
        self._calc = None
        self.calc = calculator

    # -- the attach hook of qse/qbits.py:164-173 -------------------------------------
    @property
    def calc(self):
        return self._calc

    @calc.setter
    def calc(self, calc):
        self._calc = calc
        if hasattr(calc, "set_qbits"):
            calc.set_qbits(self)

    @property
    def nqbits(self) -> int:
        return self.positions.shape[0]

    @property
    def dim(self) -> int:
        return self.positions.shape[1]

    def __len__(self) -> int:
        return self.nqbits

    def get_distance(self, i: int, j: int) -> float:
        return float(np.linalg.norm(self.positions[i] - self.positions[j]))

    def get_all_distances(self) -> np.ndarray:
        n = self.nqbits
        d = np.zeros((n, n))
        for i, j in itertools.combinations(range(n), 2):
            d[i, j] = d[j, i] = self.get_distance(i, j)
        return d

    def compute_interaction_hamiltonian(self, distance_func, interaction, tol=1e-8):
        """``Operators`` with one term ``distance_func(|R_i - R_j|) * interaction_i interaction_j`` per pair i < j,
        kept iff its magnitude exceeds ``tol`` (qse/qbits.py:1239-1304: the TFIM / XY builders of
        docs/use-cases/time_evo.ipynb and ssh_model.ipynb)."""
        terms = []
        for i, j in itertools.combinations(range(self.nqbits), 2):
            coef = distance_func(self.get_distance(i, j))
            if np.abs(coef) > tol:
                terms.append(Operator(interaction, (i, j), self.nqbits, coef))
        return Operators(terms)

    def repeat(self, reps):
        """Tile the register over its cell; site index = (m0*r1 + m1)*nbasis + b."""
        if self.cell is None:
            raise ValueError("Cannot repeat without a cell")
        if isinstance(reps, (int, np.integer)):
            reps = (int(reps),) * self.dim
        reps = tuple(int(r) for r in reps)
        if len(reps) != self.dim:
            raise ValueError("one repeat count per lattice vector is required")
        vecs = self.cell.lattice_vectors
        blocks = [self.positions + np.dot(m, vecs) for m in itertools.product(*(range(r) for r in reps))]
        out = Qbits(np.concatenate(blocks), cell=np.array([r * v for r, v in zip(reps, vecs)]))
        out.shape = tuple(int(s * r) for s, r in zip(self.shape, reps))
        return out


class Operator:
    """One Pauli string ``coef * prod op_q`` over "X", "Y", "Z", "N" (N = diag(0, 1)): the duck-typed subset of
    ``qse.Operator`` (qse/operator/operator.py:4-60) a calculator reads -- ``operator``, ``indices``, ``nqbits``,
    ``coef``."""

    def __init__(self, operator, indices, nqbits, coef=1.0):
        indices = [int(indices)] if isinstance(indices, (int, np.integer)) else [int(i) for i in indices]
        operator = [operator] * len(indices) if isinstance(operator, str) else list(operator)
        if any(op not in ("X", "Y", "Z", "N") for op in operator):
            raise Exception("An operator must be one of 'X', 'Y', 'Z' or 'N'.")
        if any(i < 0 or i >= nqbits for i in indices):
            raise Exception("Qubit indices must be between 0 and nqbits-1, inclusive.")
        if len(indices) != len(operator):
            raise Exception("The number of passed qubits must equal the number of passed operators.")
        self.operator, self.indices, self.nqbits, self.coef = operator, indices, int(nqbits), coef

    def copy(self):
        return Operator(self.operator, self.indices, self.nqbits, self.coef)

    def __mul__(self, k):
        if not isinstance(k, (int, float)):
            raise NotImplementedError("Multiplication only supported for scalars.")
        return Operator(self.operator, self.indices, self.nqbits, self.coef * k)

    __rmul__ = __mul__

    def __repr__(self):
        return f"{self.coef:.2f} " + " ".join(f"{op}{q}" for op, q in zip(self.operator, self.indices))


class Operators:
    """A sum of :class:`Operator` terms (``qse.Operators``, qse/operator/operators.py:4-30, 80-120)."""

    def __init__(self, operator_list=None):
        self.operator_list = list(operator_list) if operator_list is not None else []
        if self.operator_list and not all(op.nqbits == self.operator_list[0].nqbits for op in self.operator_list):
            raise Exception("All operators in operator_list must act on the same number of qubits.")

    @property
    def nterms(self) -> int:
        return len(self.operator_list)

    @property
    def nqbits(self):
        return self.operator_list[0].nqbits if self.operator_list else None

    def extend(self, other):
        if self.nterms and other.nqbits != self.nqbits:
            raise Exception("other must have the same number of qubits.")
        if hasattr(other, "operator_list"):
            self.operator_list += other.operator_list
        elif hasattr(other, "operator") and hasattr(other, "indices"):
            self.operator_list += [other]
        else:
            raise Exception("other must be Operators or Operator.")

    def copy(self):
        return Operators(self.operator_list.copy())

    def __add__(self, other):
        out = self.copy()
        out.extend(other)
        return out

    def __iadd__(self, other):
        self.extend(other)
        return self

    def __mul__(self, k):
        if not isinstance(k, (int, float)):
            raise NotImplementedError("Multiplication only supported for scalars.")
        return Operators([op * k for op in self.operator_list])

    __rmul__ = __mul__

    def __getitem__(self, i):
        return self.operator_list[i]

    def __repr__(self):
        return f"Number of terms: {self.nterms}" + "".join("\n" + repr(op) for op in self.operator_list)


class Signal:
    """Piecewise-constant pulse: ``values`` equally spaced over ``duration`` (ns)."""

    def __init__(self, values, duration=None):
        self.values = np.asarray(values, dtype=float)
        self.duration = len(self.values) if duration is None else int(duration)

    @property
    def duration(self) -> int:
        return self._duration

    @duration.setter
    def duration(self, new_duration):
        if not isinstance(new_duration, int):
            raise ValueError("The duration must be an ints")
        if new_duration % len(self) != 0:
            raise ValueError("The number of values must divide the duration.")
        self._duration = new_duration

    def time_per_value(self) -> int:
        return self.duration // len(self)

    def expand(self) -> np.ndarray:
        return np.repeat(self.values, self.time_per_value())

    def __len__(self):
        return len(self.values)

    def __iter__(self):
        return iter(self.values)

    def __getitem__(self, i):
        return self.values[i]

    def __eq__(self, other):
        return self.duration == other.duration and bool((self.values == other.values).all())

    # scalar arithmetic only, as in the reference (qse/signal/signal.py:104-225)
    def __mul__(self, k):
        if not isinstance(k, (float, int)):
            raise TypeError(f"Unsupported operand type for *: {type(k)}")
        return Signal(self.values * k, self.duration)

    __rmul__ = __mul__

    def __add__(self, k):
        if not isinstance(k, (float, int)):
            raise TypeError(f"Unsupported operand type for +: {type(k)}")
        return Signal(self.values + k, self.duration)

    __radd__ = __add__

    def __repr__(self):
        return f"Signal(duration={self.duration}, values={self.values})"


class Signals:
    """An ordered collection of :class:`Signal` segments."""

    def __init__(self, signals=None):
        if signals is None:
            signals = []
        elif not isinstance(signals, list) or not _is_signal(signals[0]):
            raise Exception("signals must be a list of Signal objects.")
        self._signals = signals

    @property
    def signals(self):
        return self._signals

    @property
    def duration(self) -> int:
        return sum(s.duration for s in self._signals)

    def __len__(self):
        return len(self._signals)

    def __iter__(self):
        return iter(self._signals)

    def __getitem__(self, i):
        return self._signals[i]

    def __add__(self, other):
        if not _is_signal(other):
            raise TypeError(f"Unsupported operand type for +: {type(other)}")
        return Signals(self._signals + [other])

    def __iadd__(self, other):
        if not _is_signal(other):
            raise TypeError(f"Unsupported operand type for +=: {type(other)}")
        self._signals = self._signals + [other]
        return self

    def expand(self) -> np.ndarray:
        return np.concatenate([s.expand() for s in self._signals])

    def __repr__(self):
        return f"Total duration={self.duration}\n" + "\n".join(f"  {s!r}" for s in self._signals)


def _is_signal(obj) -> bool:
    """Duck-typed: the reference's qse.Signal qualifies as well as ours."""
    return hasattr(obj, "values") and hasattr(obj, "duration") and hasattr(obj, "time_per_value")


def _is_signals(obj) -> bool:
    return hasattr(obj, "signals") and hasattr(obj, "duration") and not hasattr(obj, "values")


# --------------------------------------------------------------------------------------
# Lattice generators (qse/lattices/lattices.py) — Bravais cell + basis, tiled by Qbits.repeat
# --------------------------------------------------------------------------------------
def _need_repeats(value, name):
    if not isinstance(value, int) or value < 2:
        raise Exception(f"The {name} must be an integer greater than 1. {name}={value}")


def _bravais(cell, basis, reps):
    return Qbits(np.array(basis, dtype=float), cell=np.array(cell, dtype=float)).repeat(reps)


def _chain(lattice_spacing: float, repeats: int) -> Qbits:
    _need_repeats(repeats, "repeats")
    return _bravais([[lattice_spacing]], [[0.0]], (repeats,))


def _square(lattice_spacing: float, repeats_x: int, repeats_y: int) -> Qbits:
    _need_repeats(repeats_x, "repeats_x")
    _need_repeats(repeats_y, "repeats_y")
    a = lattice_spacing
    return _bravais([[a, 0], [0, a]], [[0.0, 0.0]], (repeats_x, repeats_y))


def _triangular(lattice_spacing: float, repeats_x: int, repeats_y: int) -> Qbits:
    _need_repeats(repeats_x, "repeats_x")
    _need_repeats(repeats_y, "repeats_y")
    cell = np.array([[1, 0], [0.5, np.sqrt(3) / 2]]) * lattice_spacing
    return _bravais(cell, [[0.0, 0.0]], (repeats_x, repeats_y))


def _hexagonal(lattice_spacing: float, repeats_x: int, repeats_y: int) -> Qbits:
    _need_repeats(repeats_x, "repeats_x")
    _need_repeats(repeats_y, "repeats_y")
    cell = np.array([[3 / 2, np.sqrt(3) / 2], [3 / 2, -np.sqrt(3) / 2]]) * lattice_spacing
    return _bravais(cell, [[0, 0], [lattice_spacing, 0]], (repeats_x, repeats_y))


def _kagome(lattice_spacing: float, repeats_x: int, repeats_y: int) -> Qbits:
    _need_repeats(repeats_x, "repeats_x")
    _need_repeats(repeats_y, "repeats_y")
    a = lattice_spacing
    cell = np.array([[2, 0], [1, np.sqrt(3)]]) * a
    basis = [[0, 0], [a, 0], [a / 2, np.sqrt(3) * a / 2]]
    return _bravais(cell, basis, (repeats_x, repeats_y))


def _ring(spacing: float, nqbits: int) -> Qbits:
    radius = spacing * np.sqrt(0.5 / (1.0 - np.cos(2 * np.pi / nqbits)))
    theta = np.arange(nqbits, dtype=float)
    theta *= 2.0 * np.pi / nqbits
    return Qbits(radius * np.array([np.cos(theta), np.sin(theta)]).T)


def _torus(n_outer: int, n_inner: int, inner_radius: float, outer_radius: float) -> Qbits:
    theta = (2.0 * np.pi / n_outer) * np.arange(n_outer)
    phi = (2.0 * np.pi / n_inner) * np.arange(n_inner)
    rho = outer_radius + inner_radius * np.cos(theta)
    xyz = np.array([np.outer(rho, np.cos(phi)), np.outer(rho, np.sin(phi)),
                    np.outer(inner_radius * np.sin(theta), np.ones(phi.shape))])
    return Qbits(xyz.reshape(3, n_outer * n_inner).T)


def _random_register(nqbits: int, spacing: float, seed: int = 20261019, min_frac: float = 0.6) -> Qbits:
    """Synthetic unit-disk register (BASELINE config 5; SURVEY.md §8d 'C5').

    ``nqbits`` points uniform in a square of side sqrt(nqbits)*spacing, rejecting any point
    closer than ``min_frac*spacing`` to an accepted one.  Not a reference function.
    """
    rng = np.random.default_rng(seed)
    side = math.sqrt(nqbits) * spacing
    pts: list[np.ndarray] = []
    while len(pts) < nqbits:
        p = rng.uniform(0.0, side, size=2)
        if all(np.linalg.norm(p - q) >= min_frac * spacing for q in pts):
            pts.append(p)
    return Qbits(np.array(pts))


lattices = SimpleNamespace(chain=_chain, square=_square, triangular=_triangular,
                           hexagonal=_hexagonal, kagome=_kagome, ring=_ring, torus=_torus,
                           random_register=_random_register)
