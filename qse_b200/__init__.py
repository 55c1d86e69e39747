"""qse_b200 — a B200-native state-vector Calculator for ICHEC/qse.

    import qse_b200 as qb
    qbits = qb.lattices.square(0.9 * qb.blockade_radius(2 * np.pi), 3, 3)   # or qse.lattices.square(...)
    calc = qb.B200(qbits, amplitude=qb.Signal(...), detuning=qb.Signal(...))
    out = calc.calculate()           # {"state": (2^N, 1) complex128}
    spins = calc.get_spins()         # (N, 3)

The compute path is hand-written sm_100a CUDA behind the C ABI of ``include/qsb.h``
(``qse_b200/libqsb.so``, called through ctypes).  There is no CPU fallback.
"""

from .host import C6, Cell, Operator, Operators, Qbits, Signal, Signals, blockade_radius, lattices  # noqa: F401
from .calculator import B200, Hamiltonian, evolve_operators, flatten_schedule, interaction_matrix, operator_terms  # noqa: F401
from . import magnetic, mis  # noqa: F401

__all__ = ["B200", "Hamiltonian", "Cell", "Qbits", "Signal", "Signals", "Operator", "Operators", "lattices",
           "blockade_radius", "C6", "magnetic", "mis", "interaction_matrix", "flatten_schedule", "operator_terms",
           "evolve_operators"]
__version__ = "0.1.0"
