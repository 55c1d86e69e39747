"""The reference's generic-Hamiltonian notebook (docs/use-cases/time_evo.ipynb) on the B200 engine.

A transverse-field Ising chain built with `Qbits.compute_interaction_hamiltonian` + `Operator` exactly as the
notebook does, evolved from |0...0> for t_final -- what `qutip.sesolve(hamiltonian.to_qutip(), psi0, [0, t_final])`
computes there.  The ZZ strings fold into the diagonal and the single-X strings into the per-site drive, so the
whole model runs in the tiled H psi kernels.  Works with `qse.Operators` as well (same attributes).

    python examples/tfim_time_evolution.py [n_qubits]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qse_b200 as qb  # noqa: E402


def main(n: int = 20, j: float = 1.0, h: float = 0.5, spacing: float = 1.0, t_final: float = 1.0):
    qbits = qb.lattices.chain(spacing, n)
    hamiltonian = qbits.compute_interaction_hamiltonian(lambda d: -j * np.isclose(d, spacing), "Z")
    for i in range(n):
        hamiltonian += qb.Operator("X", i, n, coef=-h)
    t0 = time.perf_counter()
    psi = qb.evolve_operators(hamiltonian, t_final)
    dt = time.perf_counter() - t0
    print(f"TFIM chain, {n} qubits, {hamiltonian.nterms} terms: exp(-i H {t_final}) |0...0> in {dt:.3f} s")
    print("norm:", float(np.vdot(psi, psi).real), " <psi(t)|0...0> =", complex(psi[0]))
    # magnetisation per site through the drop-in observables of qse.magnetic
    mz = qb.magnetic.get_spins(psi, n)[:, 2]
    print("<Z_i>:", np.round(mz, 6))
    return psi


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 20)
