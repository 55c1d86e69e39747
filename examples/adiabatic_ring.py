"""The reference's canonical workload (docs/use-cases/adiabatic.ipynb) on the B200 calculator.

Ring of 8 qubits at the blockade radius of Omega = 1, adiabatic ramp Omega: 0 -> 2, delta: -2 -> 0
(the notebook's 5000 x 1 ns schedule), energies <H_i>, <H_f> recorded after every step, final spins,
<s_i s_j>, and the most probable bit strings.  Works with `qse` objects as well as with the
stand-ins of qse_b200 (same names): replace `qb.` by `qse.` for geometry and signals.

    python examples/adiabatic_ring.py [n_values]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qse_b200 as qb  # noqa: E402


def main(n_values: int = 5000):
    qbits = qb.lattices.ring(qb.blockade_radius(1.0), 8)
    amplitude = qb.Signal(np.linspace(0.0, 2.0, n_values), n_values)   # 1 ns per value
    detuning = qb.Signal(np.linspace(-2.0, 0.0, n_values), n_values)
    calc = qb.B200(qbits, amplitude, detuning)
    h_i, h_f = calc.get_hamiltonian(0.0, -2.0), calc.get_hamiltonian(2.0, 0.0)
    t0 = time.perf_counter()
    result = calc.calculate(e_ops=[h_i, h_f])
    dt = time.perf_counter() - t0
    energies = result["expectations"]
    print(f"{n_values} steps in {dt:.3f} s; <H_i>: {energies[0, 0]:+.6f} -> {energies[-1, 0]:+.6f}; "
          f"<H_f>: {energies[0, 1]:+.6f} -> {energies[-1, 1]:+.6f}")
    print("final <Z_i>:", np.round(calc.get_spins()[:, 2], 6))
    print("s_ij[0, :]:", np.round(calc.get_sij()[0], 6))
    for bits, p in calc.bitstring_probabilities(4):
        print(f"  |{bits}>  p = {p:.6f}")
    print("1000 shots:", calc.sample_final_state(1000, seed=1).most_common(3))
    calc.close()
    return result


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 5000)
