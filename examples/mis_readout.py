"""BASELINE config 5 in small: a random-position register as a unit-disk maximum-independent-set instance.

Positions -> blockade graph (`qse_b200.mis.unit_disk_graph`), a detuning sweep at constant drive on the B200
calculator, the most probable bit strings read out on the device (`bitstring_probabilities`) and scored against
the graph: probability mass on independent sets, the largest independent set found, a greedy classical baseline.

    python examples/mis_readout.py [n_qubits] [n_values]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qse_b200 as qb  # noqa: E402
from qse_b200 import mis  # noqa: E402


def main(n: int = 16, n_values: int = 1500):
    omega = 2 * np.pi
    rb = qb.blockade_radius(omega)
    qbits = qb.lattices.random_register(n, 0.9 * rb, seed=20261019)
    adjacency = mis.unit_disk_graph(qbits, rb)
    amp = qb.Signal(np.concatenate([np.linspace(0.0, omega, n_values // 10), np.full(n_values - n_values // 10, omega)]), n_values)
    det = qb.Signal(np.linspace(-2 * omega, 2 * omega, n_values), n_values)  # 1 ns per value
    calc = qb.B200(qbits, amp, det, return_state=False)
    calc.calculate()
    top = calc.bitstring_probabilities(16)
    score = mis.score_readout(top, adjacency)
    print(f"{n} qubits, {int(adjacency.sum()) // 2} blockade edges; greedy independent set: {score['greedy_size']}")
    print(f"top-16 read-out: p(independent) = {score['p_independent']:.4f}, largest independent set "
          f"{score['largest_independent_set']} with p = {score['p_largest']:.4f}")
    for bits, p in top[:5]:
        tag = "maximal IS" if mis.is_maximal_independent_set(bits, adjacency) else \
            ("IS" if mis.is_independent_set(bits, adjacency) else f"{mis.violations(bits, adjacency)} violations")
        print(f"  |{bits}>  p = {p:.5f}  {tag}")
    counts = calc.sample_final_state(2000, seed=7)
    good = sum(c for b, c in counts.items() if mis.is_independent_set(b, adjacency))
    print(f"2000 shots: {good} land on independent sets")
    calc.close()
    return score


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:3]))
