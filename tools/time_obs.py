"""Developer tool: wall-clock of the observable read-outs at a given size."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qse_b200 as qb  # noqa: E402
from qse_b200 import lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
reg = qb.lattices.random_register(n, 0.9 * qb.blockade_radius(2 * np.pi), seed=n)
with lib.Engine(n) as eng:
    eng.set_interaction(qb.interaction_matrix(reg.positions))
    eng.evolve_schedule(np.full(3, 5.0), np.full(3, 1.0), np.full(3, 0.002))
    for name, fn in (("spins", eng.spins), ("number", eng.number), ("sisj", eng.sisj), ("topk16", lambda: eng.topk(16)),
                     ("sample1e5", lambda: eng.sample(100000, 1)), ("energy", lambda: eng.energy(5.0, 1.0))):
        fn()
        t0 = time.perf_counter()
        fn()
        print(f"n={n} {name}: {1e3 * (time.perf_counter() - t0):.2f} ms")
