"""Developer tool: a handful of H psi applications for ncu (python tools/prof_hpsi.py N sb_bits path)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qse_b200 as qb  # noqa: E402
from qse_b200 import lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
sb = int(sys.argv[2]) if len(sys.argv) > 2 else 19
path = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reg = qb.lattices.random_register(n, 0.9 * qb.blockade_radius(2 * np.pi), seed=n)
with lib.Engine(n) as eng:
    eng.set_interaction(qb.interaction_matrix(reg.positions))
    eng.set_option("hpsi_path", path)
    eng.set_option("sb_bits", sb)
    eng.reset_state()
    eng.evolve_schedule(np.full(2, 3.0), np.full(2, -4.0), np.full(2, 0.001))
    ms = eng.time_apply_h(3.0, -4.0, 6, False)
    print("hpsi ms", ms)
