"""Developer tool: a handful of stand-alone H psi applications for ncu.

    python tools/prof_hpsi.py N [sb_bits] [hpsi_path]

Launch order (hpsi_ws_kernel only): a 1-step warm-up sweep, then 6 stand-alone H psi = 12 launches of
hpsi_ws_kernel (fused, plain) for N >= 22; `-k regex:hpsi_ws --launch-skip-before-match 0 -s <warm-up launches>`:
the script prints the number of hpsi_ws launches of the warm-up so that the capture can skip exactly those.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qse_b200 as qb  # noqa: E402
from qse_b200 import lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
sb = int(sys.argv[2]) if len(sys.argv) > 2 else 0
path = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reg = qb.lattices.square(0.9 * qb.blockade_radius(2 * np.pi), 5, 5) if n == 25 else \
    qb.lattices.random_register(n, 0.9 * qb.blockade_radius(2 * np.pi), seed=n)
with lib.Engine(n) as eng:
    eng.set_interaction(qb.interaction_matrix(reg.positions))
    eng.set_option("hpsi_path", path)
    eng.set_option("sb_bits", sb)
    eng.set_option("fixed_terms", 4)  # a known number of launches in the warm-up: 4 terms x (fused + plain)
    eng.reset_state()
    eng.evolve_schedule(np.full(1, 3.0), np.full(1, -4.0), np.full(1, 0.001))
    m = eng.metrics()
    print("warm-up hpsi_ws launches:", m["hpsi_calls"] * m["hpsi_passes"])
    ms = eng.time_apply_h(3.0, -4.0, 6, False)
    print("hpsi ms", ms)
