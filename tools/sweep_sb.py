"""Developer tool: H psi time against super-block size / lead at one size, one process (GPU box).

    python tools/sweep_sb.py 25 16:8 16:16 17:6 18:3      # sb_bits:sb_lag pairs
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import qse_b200 as qb
from qse_b200 import lib

n = int(sys.argv[1])
pairs = [tuple(int(v) for v in a.split(":")) for a in sys.argv[2:]] or [(18, 3)]
reg = qb.lattices.square(0.9 * qb.blockade_radius(2 * np.pi), 5, 5) if n == 25 else \
    qb.lattices.random_register(n, 0.9 * qb.blockade_radius(2 * np.pi), seed=n)
with lib.Engine(n) as eng:
    eng.set_interaction(qb.interaction_matrix(reg.positions))
    eng.reset_state()
    eng.evolve_schedule(np.full(2, 3.0), np.full(2, -4.0), np.full(2, 0.001))
    for rep in range(2):
        for sb, lag in pairs:
            eng.set_option("sb_bits", sb)
            eng.set_option("sb_lag", lag)
            ms = eng.time_apply_h(3.0, -4.0, 12, False)[3:]
            steps = 6
            st = eng.time_schedule(np.linspace(2.0, 6.0, steps), np.linspace(-5, 10, steps), np.full(steps, 0.001))[2:]
            print(json.dumps({"n": n, "sb": sb, "lag": lag, "hpsi_ms": round(float(np.median(ms)), 4),
                              "hpsi_min": round(float(np.min(ms)), 4), "step_ms": round(float(np.median(st)), 3),
                              "terms": eng.metrics()["last_terms"], "norm": eng.norm()}), flush=True)
