"""Developer tool: H psi / sweep-step time under different engine options, one process (GPU box).

    python tools/ab_opts.py 25 diag_on_the_fly=1 diag_on_the_fly=0 sb_bits=17,sb_lag=8

Each argument after the size is one comma-separated option set (qsb_set_option keys); options not named in
a set are reset to their defaults (0 = auto for the plan options).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import qse_b200 as qb
from qse_b200 import lib

DEFAULTS = {"diag_on_the_fly": -1, "sb_bits": 0, "sb_lag": 0, "tile_bits": 12, "tensor_map": 1, "hpsi_path": 0}

n = int(sys.argv[1])
sets = [dict((kv.split("=")[0], float(kv.split("=")[1])) for kv in a.split(",")) for a in sys.argv[2:]] or [{}]
reg = qb.lattices.square(0.9 * qb.blockade_radius(2 * np.pi), 5, 5) if n == 25 else \
    qb.lattices.random_register(n, 0.9 * qb.blockade_radius(2 * np.pi), seed=n)
with lib.Engine(n) as eng:
    eng.set_interaction(qb.interaction_matrix(reg.positions))
    eng.reset_state()
    eng.evolve_schedule(np.full(2, 3.0), np.full(2, -4.0), np.full(2, 0.001))
    for rep in range(2):
        for opts in sets:
            for k, v in {**DEFAULTS, **opts}.items():
                eng.set_option(k, v)
            ms = eng.time_apply_h(3.0, -4.0, 12, False)[3:]
            steps = 6
            st = eng.time_schedule(np.linspace(2.0, 6.0, steps), np.linspace(-5, 10, steps), np.full(steps, 0.001))[2:]
            print(json.dumps({"n": n, "opts": opts, "hpsi_ms": round(float(np.median(ms)), 4),
                              "hpsi_min": round(float(np.min(ms)), 4), "step_ms": round(float(np.median(st)), 3),
                              "terms": eng.metrics()["last_terms"], "norm": eng.norm()}), flush=True)
