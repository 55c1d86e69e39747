"""Developer tool: time H psi and sweep steps at a few sizes on one GPU (not the bench contract)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import qse_b200 as qb
from qse_b200 import lib

PEAK = 6563.2
sizes = [int(a) for a in sys.argv[1:]] or [20, 22, 25]
for n in sizes:
    reg = qb.lattices.random_register(n, 0.9 * qb.blockade_radius(2 * np.pi), seed=n)
    t0 = time.time()
    v = qb.interaction_matrix(reg.positions)
    with lib.Engine(n) as eng:
        eng.set_interaction(v)
        t_diag = time.time() - t0
        eng.reset_state()
        eng.evolve_schedule(np.full(3, 3.0), np.full(3, -4.0), np.full(3, 0.001))
        for path, sb in ((2, 20), (0, 16), (0, 18), (0, 19), (0, 20), (0, 21)):
            if sb > n:
                continue
            eng.set_option("hpsi_path", path)
            eng.set_option("sb_bits", sb)
            ms = eng.time_apply_h(3.0, -4.0, 12, False)[2:]
            gbs = 40.0 * (1 << n) / (np.median(ms) * 1e-3) / 1e9
            print(json.dumps({"n": n, "path": path, "sb": sb, "passes": eng.metrics()["hpsi_passes"], "hpsi_ms": round(float(np.median(ms)), 4),
                              "hpsi_gbs_40B": round(gbs, 1), "frac_measured_peak": round(gbs / PEAK, 4)}))
        eng.set_option("hpsi_path", 0)
        eng.set_option("sb_bits", 0)  # auto
        steps = 10
        om = np.linspace(1.0, 6.0, steps); de = np.linspace(-10, 10, steps); dt = np.full(steps, 0.001)
        ms = eng.time_schedule(om, de, dt)
        m = eng.metrics()
        print(json.dumps({"n": n, "step_ms": float(np.median(ms)), "terms": m["last_terms"], "steps_per_s": 1e3 / float(np.median(ms)),
                          "diag_build_s": t_diag, "norm": eng.norm()}))
