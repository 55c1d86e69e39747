"""Developer tool: BASELINE config 1 (square 3x3, 5000 x 1 ns ramp) wall-clock, both stepper paths."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qse_b200 as qb  # noqa: E402

om_max = 2 * np.pi
qbits = qb.lattices.square(0.9 * qb.blockade_radius(om_max), 3, 3)
T = 5000
amp = qb.Signal(np.linspace(0, om_max, T), T)
det = qb.Signal(np.linspace(-2 * om_max, 2 * om_max, T), T)
for small in (1, 0):
    calc = qb.B200(qbits, amp, det)
    calc._engine_ready().set_option("small_kernel", small)
    calc.calculate()
    t0 = time.perf_counter()
    out = calc.calculate(e_ops=[calc.get_hamiltonian(om_max, 2 * om_max)])
    t1 = time.perf_counter()
    print(f"small_kernel={small}: {T} steps in {t1 - t0:.3f} s ({T / (t1 - t0):.0f} steps/s), "
          f"hpsi {calc.metrics['hpsi_calls']}, launches {calc.metrics['kernel_launches']}, "
          f"final <H_f> {out['expectations'][-1, 0]:.9f}")
    calc.close()
