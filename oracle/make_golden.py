"""Generate tests/golden/*.npz by EXECUTING the unmodified reference (TEST INFRASTRUCTURE).

Runs only where /root/reference exists (this container); the GPU box uses the committed
fixtures.  Usage:  python -m oracle.make_golden

Every array below is an OUTPUT of reference code (or a stored fixture of the reference's own
tests), with the inputs that produced it, so the tests can replay the inputs through the
oracle and through the CUDA path:

  lattices.npz      qse.lattices.* positions + cells            (qse/lattices/lattices.py)
  interaction.npz   Qbits.compute_interaction_hamiltonian pairs (qse/qbits.py:1239-1304)
  basis.npz         magnetic.get_basis(N) bit tables            (qse/magnetic.py:14-52)
  magnetic.npz      magnetic.get_spins / get_sisj / get_number_operator on seeded states
  exact_n1.npz      calc.ExactSimulator on the 9 (omega, delta) pairs of tests/qse/calc_test.py:32-53
  local_tests.npz   tests/local_tests/results_{myqlm,pulser}.npy (stored 4-qubit states+spins)
  signals.npz       Signal/Signals schedule arithmetic          (qse/signal/signal.py:253-263)
"""

from __future__ import annotations

import os

import numpy as np

from oracle.refshim import REFERENCE_ROOT, import_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

LATTICE_CASES = [
    ("chain", (4.0, 4)), ("chain", (1.7, 9)),
    ("square", (2.0, 3, 3)), ("square", (6.1, 2, 2)), ("square", (5.5, 5, 5)), ("square", (1.0, 2, 5)),
    ("triangular", (1.3, 3, 2)), ("triangular", (5.0, 4, 4)), ("triangular", (6.0, 11, 3)),
    ("hexagonal", (1.1, 2, 3)), ("kagome", (1.0, 2, 2)), ("kagome", (6.0, 5, 2)),
    ("ring", (3.0, 8)), ("ring", (1.0, 5)), ("torus", (4, 3, 5.0, 6.0)),
]


def main():
    qse = import_reference()
    os.makedirs(OUT, exist_ok=True)

    # ---- lattices -------------------------------------------------------------------
    lat = {}
    for idx, (name, args) in enumerate(LATTICE_CASES):
        q = getattr(qse.lattices, name)(*args)
        lat[f"{idx}_name"] = np.array(name)
        lat[f"{idx}_args"] = np.array(args, dtype=float)
        lat[f"{idx}_positions"] = q.positions.copy()
        if q.cell is not None:
            lat[f"{idx}_cell"] = q.cell.lattice_vectors.copy()
            lat[f"{idx}_reciprocal"] = q.cell.reciprocal()
    lat["n_cases"] = np.array(len(LATTICE_CASES))
    np.savez(os.path.join(OUT, "lattices.npz"), **lat)

    # ---- interaction terms ----------------------------------------------------------
    c6 = 5420158.53
    inter = {}
    cases = [("square", (0.9 * qse.calc.blockade_radius(2 * np.pi), 3, 3)),
             ("chain", (4.0, 4)),
             ("kagome", (6.3, 2, 2)),
             ("chain", (60.0, 6))]  # long chain: far pairs fall under the 1e-8 cut-off
    for idx, (name, args) in enumerate(cases):
        q = getattr(qse.lattices, name)(*args)
        ops = q.compute_interaction_hamiltonian(lambda d: c6 / d**6, "N")
        inter[f"{idx}_positions"] = q.positions.copy()
        inter[f"{idx}_pairs"] = np.array([op.indices for op in ops.operator_list], dtype=np.int64).reshape(-1, 2)
        inter[f"{idx}_coefs"] = np.array([op.coef for op in ops.operator_list], dtype=float)
    inter["n_cases"] = np.array(len(cases))
    inter["blockade_radius_2pi"] = np.array(qse.calc.blockade_radius(2 * np.pi))
    np.savez(os.path.join(OUT, "interaction.npz"), **inter)

    # ---- basis tables ---------------------------------------------------------------
    np.savez(os.path.join(OUT, "basis.npz"),
             **{f"n{n}": qse.magnetic.get_basis(n) for n in range(1, 8)})

    # ---- magnetic observables on seeded random states ---------------------------------
    rng = np.random.default_rng(20261019)
    mag = {}
    sizes = [1, 2, 3, 4, 5, 6, 7]
    for n in sizes:
        psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        psi /= np.linalg.norm(psi)
        mag[f"n{n}_state"] = psi
        mag[f"n{n}_spins"] = qse.magnetic.get_spins(psi, n)
        mag[f"n{n}_sisj"] = qse.magnetic.get_sisj(psi, n)
        mag[f"n{n}_number"] = qse.magnetic.get_number_operator(psi, n)
    mag["sizes"] = np.array(sizes)
    # the basis-state cases of tests/qse/magnetic_test.py:46-53
    for bits in ("1", "00", "001", "101011"):
        n = len(bits)
        psi = np.zeros(1 << n, dtype=complex)
        psi[int(bits, 2)] = 1.0
        mag[f"basis_{bits}_number"] = qse.magnetic.get_number_operator(psi, n)
    np.savez(os.path.join(OUT, "magnetic.npz"), **mag)

    # ---- N = 1 closed form ----------------------------------------------------------
    ex = {"omega": [], "delta": [], "state": [], "closed_form": []}
    duration = 400
    for omega in (10.01, 0.453, 5.6):
        for delta in (0.12, 5.242, -13.0):
            calc = qse.calc.ExactSimulator(amplitude=qse.Signal([omega], duration),
                                           detuning=qse.Signal([delta], duration))
            calc.calculate()
            mag_ = np.sqrt(delta**2 + omega**2)
            ang = mag_ * (duration / 1000) * 0.5
            closed = np.cos(ang) * np.array([1, 0]) - 1j * np.sin(ang) * np.array([delta, omega]) / mag_
            ex["omega"].append(omega)
            ex["delta"].append(delta)
            ex["state"].append(calc.statevector.reshape(-1))
            ex["closed_form"].append(closed)
    np.savez(os.path.join(OUT, "exact_n1.npz"), duration=np.array(duration),
             **{k: np.array(v) for k, v in ex.items()})

    # ---- stored states of the reference's local tests ---------------------------------
    loc = {}
    for name in ("myqlm", "pulser"):
        r = np.load(os.path.join(REFERENCE_ROOT, "tests", "local_tests", f"results_{name}.npy"),
                    allow_pickle=True).item()
        loc[f"{name}_state"] = np.asarray(r["state"]).reshape(-1)
        loc[f"{name}_spins"] = np.asarray(r["spins"])
    loc["positions"] = qse.lattices.chain(4.0, 4).positions
    loc["omega"] = np.array(10.01)
    loc["delta"] = np.array(0.12)
    loc["duration"] = np.array(400)
    np.savez(os.path.join(OUT, "local_tests.npz"), **loc)

    # ---- schedule arithmetic --------------------------------------------------------
    s1 = qse.Signal(np.linspace(0, 1, 5), 20)
    s2 = qse.Signal([3.0], 400)
    ss = qse.Signals([s1, s2])
    np.savez(os.path.join(OUT, "signals.npz"),
             s1_tpv=np.array(s1.time_per_value()), s2_tpv=np.array(s2.time_per_value()),
             s1_expand=s1.expand(), total=np.array(ss.duration), n=np.array(len(ss)),
             expand=ss.expand())
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f"  {f}: {os.path.getsize(os.path.join(OUT, f))} bytes")


if __name__ == "__main__":
    main()
