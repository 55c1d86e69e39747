#!/usr/bin/env python
"""bench.py — the measurement contract for the Rydberg sweep hot path.

    python bench.py --gpus 1 --steps 20 --warmup 3            # this repo's arm (default)
    python bench.py --impl reference --steps 20 --warmup 3    # the CPU arm (oracle port, host cores)
    torchrun --nproc-per-node P bench.py --gpus P ...          # one rank per GPU, sharded state

A "step" is one sweep step: psi <- exp(-i H(omega_s, delta_s) dt) psi (one signal value of
Qutip.calculate's loop, qse/calc/qutip.py:42-45).  Workload (BASELINE config 2):
square lattice 5x5 (25 qubits, 2^25 amplitudes per GPU) at 0.9 r_b(Omega_max = 2 pi), adiabatic
ramp Omega: 0 -> Omega_max, delta: -2 Omega_max -> +2 Omega_max, 1 ns per value.  With P GPUs the
per-GPU shard stays 2^25 amplitudes (weak scaling): 25 + log2(P) qubits on a 5x5 / 2x13 / 3x9 /
4x7 rectangle of the same square lattice.

`value` = H psi HBM GB/s in the 40 B/amplitude currency of SURVEY.md §8(d) sustained over the
timed sweep steps: 40 B * 2^N * (number of H psi applied) / time, whole job.  `steps_per_s` is the
other half of BASELINE.json's metric.  One JSON line on stdout (rank 0).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

OMEGA_MAX = 2 * np.pi
BYTES_PER_AMP = 40.0  # 16 (read psi) + 8 (read E_int) + 16 (write y): SURVEY.md §8(d)
SHAPES = {25: (5, 5), 26: (2, 13), 27: (3, 9), 28: (4, 7), 20: (4, 5), 22: (2, 11), 24: (4, 6), 16: (4, 4),
          12: (3, 4), 9: (3, 3), 30: (5, 6)}


def workload(n_qubits: int, n_values: int):
    import qse_b200 as qb

    a = 0.9 * qb.blockade_radius(OMEGA_MAX)
    qbits = qb.lattices.square(a, *SHAPES[n_qubits])
    amp = qb.Signal(np.linspace(0.0, OMEGA_MAX, n_values), n_values)  # 1 ns per value
    det = qb.Signal(np.linspace(-2 * OMEGA_MAX, 2 * OMEGA_MAX, n_values), n_values)
    return qbits, amp, det


def make_config(n: int, world: int) -> dict:
    """The workload description both arms print (identical keys and values: the driver compares them)."""
    p = world.bit_length() - 1
    return {"workload": f"square {SHAPES[n][0]}x{SHAPES[n][1]} ({n} qubits) adiabatic sweep, complex128, 1 ns steps",
            "qubits": n, "amplitudes_per_gpu": 1 << (n - p),
            "l2": "inputs larger than L2 (state 512 MiB per 2^25-amplitude shard vs 126 MB L2); no flush",
            "parallelism": f"state sharded by top {p} qubit(s) over {world} GPU(s) (the CPU arm runs the whole register on the host)",
            "stepper": "taylor tol 1e-14"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, every 5 ms).

    Same quantities as the nvidia-smi line of B200_PROFILING.md (clocks.sm, clocks.max.sm,
    clocks_event_reasons.*), polled in-process so that sub-second regions still get samples.
    """

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, device: int):
        self.device = device
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = None
        self._thread = None
        self._err = None

    def _run(self, nv, handle):
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(handle) / 1000.0)
            except Exception as exc:  # keep timing even if NVML hiccups
                self._err = str(exc)
                return
            time.sleep(0.005)

    def start(self):
        import threading

        try:
            import pynvml as nv

            nv.nvmlInit()
            idx = int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[self.device]) \
                if os.environ.get("CUDA_VISIBLE_DEVICES") else self.device
            handle = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM)
        except Exception as exc:
            self._err = str(exc)
            return
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, args=(nv, handle), daemon=True)
        self._thread.start()

    def stop(self) -> dict:
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "samples": len(self.samples), "reasons": sorted(self.reasons),
               "power_w_max": max(self.power) if self.power else None}
        if self._err:
            out["error"] = self._err
        return out


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle's C port on the host cores (the reference itself is Python + QuTiP, absent)
# --------------------------------------------------------------------------------------------
def cpu_port_run(n_qubits: int, steps: int, warmup: int, budget_s: float):
    """Time the C restatement on the same workload.  Returns (gbs, steps_per_s, info)."""
    from oracle import c_port
    from oracle import rydberg_oracle as orc

    qbits, amp, det = workload(n_qubits, warmup + steps)
    v = orc.interaction_matrix(qbits.positions)
    e = c_port.build_diagonal(v, n_qubits)
    om, de, dt = orc.flatten_schedule([(amp.values, amp.duration)], [(det.values, det.duration)])
    # torchrun exports OMP_NUM_THREADS=1: the CPU arm is entitled to every host core
    c_port.set_threads(os.cpu_count() or 1)
    cores = c_port.threads()
    psi = np.zeros(1 << n_qubits, dtype=np.complex128)
    psi[0] = 1.0
    # probe one H psi to decide what a bounded "step" can be
    t0 = time.perf_counter()
    c_port.apply_h(psi, e, n_qubits, om[-1], de[-1])
    t_probe = time.perf_counter() - t0
    full_steps = 10.0 * t_probe * (steps + warmup) <= budget_s
    amps = float(1 << n_qubits)
    if full_steps:
        psi, _ = c_port.evolve_schedule(psi, e, n_qubits, om[:warmup], de[:warmup], dt[:warmup])
        t0 = time.perf_counter()
        psi, n_hpsi = c_port.evolve_schedule(psi, e, n_qubits, om[warmup:], de[warmup:], dt[warmup:])
        t = time.perf_counter() - t0
        sample = f"{steps} full sweep steps of the {n_qubits}-qubit workload ({n_hpsi} H psi), after {warmup} warm-up steps"
        return BYTES_PER_AMP * amps * n_hpsi / t / 1e9, steps / t, {
            "cores": cores, "sample": sample, "n_hpsi": n_hpsi, "seconds": t, "ms_per_step": 1e3 * t / steps}
    # too slow for whole steps: each step = ONE H psi application of the same workload
    rng = np.random.default_rng(1234)
    psi = rng.standard_normal(1 << n_qubits) + 1j * rng.standard_normal(1 << n_qubits)
    psi /= np.linalg.norm(psi)
    for _ in range(warmup):
        c_port.apply_h(psi, e, n_qubits, om[-1], de[-1])
    t0 = time.perf_counter()
    for _ in range(steps):
        c_port.apply_h(psi, e, n_qubits, om[-1], de[-1])
    t = time.perf_counter() - t0
    sample = (f"{steps} single H psi applications of the {n_qubits}-qubit workload (a full sweep step is ~9 of them; "
              f"whole steps would exceed the {budget_s:.0f} s budget)")
    return BYTES_PER_AMP * amps * steps / t / 1e9, steps / t / 9.0, {
        "cores": cores, "sample": sample, "n_hpsi": steps, "seconds": t, "ms_per_step": 1e3 * t / steps}


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    n = args.qubits + (world.bit_length() - 1)
    gbs, sps, info = cpu_port_run(n, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": "hpsi_hbm_gbs_in_sweep", "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": info["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
        "steps_per_s": sps,
        "config": make_config(n, world),
        "arm": "CPU port of the reference path (oracle/c/rydberg_ref.c, OpenMP): QuTiP is not installable here",
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": info["cores"], "kind": "port", "sample": info["sample"]},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------
# parity of the timed workload (outside every timed region): GPU against the oracle's C port
# --------------------------------------------------------------------------------------------
def parity_check(eng, n: int, rank: int, world: int, qbits, om, de, dt, steps: int = 2) -> dict | None:
    """`steps` mid-ramp sweep steps of the benchmarked workload from a seeded random state, on the GPU (all
    ranks) and through the C port of the reference path (rank 0, host cores): infidelity and max |delta psi|.
    qse/calc/qutip.py:34-53; tolerances of BASELINE.json: 1e-10 / 1e-9."""
    from oracle import c_port
    from oracle import rydberg_oracle as orc

    p = world.bit_length() - 1
    nl = n - p

    def shard(r):
        rng = np.random.default_rng(777 + 64 * n + r)
        return (rng.standard_normal(1 << nl) + 1j * rng.standard_normal(1 << nl)) / np.sqrt(2.0 ** (n + 1))

    mid = len(om) // 2
    sel = slice(mid, mid + steps)
    eng.set_state(shard(rank))
    eng.evolve_schedule(om[sel], de[sel], dt[sel])
    got = eng.get_state()
    if world > 1:
        from qse_b200 import distributed as qdist

        got = qdist.gather_shards(got)
    if rank != 0:
        return None
    t0 = time.perf_counter()
    psi = np.concatenate([shard(r) for r in range(world)])
    c_port.set_threads(os.cpu_count() or 1)
    e = c_port.build_diagonal(orc.interaction_matrix(qbits.positions), n)
    want, n_hpsi = c_port.evolve_schedule(psi, e, n, om[sel], de[sel], dt[sel])
    nrm = float(np.vdot(psi, psi).real)
    return {"against": "oracle C port (oracle/c/rydberg_ref.c) on the host cores, same seeded state and schedule values",
            "steps": steps, "infidelity": float(orc.infidelity(got / np.sqrt(nrm), want / np.sqrt(nrm))),
            "max_abs_dpsi": float(np.max(np.abs(got - want))), "tolerance": {"infidelity": 1e-10, "max_abs_dpsi": 1e-9},
            "oracle_hpsi": n_hpsi, "seconds": time.perf_counter() - t0}


# --------------------------------------------------------------------------------------------
# north-star configurations (BASELINE.json configs 3, 4, 5): extra keys, outside the headline timing
# --------------------------------------------------------------------------------------------
def north_star_single_gpu(local_rank: int, peak: float) -> dict:
    """Config 3's per-GPU shard size on ONE GPU (kagome 5x2, 30 qubits: 16 GiB of state) and config 5 (28
    random-position qubits, detuning sweep, bit-string read-out on the device)."""
    import qse_b200 as qb
    from qse_b200 import lib

    out = {}
    a = 0.9 * qb.blockade_radius(OMEGA_MAX)
    # --- config 3 shard: kagome 30 qubits, quench ------------------------------------------------
    reg = qb.lattices.kagome(a, 5, 2)
    n = reg.nqbits
    t0 = time.perf_counter()
    with lib.Engine(n, device=local_rank) as eng:
        eng.set_interaction(qb.interaction_matrix(reg.positions))
        setup = time.perf_counter() - t0
        eng.reset_state()
        T = 2
        ms = eng.time_schedule(np.full(T, OMEGA_MAX), np.zeros(T), np.full(T, 0.001))  # quench: constant Omega, delta = 0
        terms = eng.metrics()["last_terms"]
        norm = eng.norm()
        hp = np.median(eng.time_apply_h(OMEGA_MAX, 0.0, 5, False)[1:])
        eng.evolve_schedule(np.full(T, OMEGA_MAX), np.zeros(T), np.full(T, -0.001))
        idx, pr = eng.topk(1)
        out["config3_kagome_30q_1gpu"] = {
            "qubits": n, "hpsi_ms": float(hp), "hpsi_gbs_40B": BYTES_PER_AMP * 2.0 ** n / hp / 1e6,
            "frac_hbm": BYTES_PER_AMP * 2.0 ** n / hp / 1e6 / peak, "step_ms": [float(x) for x in ms],
            "steps_per_s": 1e3 / float(np.mean(ms)), "taylor_terms": terms, "norm": norm,
            "round_trip_infidelity": 1.0 - float(pr[0]) if int(idx[0]) == 0 else 1.0, "setup_s": setup}
    # --- config 5: 28 random-position qubits, detuning sweep, top-k bit strings --------------------
    reg = qb.lattices.random_register(28, a, seed=20261019)
    T = 3
    amp = qb.Signal(np.full(T, OMEGA_MAX), T)
    det = qb.Signal(np.linspace(-OMEGA_MAX, 2 * OMEGA_MAX, T), T)
    calc = qb.B200(reg, amp, det, device=local_rank, return_state=False)
    t0 = time.perf_counter()
    calc.calculate()
    t_calc = time.perf_counter() - t0
    t0 = time.perf_counter()
    top = calc.bitstring_probabilities(8)
    t_top = time.perf_counter() - t0
    eng = calc._engine
    hp = np.median(eng.time_apply_h(OMEGA_MAX, 1.0, 5, False)[1:])
    out["config5_random_28q_mis"] = {
        "qubits": 28, "calculate_s": t_calc, "steps": T, "topk_s": t_top, "top_bitstring": top[0][0], "top_probability": top[0][1],
        "norm": eng.norm(), "hpsi_ms": float(hp), "frac_hbm": BYTES_PER_AMP * 2.0 ** 28 / hp / 1e6 / peak,
        "evolve_s": calc.metrics["evolve_s"], "hpsi_calls": calc.metrics["hpsi_calls"]}
    calc.close()
    return out


def north_star_8gpu(rank: int, world: int, local_rank: int, peak: float, qdist) -> dict | None:
    """Config 4: triangular 11x3, 33 qubits, 128 GiB of state over 8 GPUs: H psi x5, two sweep steps, norm,
    exact time reversal back to |0...0>."""
    import qse_b200 as qb
    from qse_b200 import lib

    reg = qb.lattices.triangular(0.9 * qb.blockade_radius(OMEGA_MAX), 11, 3)
    n = reg.nqbits
    t0 = time.perf_counter()
    eng = lib.Engine(n, device=local_rank, rank=rank, nranks=world, nccl_uid=qdist.broadcast_unique_id())
    eng.set_interaction(qb.interaction_matrix(reg.positions))
    setup = time.perf_counter() - t0
    T = 2
    om = np.linspace(0.3 * OMEGA_MAX, OMEGA_MAX, T)
    de = np.linspace(-OMEGA_MAX, OMEGA_MAX, T)
    dt = np.full(T, 0.001)
    eng.reset_state()
    ms = eng.time_schedule(om, de, dt)
    terms = eng.metrics()["last_terms"]
    norm = eng.norm()
    hp = qdist.max_over_ranks(float(np.median(eng.time_apply_h(OMEGA_MAX, 0.5, 5, False)[1:])))
    step = [qdist.max_over_ranks(float(x)) for x in ms]
    eng.evolve_schedule(om[::-1], de[::-1], -dt[::-1])
    idx, pr = eng.topk(1)
    eng.close()
    if rank != 0:
        return None
    amps = 2.0 ** n
    inbound = 16.0 * amps / world * 2.0 * (world - 1) / world
    return {"config4_triangular_33q_8gpu": {
        "qubits": n, "state_gib": amps * 16 / 2 ** 30, "hpsi_ms": hp, "hpsi_gbs_40B_aggregate": BYTES_PER_AMP * amps / hp / 1e6,
        "frac_hbm": BYTES_PER_AMP * amps / hp / 1e6 / (peak * world), "frac_nvlink": inbound / hp / 1e6 / 770.0,
        "nvlink_inbound_bytes_per_hpsi_per_gpu": inbound, "step_ms": step, "steps_per_s": 1e3 / float(np.mean(step)),
        "taylor_terms": terms, "norm": norm, "round_trip_infidelity": 1.0 - float(pr[0]) if int(idx[0]) == 0 else 1.0,
        "setup_s": setup}}


# --------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------
def run_b200(args, rank: int, world: int, local_rank: int):
    import torch

    import qse_b200 as qb
    from qse_b200 import lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    p = world.bit_length() - 1
    n = args.qubits + p
    K, W = args.steps, args.warmup
    qbits, amp, det = workload(n, W + K)
    om, de, dt = qb.flatten_schedule(qb.Signals([amp]), qb.Signals([det]))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    uid = None
    if world > 1:
        from qse_b200 import distributed as qdist

        uid = qdist.broadcast_unique_id()
    eng = lib.Engine(n, device=local_rank, rank=rank, nranks=world, nccl_uid=uid)
    eng.set_interaction(qb.interaction_matrix(qbits.positions))
    amps_total = float(1 << n)

    # ---- value: device-resident sweep ---------------------------------------------------
    eng.reset_state()
    eng.evolve_schedule(om[:W], de[:W], dt[:W])  # W untimed warm-up steps
    m0 = eng.metrics()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    step_ms = eng.time_schedule(om[W:], de[W:], dt[W:])  # CUDA events on the library's stream, per step
    barrier()
    t_wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    m1 = eng.metrics()
    n_hpsi = m1["hpsi_calls"] - m0["hpsi_calls"]
    launches = m1["kernel_launches"] - m0["kernel_launches"]
    t_dev = float(np.sum(step_ms)) * 1e-3
    t_region = max(t_wall, t_dev)
    if dist is not None:
        from qse_b200 import distributed as qdist

        t_region = qdist.max_over_ranks(t_region)
    value = BYTES_PER_AMP * amps_total * n_hpsi / t_region / 1e9
    norm = eng.norm()

    # ---- roofline leg: H psi alone, CUDA events around each application ----------------------
    hp_ms = eng.time_apply_h(float(om[-1]), float(de[-1]), 12, False)[2:]
    if dist is not None:
        hp_ms = np.array([qdist.max_over_ranks(float(np.median(hp_ms)))])
    t_hpsi = float(np.median(hp_ms)) * 1e-3
    peak, peak_src = measured_peaks()
    achieved = BYTES_PER_AMP * amps_total / t_hpsi / 1e9 / world  # per GPU
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get(str(n - p))  # by LOCAL qubits: DRAM bytes of one GPU's own passes
    met = eng.metrics()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": f"H psi = {met['hpsi_passes']} launches of hpsi_ws_kernel (fused super-block launch + plain pass), timed as a group",
                "algorithmic_bytes_per_launch_group": BYTES_PER_AMP * amps_total / world,
                "planned_hbm_bytes_per_amp": met["hpsi_bytes_per_amp"],
                "hpsi_ms": t_hpsi * 1e3,
                "in_sweep": {"achieved": value / world, "frac": value / world / peak,
                             "note": "fused Taylor step: same 40 B/amp currency, region time / number of H psi"}}

    if world > 1:
        # bytes pulled over NVLink per H psi per GPU: p partner shards (direct exchange, P = 2) or
        # 2 * (P-1)/P of a shard (qubit-remap exchange, P >= 4)
        inbound = 16.0 * (1 << (n - p)) * (p if world < 4 else 2.0 * (world - 1) / world)
        roofline["nvlink"] = {"bound": "nvlink", "achieved": inbound / t_hpsi / 1e9, "peak": 770.0, "unit": "GB/s",
                              "frac": inbound / t_hpsi / 1e9 / 770.0, "peak_source": "measured peer copy per direction "
                              "(B200_PROFILING.md); nominal 900", "bytes_per_hpsi_per_gpu": inbound,
                              "note": "partner amplitudes are read inside pass 0; time = whole H psi"}

    # ---- parity of this very workload against the oracle (not timed) ---------------------------
    parity = None if args.no_parity else parity_check(eng, n, rank, world, qbits, om, de, dt)

    # ---- e2e: the plugin call with host buffers -------------------------------------------
    eng.close()
    k_amp = qb.Signal(amp.values[W:], K)
    k_det = qb.Signal(det.values[W:], K)
    calc = qb.B200(qbits, qb.Signal(amp.values[:W], W), qb.Signal(det.values[:W], W), device=local_rank,
                   return_state=(True if world == 1 else "shard"))
    calc.calculate()  # warm-up: builds the context and the diagonal
    calc.amplitude, calc.detuning = qb.Signals([k_amp]), qb.Signals([k_det])
    barrier()
    t0 = time.perf_counter()
    out = calc.calculate()
    barrier()
    t_e2e = time.perf_counter() - t0
    if dist is not None:
        t_e2e = qdist.max_over_ranks(t_e2e)
    e2e_hpsi = calc.metrics["hpsi_calls"]
    e2e_val = BYTES_PER_AMP * amps_total * e2e_hpsi / t_e2e / 1e9
    state_bytes = out["state"].nbytes if "state" in out else 0
    h2d = (3 * 8 * K) / K
    d2h = (state_bytes + n * 3 * 8) / K  # per rank: its own shard of the state + the spins
    e2e_launches = calc.metrics["kernel_launches"]
    calc.close()

    line = {
        "metric": "hpsi_hbm_gbs_in_sweep", "value": value, "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * t_region / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "complex128", "data": "synthetic",
        "steps_per_s": K / t_region,
        "config": make_config(n, world),
        "hpsi_per_step": n_hpsi / K,
        "roofline": roofline,
        "parity": parity,
        "e2e": {"value": e2e_val, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps_per_s": K / t_e2e, "seconds": t_e2e,
                "note": "qse_b200.B200.calculate() for the K timed values: schedule upload, sweep, spins reduced on "
                        "the device, final state copied back to a page-locked host ndarray"
                        + (" (every rank its own shard, return_state='shard'; bytes are per rank)" if world > 1 else "")},
        "gpu_launches": int(launches),
        "e2e_gpu_launches": int(e2e_launches),
        "clocks": clocks,
        "norm_after_sweep": norm,
        "device_event_ms_per_step": float(np.mean(step_ms)),
    }
    if not args.no_north_star:
        if world == 1:
            line["north_star"] = north_star_single_gpu(local_rank, peak)
        elif world == 8:
            ns = north_star_8gpu(rank, world, local_rank, peak, qdist)
            if ns is not None:
                line["north_star"] = ns
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gbs, sps, info = cpu_port_run(n, 4, 1, budget_s=40.0)
        line["cpu_baseline"] = {"value": gbs, "unit": "GB/s", "cores": info["cores"], "kind": "port",
                                "sample": info["sample"], "steps_per_s": sps}
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) to stderr."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)  # libraries that write to fd 1 (NCCL's version banner) now land on stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--qubits", type=int, default=25, help="qubits per 1-GPU shard (default: BASELINE config 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity leg (developer runs)")
    ap.add_argument("--no-north-star", action="store_true", help="skip the BASELINE config 3/4/5 extras")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rules: W >= 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
