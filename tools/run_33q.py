"""North-star demonstration (torchrun, 8 GPUs): triangular 11x3 lattice, 33 qubits (2^33 amplitudes =
128 GiB of state sharded by the top 3 qubits), a short adiabatic sweep, checked through
size-independent properties; prints one JSON line on rank 0.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_33q.py [--qubits 33]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import qse_b200 as qb  # noqa: E402
from qse_b200 import distributed as qdist  # noqa: E402
from qse_b200 import lib  # noqa: E402

SHAPES = {33: (11, 3), 30: (10, 3), 27: (9, 3), 24: (8, 3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=33)
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.qubits
    om_max = 2 * np.pi
    qbits = qb.lattices.triangular(0.9 * qb.blockade_radius(om_max), *SHAPES[n])
    assert qbits.nqbits == n
    t0 = time.perf_counter()
    eng = lib.Engine(n, device=local, rank=rank, nranks=world, nccl_uid=qdist.broadcast_unique_id())
    eng.set_interaction(qb.interaction_matrix(qbits.positions))
    t_setup = time.perf_counter() - t0
    T = args.steps
    om = np.linspace(0.3 * om_max, om_max, T)
    de = np.linspace(-om_max, om_max, T)
    dt = np.full(T, 0.001)
    eng.reset_state()
    ms = eng.time_schedule(om, de, dt)
    norm = eng.norm()
    number = eng.number()
    hp = eng.time_apply_h(om_max, 0.5, 3)[1:]
    eng.evolve_schedule(om[::-1], de[::-1], -dt[::-1])  # exact time reversal back to |0...0>
    idx, p = eng.topk(1)
    met = eng.metrics()
    t_hpsi = qdist.max_over_ranks(float(np.median(hp))) * 1e-3
    if rank == 0:
        amps = float(1 << n)
        p_ = world.bit_length() - 1
        print(json.dumps({
            "workload": f"triangular {SHAPES[n][0]}x{SHAPES[n][1]} ({n} qubits) sweep on {world} GPUs, complex128",
            "state_gib": amps * 16 / 2**30, "amplitudes_per_gpu": int(amps) // world, "setup_s": t_setup,
            "step_ms": [float(x) for x in ms], "taylor_terms": met["last_terms"],
            "hpsi_ms": t_hpsi * 1e3, "hpsi_gbs_40B_aggregate": 40.0 * amps / t_hpsi / 1e9,
            "hpsi_frac_of_aggregate_hbm_peak": 40.0 * amps / t_hpsi / 1e9 / (6563.2 * world),
            # qubit-remap exchange from 4 ranks up: 2 * (P-1)/P of a shard per H psi; else p full shards
            "nvlink_inbound_gbs_per_gpu": 16.0 * amps / world * (p_ if world < 4 else 2.0 * (world - 1) / world) / t_hpsi / 1e9,
            "norm_after_sweep": norm, "max_rydberg_population": float(number.max()),
            "round_trip_infidelity": 1.0 - float(p[0]), "round_trip_argmax": int(idx[0]),
            "hpsi_launches": met["hpsi_passes"], "planned_hbm_bytes_per_amp": met["hpsi_bytes_per_amp"]}), flush=True)
        assert abs(norm - 1.0) < 1e-10 and int(idx[0]) == 0 and 1.0 - float(p[0]) < 1e-10
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
