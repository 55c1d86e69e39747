"""SPMD worker for the sharded path: run under torchrun with one rank per GPU.

    python -m torch.distributed.run --nproc-per-node P --master-addr 127.0.0.1 tests/multi_gpu_worker.py

Every rank builds the same calculator (state sharded by the top log2(P) qubits) and rank 0 checks
the gathered results against the CPU oracle.  Exit code 0 = parity green.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import qse_b200 as qb  # noqa: E402
from oracle import rydberg_oracle as orc  # noqa: E402
from qse_b200 import distributed as qdist  # noqa: E402
from qse_b200 import lib  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    failures = []

    def check(name, ok, detail=""):
        if rank == 0:
            print(f"[{'ok' if ok else 'FAIL'}] {name} {detail}", flush=True)
        if not ok:
            failures.append(name)

    omega_max = 2 * np.pi
    for exchange in ("ipc", "nccl"):
        if exchange == "nccl":
            os.environ["QSB_NO_IPC"] = "1"
        else:
            os.environ.pop("QSB_NO_IPC", None)
        for n in ((8, 14, 17, 19) if world >= 4 else (8, 14, 17)):
            reg = qb.lattices.random_register(n, 0.9 * qb.blockade_radius(omega_max), seed=n)
            plan = qdist.ShardPlan(n, world)
            v = orc.interaction_matrix(reg.positions)
            e = orc.diagonal_energies(v, n)
            # --- raw engine: diagonal, H psi on a random state ------------------------------
            uid = qdist.broadcast_unique_id()
            eng = lib.Engine(n, device=local, rank=rank, nranks=world, nccl_uid=uid)
            eng.set_interaction(qb.interaction_matrix(reg.positions))
            sl = plan.shard_slice(rank)
            check(f"{exchange} n={n} diagonal bit-exact", np.array_equal(eng.diagonal(), e[sl]))
            rng = np.random.default_rng(n)
            psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
            psi /= np.linalg.norm(psi)
            eng.set_state(psi[sl])
            y = eng.apply_h(3.1, -2.2)
            want = orc.apply_hamiltonian(psi, e, n, 3.1, -2.2)
            err = float(np.max(np.abs(y - want[sl])) / np.max(np.abs(want)))
            check(f"{exchange} n={n} H psi", err <= 1e-13, f"rel err {err:.2e}")
            if exchange == "ipc" and world >= 4:
                eng.set_option("exchange", 0)  # direct partner pulls instead of the remap exchange
                y0 = eng.apply_h(3.1, -2.2)
                err0 = float(np.max(np.abs(y0 - want[sl])) / np.max(np.abs(want)))
                check(f"n={n} H psi, direct exchange", err0 <= 1e-13, f"rel err {err0:.2e}")
                eng.set_option("exchange", -1)
            if exchange == "ipc":
                spins = eng.spins()
                d = float(np.max(np.abs(spins - orc.get_spins(psi, n))))
                check(f"n={n} spins on a random state", d <= 1e-12, f"{d:.2e}")
                if n <= 14:
                    d = float(np.max(np.abs(eng.sisj() - orc.get_sisj(psi, n))))
                    check(f"n={n} sisj on a random state", d <= 1e-12, f"{d:.2e}")
                idx, p = eng.topk(8)
                prob = np.abs(psi) ** 2
                order = np.lexsort((np.arange(prob.size), -prob))[:8]
                check(f"n={n} topk", np.array_equal(idx, order.astype(np.uint64)) and np.allclose(p, prob[order], rtol=1e-15))
            if exchange == "ipc":
                shots = 50000
                smp = eng.sample(shots, seed=17).astype(np.int64)
                prob = np.abs(psi) ** 2
                owner_freq = np.bincount(smp >> plan.nl, minlength=world) / shots
                owner_p = prob.reshape(world, -1).sum(axis=1)
                check(f"n={n} sampling across shards", bool(np.max(np.abs(owner_freq - owner_p)) < 6 * np.sqrt(0.25 / shots)
                                                            and np.all(prob[smp] > 0)))
            d = float(np.max(np.abs(eng.number() - orc.get_number_operator(psi, n))))
            check(f"{exchange} n={n} number", d <= 1e-12, f"{d:.2e}")
            check(f"{exchange} n={n} norm", abs(eng.norm() - 1.0) <= 1e-12)
            en = eng.energy(3.1, -2.2)
            check(f"{exchange} n={n} energy", abs(en - orc.energy(psi, e, n, 3.1, -2.2)) <= 1e-10 * max(1, abs(en)))
            eng.close()
            # --- calculator: sweep + e_ops, state gathered over ranks -------------------------
            t = 8
            amp = qb.Signal(np.linspace(0, omega_max, t), 4 * t)
            det = qb.Signal(np.linspace(-2 * omega_max, 2 * omega_max, t), 4 * t)
            calc = qb.B200(reg, amp, det, compute_spins=(exchange == "ipc"))
            h_f = calc.get_hamiltonian(omega_max, 2 * omega_max)
            out = calc.calculate(e_ops=[h_f])
            method = "eigh" if n <= 10 else "expm_multiply"
            ref = orc.calculate(reg.positions, [(amp.values, amp.duration)], [(det.values, det.duration)], method=method,
                                e_ops=[lambda s: orc.energy(s, e, n, omega_max, 2 * omega_max)])
            infid = orc.infidelity(out["state"], ref["state"])
            check(f"{exchange} n={n} sweep infidelity", infid <= 1e-10, f"{infid:.2e}")
            dmax = float(np.max(np.abs(out["state"] - ref["state"])))
            check(f"{exchange} n={n} sweep state", dmax <= 1e-9, f"{dmax:.2e}")
            dexp = float(np.max(np.abs(out["expectations"] - ref["expectations"])))
            check(f"{exchange} n={n} expectations", dexp <= 1e-9 * max(1.0, float(np.max(np.abs(ref["expectations"])))), f"{dexp:.2e}")
            if exchange == "ipc":
                d = float(np.max(np.abs(calc.spins - orc.get_spins(ref["state"], n))))
                check(f"n={n} sweep spins", d <= 1e-9, f"{d:.2e}")
            calc.close()
    t = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(t)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI-GPU PARITY", "GREEN" if t.item() == 0 else f"RED ({int(t.item())} failures)", flush=True)
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()
