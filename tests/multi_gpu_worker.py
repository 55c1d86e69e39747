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


def large_shard_cases(rank, world, local, check):
    """The shard sizes bench.py / SCALE run (2^25 amplitudes per GPU: 26 qubits on 2, 27 on 4, 28 on 8
    GPUs), against the oracle's C port: plain launch + fused Taylor epilogue + (from 4 ranks) the remap
    exchange with its staggered super-block order -- none of which the small cases reach.  Also the
    direct exchange at the same size and a sharded checkpoint round trip."""
    import tempfile

    import bench
    from oracle import c_port

    p = world.bit_length() - 1
    n = int(os.environ.get("QSB_WORKER_LARGE_QUBITS", 25 + p))
    nl = n - p
    qbits, amp, det = bench.workload(n, 23)
    om, de, dt = qb.flatten_schedule(qb.Signals([amp]), qb.Signals([det]))
    sel = slice(10, 12)

    def shard(r):
        rng = np.random.default_rng(9000 + 16 * n + r)
        return (rng.standard_normal(1 << nl) + 1j * rng.standard_normal(1 << nl)) / np.sqrt(2.0 ** (n + 1))

    mine = shard(rank)
    uid = qdist.broadcast_unique_id()
    eng = lib.Engine(n, device=local, rank=rank, nranks=world, nccl_uid=uid)
    eng.set_interaction(qb.interaction_matrix(qbits.positions))
    eng.set_state(mine)
    y = {"default": eng.apply_h(om[11], de[11])}
    # every placement of the partner pulls: fused launch only / split between the launches / plain launch only,
    # with the direct and (from 4 ranks) the remap exchange, the remap kernel beside the fused launch or before it
    variants = [("direct, fused", {"exchange": 0, "exchange_split": 0}), ("direct, split", {"exchange": 0, "exchange_split": 1}),
                ("direct, plain", {"exchange": 0, "exchange_split": 2})]
    if world >= 4:
        variants += [("remap, fused", {"exchange": 1, "exchange_split": 0}), ("remap, plain", {"exchange": 1, "exchange_split": 2}),
                     ("remap, plain, serial", {"exchange": 1, "exchange_split": 2, "remap_ctas": 0})]
    for name, opts in variants:
        for k, v in opts.items():
            eng.set_option(k, v)
        y[name] = eng.apply_h(om[11], de[11])
        for k, v in {"exchange": -1, "exchange_split": -1, "remap_ctas": 32}.items():
            eng.set_option(k, v)
    eng.evolve_schedule(om[sel], de[sel], dt[sel])
    evolved = eng.get_state()
    spins = eng.spins()
    # rank 0 holds the reference: the same shards, the C port on the host cores
    full_y = {k: qdist.gather_shards(v) for k, v in y.items()}
    full_e = qdist.gather_shards(evolved)
    diag = qdist.gather_shards(eng.diagonal().astype(np.complex128)).real
    if rank == 0:
        psi = np.concatenate([shard(r) for r in range(world)])
        v = orc.interaction_matrix(qbits.positions)
        e = c_port.build_diagonal(v, n)
        c_port.set_threads(os.cpu_count() or 1)
        check(f"large n={n} diagonal bit-exact", np.array_equal(diag, e))
        want_h = c_port.apply_h(psi, e, n, om[11], de[11])
        for k, got in full_y.items():
            err = float(np.max(np.abs(got - want_h)) / np.max(np.abs(want_h)))
            check(f"large n={n} H psi ({k}) vs C port", err <= 1e-13, f"rel err {err:.2e}")
        want, _ = c_port.evolve_schedule(psi, e, n, om[sel], de[sel], dt[sel])
        nrm = float(np.vdot(psi, psi).real)
        infid = orc.infidelity(full_e / np.sqrt(nrm), want / np.sqrt(nrm))
        check(f"large n={n} 2-step sweep infidelity vs C port", infid <= 1e-10, f"{infid:.2e}")
        dmax = float(np.max(np.abs(full_e - want)))
        check(f"large n={n} 2-step sweep state", dmax <= 1e-9, f"{dmax:.2e}")
        d = float(np.max(np.abs(spins - c_port.spins(want, n))))
        check(f"large n={n} spins after the sweep", d <= 1e-9, f"{d:.2e}")
    eng.close()
    # sharded checkpoint: stop after 2 of 4 values, save, resume in a new calculator, compare with one go
    tmp = [tempfile.mkdtemp(prefix="qsb_ckpt_") if rank == 0 else None]
    dist.broadcast_object_list(tmp, src=0)
    prefix = os.path.join(tmp[0], "sweep")
    a4, d4 = qb.Signal(amp.values[10:14], 4), qb.Signal(det.values[10:14], 4)
    calc = qb.B200(qbits, a4, d4, return_state=False, compute_spins=False)
    calc.calculate()
    whole = calc._engine.get_state().copy()
    calc.calculate(stop_after=2)
    calc.save_checkpoint(prefix, 2)
    calc.close()
    dist.barrier()
    calc = qb.B200(qbits, a4, d4, return_state=False, compute_spins=False)
    calc.calculate(resume_from=prefix)
    resumed = calc._engine.get_state()
    # not bit-identical: a fresh context plans its first Taylor terms one at a time, the uninterrupted run in pairs
    dev = torch.tensor([float(np.max(np.abs(resumed - whole)))], device="cuda")
    dist.all_reduce(dev, op=dist.ReduceOp.MAX)
    check(f"large n={n} sharded checkpoint resume", dev.item() <= 1e-13, f"max |delta psi| {dev.item():.2e}")
    calc.close()
    dist.barrier()
    if rank == 0:
        import shutil

        shutil.rmtree(tmp[0], ignore_errors=True)


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
    quick = os.environ.get("QSB_WORKER_QUICK", "0") != "0"  # __graft_entry__.smoke(): one size, peer-memory exchange
    modes = tuple(os.environ["QSB_WORKER_MODES"].split(",")) if os.environ.get("QSB_WORKER_MODES") else None
    sizes = tuple(int(v) for v in os.environ["QSB_WORKER_SIZES"].split(",")) if os.environ.get("QSB_WORKER_SIZES") else None
    for exchange in (modes or (("ipc",) if quick else ("ipc", "nccl"))):
        if exchange == "nccl":
            os.environ["QSB_NO_IPC"] = "1"
        else:
            os.environ.pop("QSB_NO_IPC", None)
        for n in (sizes or ((14,) if quick else (8, 14, 17, 19) if world >= 4 else (8, 14, 17))):
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
            if exchange == "ipc" and n >= 14:
                # generalised H: per-site weights + operator strings that flip, sign and
                # project on SHARDED qubits (qubit 0 is always sharded), applied to a caller-supplied state
                rs = np.random.default_rng(5 * n)
                w_amp, w_det = rs.uniform(0.5, 1.5, n), rs.uniform(-1.0, 1.0, n)
                ham = reg.compute_interaction_hamiltonian(lambda dd: 2.0 / dd**3, ["X", "X"])
                ham += reg.compute_interaction_hamiltonian(lambda dd: 2.0 / dd**3, ["Y", "Y"])
                ham += qb.Operator(["Y", "Z"], [0, n - 1], n, coef=0.8)
                ham += qb.Operator(["Z", "Z"], [0, 1], n, coef=-0.9)
                ham += qb.Operator(["N", "X"], [0, 3], n, coef=0.35)
                ham += qb.Operator("X", 0, n, coef=0.25)
                eng.set_site_weights(w_amp, w_det)
                eng.set_operator_terms(qb.operator_terms(ham, n))
                yg = eng.apply_h_to(psi[sl], 3.1, -2.2)
                wantg = orc.general_apply(psi, n, e_int=e, hw_sites=0.5 * 3.1 * w_amp, d_sites=-2.2 * w_det,
                                          terms=orc.operators_as_terms(ham))
                errg = float(np.max(np.abs(yg - wantg[sl])) / np.max(np.abs(wantg)))
                check(f"n={n} generalised H psi (site weights + operator strings)", errg <= 1e-13, f"rel err {errg:.2e}")
                eng.set_operator_terms([])
                eng.set_site_weights(None, None)
                y2 = eng.apply_h_to(psi[sl], 3.1, -2.2)
                check(f"n={n} H restored after clearing the operator terms", float(np.max(np.abs(y2 - want[sl])) / np.max(np.abs(want))) <= 1e-13)
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
            calc = qb.B200(reg, amp, det)  # default arguments: spins are reduced on the device in both exchange modes
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
            d = float(np.max(np.abs(calc.spins - orc.get_spins(ref["state"], n))))
            check(f"{exchange} n={n} sweep spins", d <= 1e-9, f"{d:.2e}")
            calc.close()
    os.environ.pop("QSB_NO_IPC", None)
    if os.environ.get("QSB_WORKER_LARGE", "1") != "0" and not quick:
        large_shard_cases(rank, world, local, check)
    t = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(t)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI-GPU PARITY", "GREEN" if t.item() == 0 else f"RED ({int(t.item())} failures)", flush=True)
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()
