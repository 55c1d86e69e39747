"""world_size-2 gloo tests (CPU) of the host-side logic of the sharded path: shard gathering,
max-over-ranks timing, and the exchange algorithm libqsb implements on the GPU — restated with
numpy shards and REAL inter-process send/recv, checked against the full-vector oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, results):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import rydberg_oracle as orc
        from qse_b200 import distributed as qdist
        from qse_b200.host import lattices

        plan = qdist.ShardPlan(n, world)
        reg = lattices.random_register(n, 6.0, seed=n)
        v = orc.interaction_matrix(reg.positions)
        e = orc.diagonal_energies(v, n)
        rng = np.random.default_rng(5)
        psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        psi /= np.linalg.norm(psi)
        omega, delta = 2.7, -1.3
        sl = plan.shard_slice(rank)
        local = psi[sl].copy()

        # --- sharded H psi: local part + one pairwise exchange per sharded qubit -------------
        k = np.arange(plan.local_dim)
        pop = np.array([bin(int(x)).count("1") for x in k]) + plan.rank_popcount(rank)
        y = (e[sl] - delta * pop) * local
        for q in range(plan.nl):
            y = y + 0.5 * omega * local[k ^ (1 << q)]
        for g in plan.sharded_qubits():
            partner = plan.partner(rank, g)
            send = torch.from_numpy(local.view(np.float64).copy())
            recv = torch.empty_like(send)
            ops = [dist.P2POp(dist.isend, send, partner), dist.P2POp(dist.irecv, recv, partner)]
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            y = y + 0.5 * omega * recv.numpy().view(np.complex128)
        want = orc.apply_hamiltonian(psi, e, n, omega, delta)
        err_h = float(np.max(np.abs(y - want[sl])))

        # --- observables: per-shard partial sums + all-reduce ----------------------------------
        prob = np.abs(local) ** 2
        number = np.zeros(n)
        for q in range(n):
            bit = n - 1 - q
            if bit < plan.nl:
                number[q] = np.sum(prob * ((k >> bit) & 1))
            else:
                number[q] = prob.sum() * ((rank >> (bit - plan.nl)) & 1)
        t = torch.from_numpy(number)
        dist.all_reduce(t)
        err_n = float(np.max(np.abs(t.numpy() - orc.get_number_operator(psi, n))))

        # --- host helpers ----------------------------------------------------------------------
        full = qdist.gather_shards(local)
        ok_gather = bool(np.array_equal(full, psi))
        tmax = qdist.max_over_ranks(float(rank + 1))
        results[rank] = (err_h, err_n, ok_gather, tmax)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 7), (4, 8)])
def test_sharded_exchange_over_gloo(world, n):
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, n, results), nprocs=world, join=True)
        assert len(results) == world
        for rank in range(world):
            err_h, err_n, ok_gather, tmax = results[rank]
            assert err_h < 1e-12, (rank, err_h)
            assert err_n < 1e-13, (rank, err_n)
            assert ok_gather
            assert tmax == float(world)
