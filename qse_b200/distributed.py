"""Host-side plumbing for the sharded (one process per GPU) calculator.

``torch.distributed`` is used for bootstrap and optional gathers only; the data-path exchange
of the sharded qubits happens inside libqsb.  Everything here works on the ``gloo`` backend
too, which is how the CPU tests exercise it (tests/test_distributed_cpu.py).
"""

from __future__ import annotations

import numpy as np


def _dist():
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("torch.distributed process group is not initialised")
    return dist


class ShardPlan:
    """Index arithmetic of sharding a 2^N state vector over P = 2^p ranks by its top p qubits.

    Rank r owns basis indices [r * 2^(N-p), (r+1) * 2^(N-p)); qubit q is bit N-1-q of the index
    (qse/magnetic.py:51), so qubits 0..p-1 are the sharded ones and qubit g's partner of rank r is
    r ^ (1 << (p-1-g)).
    """

    def __init__(self, nqbits: int, world: int):
        if world < 1 or world & (world - 1):
            raise ValueError("the number of ranks must be a power of two")
        self.n = int(nqbits)
        self.world = int(world)
        self.p = self.world.bit_length() - 1
        if self.p >= self.n and world > 1:
            raise ValueError("more ranks than half the basis")
        self.nl = self.n - self.p
        self.local_dim = 1 << self.nl

    def owner(self, k: int) -> int:
        return k >> self.nl

    def local_index(self, k: int) -> int:
        return k & (self.local_dim - 1)

    def shard_slice(self, rank: int) -> slice:
        return slice(rank * self.local_dim, (rank + 1) * self.local_dim)

    def sharded_qubits(self) -> list[int]:
        return list(range(self.p))

    def partner(self, rank: int, qubit: int) -> int:
        if not 0 <= qubit < self.p:
            raise ValueError("qubit is not sharded")
        return rank ^ (1 << (self.p - 1 - qubit))

    def rank_popcount(self, rank: int) -> int:
        return bin(rank).count("1")

    def rank_diagonal_offset(self, rank: int, v: np.ndarray) -> float:
        """sum_{i<j both sharded and set in rank} V_ij — the constant E_int part of a shard."""
        bits = [(rank >> (self.p - 1 - q)) & 1 for q in range(self.p)]
        return float(sum(v[i, j] for i in range(self.p) for j in range(i + 1, self.p) if bits[i] and bits[j]))


def broadcast_unique_id() -> bytes:
    """Rank 0 asks libqsb for an NCCL unique id; everyone receives it."""
    dist = _dist()
    from . import lib

    box = [lib.nccl_unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def gather_shards(local: np.ndarray) -> np.ndarray:
    """Concatenate per-rank shards in rank order = the reference's basis order."""
    dist = _dist()
    import torch

    world = dist.get_world_size()
    if world == 1:
        return local
    t = torch.from_numpy(np.ascontiguousarray(local).view(np.float64))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return np.concatenate([o.cpu().numpy().view(np.complex128) for o in out])


def max_over_ranks(value: float) -> float:
    dist = _dist()
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
