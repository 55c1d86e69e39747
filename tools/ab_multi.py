"""Developer tool (torchrun, one rank per GPU): H psi / sweep-step time of the sharded engine under different
exchange options, 2^25 amplitudes per GPU by default.

    python -m torch.distributed.run --nproc-per-node P --master-addr 127.0.0.1 tools/ab_multi.py [--local 25] \
        exchange_split=0 exchange_split=1 exchange=0,exchange_split=2
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

import bench
import qse_b200 as qb
from qse_b200 import distributed as qdist
from qse_b200 import lib

DEFAULTS = {"exchange": -1, "exchange_split": -1, "remap_ctas": 32, "sb_bits": 0, "sb_lag": 0}

args = sys.argv[1:]
local_q = 25
if args and args[0] == "--local":
    local_q = int(args[1])
    args = args[2:]
sets = [dict((kv.split("=")[0], float(kv.split("=")[1])) for kv in a.split(",")) for a in args] or [{}]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = local_q + world.bit_length() - 1
if n in bench.SHAPES:
    qbits, _, _ = bench.workload(n, 8)
else:
    qbits = qb.lattices.random_register(n, 0.9 * qb.blockade_radius(2 * np.pi), seed=n)
eng = lib.Engine(n, device=local, rank=rank, nranks=world, nccl_uid=qdist.broadcast_unique_id())
eng.set_interaction(qb.interaction_matrix(qbits.positions))
eng.reset_state()
eng.evolve_schedule(np.full(2, 3.0), np.full(2, -4.0), np.full(2, 0.001))
for rep in range(2):
    for opts in sets:
        for k, v in {**DEFAULTS, **opts}.items():
            eng.set_option(k, v)
        ms = qdist.max_over_ranks(float(np.median(eng.time_apply_h(3.0, -4.0, 10, False)[3:])))
        steps = 5
        st = eng.time_schedule(np.linspace(2.0, 6.0, steps), np.linspace(-5, 10, steps), np.full(steps, 0.001))[2:]
        st = qdist.max_over_ranks(float(np.median(st)))
        nrm = eng.norm()
        if rank == 0:
            print(json.dumps({"n": n, "gpus": world, "opts": opts, "hpsi_ms": round(ms, 4), "step_ms": round(st, 3),
                              "terms": eng.metrics()["last_terms"], "norm": nrm}), flush=True)
eng.close()
dist.destroy_process_group()
