"""A sharded sweep: one process per GPU, the state split by its top log2(P) qubits.

    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 examples/sharded_sweep.py [n_x n_y]

Every rank builds the same calculator and calls the same methods (SPMD); torch.distributed only
carries the NCCL-id broadcast, the exchange of the sharded qubits happens inside libqsb (peer
memory over NVLink).  Observables are identical on every rank; the state stays on the devices
(`return_state=False`), which is what makes 28+ qubits practical.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import qse_b200 as qb  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4, 7)
    om = 2 * np.pi
    qbits = qb.lattices.square(0.9 * qb.blockade_radius(om), nx, ny)
    t = 20
    calc = qb.B200(qbits, qb.Signal(np.linspace(0, om, t), t), qb.Signal(np.linspace(-2 * om, 2 * om, t), t),
                   return_state=False)
    out = calc.calculate(e_ops=[calc.get_hamiltonian(om, 2 * om)])
    if dist.get_rank() == 0:
        print(f"{qbits.nqbits} qubits on {dist.get_world_size()} GPUs: {calc.metrics['steps']} steps in "
              f"{calc.metrics['evolve_s']:.3f} s, <H_f> = {out['expectations'][-1, 0]:.6f}")
        print("<n_i> =", np.round(calc.get_number_operator(), 5))
        print("most probable bit strings:", calc.bitstring_probabilities(3))
    else:
        calc.get_number_operator()
        calc.bitstring_probabilities(3)  # collective calls: every rank takes part
    calc.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
