#!/usr/bin/env python
"""N-GPU check of the sharded path, run under torchrun with NCCL:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py

BASELINE config 4 at reduced size: theta[N, 24] sharded contiguously over the ranks, the fused 3-shift
energy kernel on each shard (no data-path collective), local (min, argmin) by qmps_argmin, one 16-byte
all-gather over NCCL.  Rank 0 also evaluates the UNSHARDED batch on its own GPU and checks that the global
(min energy, argmin) and the per-shard energies are bit-identical.  Device-timed, max over ranks.
Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from qmps_b200 import batched as B, represent as R
    from qmps_b200.dist import shard_range, global_argmin
    from qmps_b200.ground_state import Hamiltonian
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, P = 16384, 24
    theta_all = np.random.default_rng(3).normal(size=(N, P))
    prog = R.ShallowCNOTStateTensor_nonuniform(8, np.zeros(P)).program()
    H = Hamiltonian({'XX': 1, 'YY': 1, 'ZZ': 1}).to_matrix()
    lo, hi = shard_range(N, rank, world)
    theta = torch.from_numpy(theta_all[lo:hi]).to(dev)

    def step():
        e = B.energy_theta(prog, theta, H, coord=5, shifts=B.ROTO3_SHIFTS)       # [n_local, 3]
        bc, bi = B.argmin(e.reshape(-1), index_offset=3 * lo)
        return e, global_argmin(bc, bi)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 10
    for _ in range(K):
        e, (best, idx) = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / K], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ok = True
    if rank == 0:
        full = B.energy_theta(prog, torch.from_numpy(theta_all).to(dev), H, coord=5, shifts=B.ROTO3_SHIFTS)
        fb, fi = B.argmin(full.reshape(-1))
        ok = bool((full[lo:hi] == e).all()) and float(fb) == float(best) and int(fi) == int(idx)
        print(json.dumps({"what": "cfg 4 sharded: 3-shift D=8 energies + NCCL argmin", "world": world, "N": N,
                          "ms_per_step": float(ms), "evals_per_s": 3 * N / float(ms) * 1e3,
                          "global_min": float(best), "global_argmin": int(idx), "matches_unsharded": ok}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
