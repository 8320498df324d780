#!/usr/bin/env python
"""Host-CPU rates of the reference-style per-call path (the oracle port, one process per core) for the
units of work of BASELINE configs 3, 4, 5, the D = 2 Loschmidt step and the brick-wall cost -- the "vs
host CPU" side of the metric for everything bench.py (config 2) does not cover.  Bounded samples of a few
seconds each; prints one JSON line per unit.  No GPU needed.

  python tools/cpu_baselines.py [--seconds 4]
"""
import os
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):   # one BLAS thread per worker process:
    os.environ[_v] = "1"                                                     # must be set before numpy loads BLAS
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _setup(kind):
    import oracle as O
    from oracle import brickwall as OB
    from scipy.linalg import expm
    from scipy.stats import unitary_group
    rng = np.random.default_rng(0)
    if kind in ("loschmidt_d2", "loschmidt_d4"):
        D = 2 if kind.endswith("d2") else 4
        tens = (lambda p: O.unitary_to_tensor(O.shallow_full_state_tensor(p))) if D == 2 else (lambda p: O.unitary_to_tensor(O.shallow_cnot_state_tensor_nonuniform(4, p)))
        P = 15 if D == 2 else 12
        A0 = tens(rng.normal(size=P))
        W = expm(-1j * O.tfim_matrix(0.2) * 0.04)
        thetas = rng.normal(size=(16, P))
        # the reference evaluates gate(theta) -> tensor -> cost per call (loschmidts/time_evo.py:75-116)
        return lambda k: O.loschmidt_cost(A0, tens(thetas[k & 15]), W)
    if kind == "energy_d8":
        H = O.heisenberg_matrix()
        thetas = rng.normal(size=(16, 24))
        return lambda k: O.energy_transfer(O.unitary_to_tensor(O.shallow_cnot_state_tensor_nonuniform(8, thetas[k & 15])), H)
    if kind in ("power_d64", "power_d256"):
        D = 64 if kind.endswith("64") else 256
        A = O.unitary_to_tensor(unitary_group.rvs(2 * D, random_state=4))
        Bt = O.unitary_to_tensor(unitary_group.rvs(2 * D, random_state=5))
        state = {"r": np.eye(D, dtype=complex) / np.sqrt(D)}

        def one(k):
            r = sum(A[s] @ state["r"] @ Bt[s].conj().T for s in range(A.shape[0]))
            state["r"] = r / np.linalg.norm(r)
        return one
    if kind == "brickwall_cost":
        U1, U2 = unitary_group.rvs(4, random_state=1), unitary_group.rvs(4, random_state=2)
        Vs = [(unitary_group.rvs(4, random_state=10 + k), unitary_group.rvs(4, random_state=50 + k)) for k in range(16)]
        h = rng.normal(size=(16, 16))
        W = expm(-0.1j * (h + h.T))
        return lambda k: OB.bw_exact_cost(U1, U2, Vs[k & 15][0], Vs[k & 15][1], W)
    raise ValueError(kind)


def _worker(args):
    kind, seconds = args
    os.environ["OMP_NUM_THREADS"] = "1"
    fn = _setup(kind)
    fn(0)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        fn(n)
        n += 1
    return n, time.perf_counter() - t0


UNITS = {
    "loschmidt_d2": ("Loschmidt / TDVP-step cost, D=2 (cfg 1/7 unit)", "steps/s"),
    "loschmidt_d4": ("Loschmidt / TDVP-step cost, D=4 (cfg 3 unit)", "steps/s"),
    "energy_d8": ("energy evaluation, D=8 Heisenberg ansatz (cfg 4 unit)", "evals/s"),
    "power_d64": ("transfer-matrix application, D=64 (cfg 5 unit)", "applications/s"),
    "power_d256": ("transfer-matrix application, D=256 (cfg 5 unit)", "applications/s"),
    "brickwall_cost": ("brick-wall Evolve.exact_cost_function (cfg 6 unit)", "costs/s"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--cores", type=int, default=0)
    args = ap.parse_args()
    cores = args.cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for kind, (what, unit) in UNITS.items():
            res = pool.map(_worker, [(kind, args.seconds)] * cores)
            rate = sum(n / dt for n, dt in res)
            print(json.dumps({"unit_of_work": what, "value": rate, "unit": unit, "cores": cores, "kind": "port",
                              "per_core": rate / cores,
                              "sample": f"{args.seconds:.0f} s of per-call oracle evaluations on each of {cores} cores (OMP_NUM_THREADS=1)"}), flush=True)


if __name__ == "__main__":
    main()
