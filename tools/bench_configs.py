#!/usr/bin/env python
"""Device-time throughput of BASELINE.json configs 2..5 at their full sizes (one GPU); 6 = brick-wall cost.

Not the driver's bench (that is ../bench.py, config 2): this is the measurement tool
behind DESIGN.md's per-kernel roofline table.  CUDA events on the launching stream,
3 warm-ups, inputs resident in HBM.  Prints one JSON object per config.

  python tools/bench_configs.py [--cfg 2,3,4,5] [--scale 1.0] [--reps 3]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(torch, fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(min(ts))


def left_canonical(torch, n, D, seed, dev, dtype):
    g = torch.Generator(device=dev).manual_seed(seed)
    rd = torch.float64
    Z = torch.randn((n, 2 * D, D), dtype=rd, device=dev, generator=g) + 1j * torch.randn((n, 2 * D, D), dtype=rd, device=dev, generator=g)
    Q, _ = torch.linalg.qr(Z)
    return Q.reshape(n, D, 2, D).permute(0, 2, 1, 3).contiguous().to(dtype)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="2,3,4,5")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the batch (smoke runs)")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--c64", action="store_true", help="also time the complex64 mode")
    args = ap.parse_args()
    import torch
    from scipy.linalg import expm
    from qmps_b200 import batched as B, represent as R
    from qmps_b200.ground_state import Hamiltonian
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    cfgs = [int(c) for c in args.cfg.split(",")]
    dtypes = [torch.complex128] + ([torch.complex64] if args.c64 else [])
    out = []
    for cdt in dtypes:
        tag = "c128" if cdt == torch.complex128 else "c64"
        if 2 in cfgs:
            N = int((1 << 20) * args.scale)
            A = left_canonical(torch, N, 2, 1, dev, cdt)
            med, best = timed(torch, lambda: B.env_exact(A=A, want_C=False, want_status=False), max(args.reps, 10), 3)
            out.append({"cfg": 2, "dtype": tag, "what": "exact env D=2", "N": N, "ms": med, "ms_best": best, "solves_per_s": N / med * 1e3})
        if 3 in cfgs:
            NP, NT = int(4096 * args.scale), 1000
            rng = np.random.default_rng(2)
            theta = torch.from_numpy(rng.normal(size=(NP, 12))).to(dev)
            gate = R.ShallowCNOTStateTensor_nonuniform(4, np.zeros(12))
            prog = gate.program()
            A0 = B.ansatz_tensors(prog, theta[:1], dtype=cdt)[0]
            W = torch.from_numpy(np.stack([expm(-1j * Hamiltonian({'ZZ': -1, 'X': 0.2}).to_matrix() * 2 * 0.02 * k) for k in range(NT)])).to(dev).to(cdt)
            med, best = timed(torch, lambda: B.loschmidt_costs(prog, theta, A0, W, dtype=cdt), args.reps, 1)
            out.append({"cfg": 3, "dtype": tag, "what": "Loschmidt D=4 (16x16 mixed two-site map, all eigenvalues)", "NP": NP, "NT": NT,
                        "ms": med, "ms_best": best, "steps_per_s": NP * NT / med * 1e3})
        if 4 in cfgs:
            N = int(65536 * args.scale)
            rng = np.random.default_rng(3)
            theta = torch.from_numpy(rng.normal(size=(N, 24))).to(dev)
            gate = R.ShallowCNOTStateTensor_nonuniform(8, np.zeros(24))
            prog = gate.program()
            H = Hamiltonian({'XX': 1, 'YY': 1, 'ZZ': 1}).to_matrix()
            med, best = timed(torch, lambda: B.energy_theta(prog, theta, H, coord=5, shifts=B.ROTO3_SHIFTS, dtype=cdt), args.reps, 1)
            flops = 8 * 2 * 8 ** 4 + (8.0 / 3.0) * 8 ** 6
            out.append({"cfg": 4, "dtype": tag, "what": "rotosolve 3-shift energy D=8 (64x64 direct solve)", "N": N, "evals": 3 * N,
                        "ms": med, "ms_best": best, "evals_per_s": 3 * N / med * 1e3, "algo_tflops": 3 * N * flops / med * 1e3 / 1e12})
        if 5 in cfgs:
            for D, N in ((64, int(512 * args.scale)), (256, max(1, int(32 * args.scale)))):
                A = left_canonical(torch, N, D, 4, dev, cdt)
                Bt = left_canonical(torch, N, D, 5, dev, cdt)
                K = 32
                med, best = timed(torch, lambda: B.tm_power(A, Bt, K), args.reps, 1)
                apps = N * (K + 1)
                out.append({"cfg": 5, "dtype": tag, "what": f"power method D={D}", "N": N, "K": K, "ms": med, "ms_best": best,
                            "applications_per_s": apps / med * 1e3, "algo_tflops": apps * 32.0 * D ** 3 / med * 1e3 / 1e12})
        if 6 in cfgs:                      # brick-wall TDVP step cost, one ket state, N candidates
            from qmps_b200 import brickwall as BW
            N = int((1 << 20) * args.scale)
            rng6 = np.random.default_rng(6)

            def haar(n):                   # numpy QR on the host (set-up only)
                Z = rng6.normal(size=(n, 4, 4)) + 1j * rng6.normal(size=(n, 4, 4))
                return torch.from_numpy(np.ascontiguousarray(np.linalg.qr(Z)[0])).to(dev).to(cdt).contiguous()
            U1, U2, V1, V2 = haar(1), haar(1), haar(N), haar(N)
            h = np.random.default_rng(6).normal(size=(16, 16))
            W = torch.from_numpy(expm(-0.1j * (h + h.T))).to(dev).to(cdt)
            from qmps_b200 import _lib as L6
            for flag in (1, 0):
                L6.load().qmps_set_option(b"bw_thread", flag)
                med, best = timed(torch, lambda: BW.bw_evolve_cost(U1, U2, V1, V2, W), args.reps, 2)
                # traffic: 2 x 16 complex in, one real out per candidate
                out.append({"cfg": 6, "dtype": tag, "what": "brick-wall Evolve.exact_cost_function (4x4 env eig + 6-qubit overlap)", "N": N,
                            "kernel": "bw_cost_thread_kernel (thread per candidate)" if flag else "bw_kernel<T,16> (16 lanes per candidate)",
                            "ms": med, "ms_best": best, "costs_per_s": N / med * 1e3,
                            "algo_gbs": N * (2 * 16 * (16 if cdt == torch.complex128 else 8) + (8 if cdt == torch.complex128 else 4)) / med * 1e3 / 1e9})
            L6.load().qmps_set_option(b"bw_thread", 1)
        if 7 in cfgs:                      # the metric's "Loschmidt-echo steps/sec at D=2": cfg 3's grid with the reference's
            NP, NT = int(4096 * args.scale), 1000     # own D = 2 ansatz (ShallowFullStateTensor, 15 parameters, represent.py:392-401)
            rng = np.random.default_rng(7)
            theta = torch.from_numpy(rng.normal(size=(NP, 15))).to(dev)
            prog = R.ShallowFullStateTensor(2, np.zeros(15)).program()
            A0 = B.ansatz_tensors(prog, theta[:1], dtype=cdt)[0]
            W = torch.from_numpy(np.stack([expm(-1j * Hamiltonian({'ZZ': -1, 'X': 0.2}).to_matrix() * 2 * 0.02 * k) for k in range(NT)])).to(dev).to(cdt)
            from qmps_b200 import _lib as L_
            for flag in (1, 0):
                L_.load().qmps_set_option(b"fp_d2", flag)
                med, best = timed(torch, lambda: B.loschmidt_costs(prog, theta, A0, W, dtype=cdt), args.reps, 1)
                out.append({"cfg": 7, "dtype": tag, "what": "Loschmidt D=2 (4x4 mixed two-site map, all eigenvalues)", "NP": NP, "NT": NT,
                            "kernel": "fp_d2_kernel (thread per problem)" if flag else "fixed_point_kernel<T,4> (generic)",
                            "ms": med, "ms_best": best, "steps_per_s": NP * NT / med * 1e3})
            L_.load().qmps_set_option(b"fp_d2", 1)
        if 8 in cfgs:                      # D = 2 rotosolve: fused 3-shift energies for the reference's 15-parameter ansatz
            N = int((1 << 20) * args.scale)
            rng = np.random.default_rng(8)
            theta = torch.from_numpy(rng.normal(size=(N, 15))).to(dev)
            prog = R.ShallowFullStateTensor(2, np.zeros(15)).program()
            H = Hamiltonian({'ZZ': -1, 'X': 1.0}).to_matrix()
            med, best = timed(torch, lambda: B.energy_theta(prog, theta, H, coord=7, shifts=B.ROTO3_SHIFTS, dtype=cdt), args.reps, 2)
            out.append({"cfg": 8, "dtype": tag, "what": "rotosolve 3-shift energy D=2 (theta -> U -> A -> env -> energy, one launch)", "N": N,
                        "evals": 3 * N, "ms": med, "ms_best": best, "evals_per_s": 3 * N / med * 1e3})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
