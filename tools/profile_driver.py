#!/usr/bin/env python
"""Launches each hot kernel a few times on cheap synthetic inputs -- the target of the
`ncu --set full -k regex:...` captures (tools/gpu_round.sh).  Inputs come from this repo's own
ansatz kernel (theta -> left-canonical A), so almost every launch ncu sees is one of ours.

  python tools/profile_driver.py [--what d2,fp4,en8,pw64,pw256] [--reps 3]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="d2,fp4,en8,pw64")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    from qmps_b200 import batched as B, represent as R
    from qmps_b200.ground_state import Hamiltonian
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    what = args.what.split(",")
    g = torch.Generator(device=dev).manual_seed(0)

    def thetas(n, p):
        return torch.randn((n, p), dtype=torch.float64, device=dev, generator=g)

    def lc_tensors(n, D, layers=3):
        nq = int(np.log2(D)) + 1
        prog = R.ShallowCNOTStateTensor_nonuniform(D, np.zeros(2 * nq * layers)).program()
        return B.ansatz_tensors(prog, thetas(n, 2 * nq * layers))

    if "d2" in what:                       # the bench kernel exactly as bench.py launches it: all four outputs, rotating buffer sets
        N = 1 << 20
        nb = 4
        As = [lc_tensors(N, 2) for _ in range(nb)]
        outs = [(torch.empty((N,), dtype=torch.complex128, device=dev), torch.empty((N, 2, 2), dtype=torch.complex128, device=dev),
                 torch.empty((N, 2, 2), dtype=torch.complex128, device=dev), torch.empty((N,), dtype=torch.int32, device=dev)) for _ in range(nb)]
        from qmps_b200 import _lib as L
        lib = L.load()
        for i in range(args.reps + 6):
            e, r, C, st = outs[i % nb]
            L.check(lib.qmps_env_exact(2, 2, N, As[i % nb].data_ptr(), 0, 1, e.data_ptr(), r.data_ptr(), C.data_ptr(), st.data_ptr(), L.C128,
                                       torch.cuda.current_stream().cuda_stream), "env_exact")
    if "fp4" in what:                      # cfg 3 tile: 16x16 mixed two-site maps
        NP, NT = 512, 64
        prog = R.ShallowCNOTStateTensor_nonuniform(4, np.zeros(12)).program()
        th = thetas(NP, 12)
        A0 = B.ansatz_tensors(prog, th[:1])[0]
        from scipy.linalg import expm
        H = Hamiltonian({'ZZ': -1, 'X': 0.2}).to_matrix()
        W = torch.from_numpy(np.stack([expm(-1j * H * 0.04 * k) for k in range(NT)])).to(dev)
        for _ in range(args.reps):
            B.loschmidt_costs(prog, th, A0, W)
    if "en8" in what:                      # cfg 4 tile: D=8 energy with 3 shifts
        prog = R.ShallowCNOTStateTensor_nonuniform(8, np.zeros(24)).program()
        th = thetas(8192, 24)
        H = Hamiltonian({'XX': 1, 'YY': 1, 'ZZ': 1}).to_matrix()
        for _ in range(args.reps):
            B.energy_theta(prog, th, H, coord=5, shifts=B.ROTO3_SHIFTS)
    for tag, D, N in (("pw64", 64, 512), ("pw256", 256, 32)):
        if tag in what:
            A = torch.view_as_complex(torch.randn((N, 2, D, D, 2), dtype=torch.float64, device=dev, generator=g)) / np.sqrt(2 * D)
            Bt = torch.view_as_complex(torch.randn((N, 2, D, D, 2), dtype=torch.float64, device=dev, generator=g)) / np.sqrt(2 * D)
            for _ in range(args.reps):
                B.tm_power(A, Bt, 2)
    for tag, D, N in (("tc64", 64, 512), ("tc256", 256, 32)):      # complex64: tcgen05 path
        if tag in what:
            A = torch.view_as_complex(torch.randn((N, 2, D, D, 2), dtype=torch.float32, device=dev, generator=g)) / np.sqrt(2 * D)
            Bt = torch.view_as_complex(torch.randn((N, 2, D, D, 2), dtype=torch.float32, device=dev, generator=g)) / np.sqrt(2 * D)
            for _ in range(args.reps):
                B.tm_power(A, Bt, 2)
    if "canon" in what:                   # gauge fixing + local expectation values
        for D in (2, 4):
            A = torch.view_as_complex(torch.randn((1 << 16, 2, D, D, 2), dtype=torch.float64, device=dev, generator=g))
            for _ in range(args.reps):
                m = B.mixed_canonical(A)
                B.expectation_values(m.AL, np.stack([np.array([[1, 0], [0, -1]]), np.array([[0, 1], [1, 0]])]))
    if "bw" in what:                      # brick-wall TDVP step cost
        from qmps_b200 import brickwall as BW

        rng = np.random.default_rng(6)

        def haar(n, m=4):                  # numpy QR on the host: torch's batched QR launches per matrix
            Z = rng.normal(size=(n, m, m)) + 1j * rng.normal(size=(n, m, m))
            return torch.from_numpy(np.ascontiguousarray(np.linalg.qr(Z)[0])).to(dev)
        U1, U2, V1, V2 = haar(1), haar(1), haar(1 << 18), haar(1 << 18)
        W = haar(1, 16)[0].contiguous()
        for _ in range(args.reps):
            BW.bw_evolve_cost(U1, U2, V1, V2, W)
    if "fpd2" in what:                    # the metric's D = 2 Loschmidt steps: thread-per-problem register kernel
        NP, NT = 4096, 256
        prog = R.ShallowFullStateTensor(2, np.zeros(15)).program()
        th = thetas(NP, 15)
        A0 = B.ansatz_tensors(prog, th[:1])[0]
        from scipy.linalg import expm
        H = Hamiltonian({'ZZ': -1, 'X': 0.2}).to_matrix()
        W = torch.from_numpy(np.stack([expm(-1j * H * 0.04 * k) for k in range(NT)])).to(dev)
        for _ in range(args.reps):
            B.loschmidt_costs(prog, th, A0, W)
    torch.cuda.synchronize()
    print("profile_driver done:", what)


if __name__ == "__main__":
    main()
