#!/usr/bin/env python
"""D = 8 Loschmidt tile: generic fixed_point_kernel<T,128> (option fp64_fast = 0) vs the warp-per-problem kernel
(kernels_fp64w.cuh, fp64_fast = 1): agreement, sweeps per problem, steps/s; plus Haar-random pairs (right / left,
d = 2 and 4) against numpy.linalg.eigvals."""
import ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from scipy.linalg import expm
    from qmps_b200 import _lib as L, batched as B, represent as R
    from qmps_b200.ground_state import Hamiltonian
    lib = L.require_device()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    rng = np.random.default_rng(8)
    # --- random pairs against numpy
    for d in (2, 4):
        for left in (False, True):
            N = 64
            Z = rng.normal(size=(2, N, d * 8, 8)) + 1j * rng.normal(size=(2, N, d * 8, 8))
            Q = np.linalg.qr(Z)[0].reshape(2, N, d, 8, 8)
            A, Bt = Q[0], Q[1]
            ref = []
            for n in range(N):
                E = sum(np.kron(A[n, s], Bt[n, s].conj()) for s in range(d))
                w = np.linalg.eigvals(E)
                ref.append(np.abs(w).max())
            ref = np.array(ref)
            out = {}
            for fast in (0, 1, 2):
                lib.qmps_set_option(b"fp64_fast", fast)
                res = B.fixed_point(torch.from_numpy(A).to(dev), torch.from_numpy(Bt).to(dev), left=left, want_vec=False)
                eta = res.eta
                out[fast] = eta.cpu().numpy()
            print(json.dumps({"check": "haar_pairs", "d": d, "left": left,
                              "max_rel_abs_eta_vs_numpy": float(np.max(np.abs(np.abs(out[1]) - ref) / ref)),
                              "generic_vs_numpy": float(np.max(np.abs(np.abs(out[0]) - ref) / ref)),
                              "max_abs_eta_fast_vs_generic": float(np.max(np.abs(out[1] - out[0]))), "max_abs_eta_packed_vs_generic": float(np.max(np.abs(out[2] - out[0])))}), flush=True)
    # --- the bench tile
    NP, NT = int(os.environ.get("NP", 256)), int(os.environ.get("NT", 100))
    theta = torch.from_numpy(np.random.default_rng(8).normal(size=(NP, 24))).to(dev)
    prog = R.ShallowCNOTStateTensor_nonuniform(8, np.zeros(24)).program()
    H = Hamiltonian({'ZZ': -1, 'X': 0.2}).to_matrix()
    Wn = np.stack([expm(-1j * H * 0.02 * k * (1000 // NT)) for k in range(NT)])
    cnt = (ctypes.c_ulonglong * 4)()
    ref = None
    for cdt, tag in ((torch.complex128, "c128"), (torch.complex64, "c64")):
        A0 = B.ansatz_tensors(prog, theta[:1], dtype=cdt)[0]
        W = torch.from_numpy(Wn).to(dev).to(cdt)
        for fast in (0, 1, 2):
            lib.qmps_set_option(b"fp64_fast", fast)
            fn = lambda: B.loschmidt_costs(prog, theta, A0, W, dtype=cdt)
            lib.qmps_debug_counters(cnt, 1)
            c = fn()[0]
            torch.cuda.synchronize()
            lib.qmps_debug_counters(cnt, 1)
            if ref is None:
                ref = c.double().clone()
            err = float((c.double() - ref).abs().max())
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            print(json.dumps({"dtype": tag, "fp64_fast": fast, "ms": round(ms, 3), "steps_per_s": NP * NT / ms * 1e3,
                              "max_abs_diff_vs_generic_c128": err,
                              "sweeps_per_problem": cnt[1] / max(cnt[0], 1), "forced": int(cnt[2])}), flush=True)
    lib.qmps_set_option(b"fp64_fast", 2)


if __name__ == "__main__":
    main()
