#!/usr/bin/env python
"""Tiny invocation of every kernel family, the target of compute-sanitizer runs:

  compute-sanitizer --tool memcheck  python tools/sanitize_driver.py
  compute-sanitizer --tool racecheck python tools/sanitize_driver.py

Batches are a few dozen problems (odd sizes: ragged tails, idle sub-warp groups)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def reduced():
    """The shared-memory group kernels only (the ones with intra-group barriers), three problems each: a racecheck
    run of this finishes in minutes.  `python tools/sanitize_driver.py reduced`"""
    import torch
    from qmps_b200 import batched as B, brickwall as BW
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    rng = np.random.default_rng(0)

    def haar(n, m):
        Z = rng.normal(size=(n, m, m)) + 1j * rng.normal(size=(n, m, m))
        return np.ascontiguousarray(np.linalg.qr(Z)[0])

    def tensors(n, D):
        U = haar(n, 2 * D)
        return torch.from_numpy(np.ascontiguousarray(U[:, :, :D].reshape(n, D, 2, D).transpose(0, 2, 1, 3))).to(dev)
    for D in (4, 8):
        A, Bt = tensors(3, D), tensors(3, D)
        B.fixed_point(A, Bt)                                 # Hessenberg + QR + inverse iteration, aliased scratch
        B.fixed_point(A, Bt, left=True, want_vec=False)
        B.env_exact(A=A * 0.9, assume_left_canonical=False)
    A = tensors(3, 4)
    m = B.mixed_canonical(A * 0.8)
    B.expectation_values(m.AL, np.stack([np.diag([1.0, -1.0])]))
    z = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)                      # noqa: E731
    V1, V2 = haar(3, 4), haar(3, 4)
    BW.bw_evolve_cost(z(V1), z(V2), z(V1[::-1].copy()), z(V2[::-1].copy()), z(haar(1, 16)[0]), want_all=True)   # group kernel
    BW.bw_expectation(z(V1), z(V2), z(haar(1, 16)[0]))
    # round 2: the D = 8 tensor-pipe elimination (panel / update hand-over through shared memory), the PXP scar cost
    from qmps_b200 import represent as R, scars as SC
    from qmps_b200.ground_state import Hamiltonian
    B.env_exact(A=tensors(3, 8))                             # env_dmma_kernel<0, 2>
    prog = R.ShallowCNOTStateTensor_nonuniform(8, np.zeros(24)).program()
    th = torch.from_numpy(rng.normal(size=(3, 24))).to(dev)
    B.energy_theta(prog, th, Hamiltonian({'ZZ': -1, 'X': 0.7}).to_matrix(), coord=1, shifts=B.ROTO3_SHIFTS)   # ansatz + env_dmma_kernel<1, 2>
    SC.scars_costs(rng.normal(size=(5, 4)), rng.normal(size=4), SC.W(0.325, 0.1))
    torch.cuda.synchronize()
    print("sanitize_driver (reduced) done")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "reduced":
        return reduced()
    import torch
    from scipy.linalg import expm
    from qmps_b200 import batched as B, represent as R, brickwall as BW
    from qmps_b200.ground_state import Hamiltonian
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    rng = np.random.default_rng(0)

    def haar(n, m):
        Z = rng.normal(size=(n, m, m)) + 1j * rng.normal(size=(n, m, m))
        return np.ascontiguousarray(np.linalg.qr(Z)[0])

    def tensors(n, D):
        U = haar(n, 2 * D)
        return torch.from_numpy(np.ascontiguousarray(U[:, :, :D].reshape(n, D, 2, D).transpose(0, 2, 1, 3))).to(dev)

    H = Hamiltonian({'ZZ': -1, 'X': 0.7}).to_matrix()
    for cdt in (torch.complex128, torch.complex64):
        for D, n in ((2, 37), (4, 21), (8, 5), (16, 2)):
            A, Bt = tensors(n, D).to(cdt), tensors(n, D).to(cdt)
            B.env_exact(A=A)
            B.env_exact(A=A * 0.9, assume_left_canonical=False)
            B.fixed_point(A, Bt)
            B.fixed_point(A, Bt, left=True, want_vec=False)
            B.energy_tensor(A, H)
            if D <= 8:
                m = B.mixed_canonical(A * 0.8)
                B.expectation_values(m.AL, np.stack([np.diag([1.0, -1.0]), np.array([[0, 1.0], [1.0, 0]])]))
        for D, P, gate in ((2, 15, R.ShallowFullStateTensor), (4, 12, R.ShallowCNOTStateTensor_nonuniform),
                           (8, 24, R.ShallowCNOTStateTensor_nonuniform)):
            prog = gate(D, np.zeros(P)).program()
            th = torch.from_numpy(rng.normal(size=(19, P))).to(dev)
            e = B.energy_theta(prog, th, H, coord=1, shifts=B.ROTO3_SHIFTS, dtype=cdt)
            B.rotosolve_fit(e.double(), th, 1)
            A0 = B.ansatz_tensors(prog, th[:1], dtype=cdt)[0]
            W = torch.from_numpy(np.stack([expm(-1j * H * 0.05 * k) for k in range(5)])).to(dev).to(cdt)
            c = B.loschmidt_costs(prog, th, A0, W, dtype=cdt)[0]
            B.argmin(c.double().reshape(-1))
        U1, U2 = haar(1, 4)[0], haar(1, 4)[0]
        V1, V2 = haar(23, 4), haar(23, 4)
        Wb = haar(1, 16)[0]
        z = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev).to(cdt)          # noqa: E731
        BW.bw_evolve_cost(z(U1), z(U2), z(V1), z(V2), z(Wb), want_all=True)              # thread per candidate
        BW.bw_evolve_cost(z(V1), z(V2), z(V1), z(V2), z(Wb), want_all=True)              # group kernel (NK = N)
        BW.bw_environment(z(U1), z(U2), z(V1), z(V2), side="left", bra_undaggered=True)
        BW.bw_expectation(z(V1), z(V2), z(Wb))
        BW.bw_env_apply(z(U1), z(U2), z(V1), z(V2), z(haar(23, 2)))
        for D, n in ((64, 3), (100, 2)):
            A, Bt = tensors(n, D).to(cdt), tensors(n, D).to(cdt)
            B.tm_power(A, Bt, 2)
    B.loschmidt_rate(np.linspace(0.1, 2.0, 9), 1.5, 0.2)
    # round 2 entries
    from qmps_b200 import scars as SC, tools as T
    from scipy.stats import unitary_group
    A2 = tensors(37, 2)
    B.unpack_env(B.env_exact_packed(A=A2).cpu().numpy())
    B.env_exact_packed_host(A2.cpu().numpy())
    T.get_env_exact(unitary_group.rvs(4, random_state=3))
    SC.scars_costs(rng.normal(size=(21, 4)), rng.normal(size=4), SC.W(0.325, 0.1))
    SC.simulate_scars(0.05, 2, 0.325, np.array([0.9, 0.4, 0.7, 1.1]), n_gen=2, npop=64, n_bfgs=2)
    hh = torch.from_numpy(H).to(dev)
    for D, n in ((2, 5), (4, 3)):
        At = tensors(n, D)
        B.tdvp_dadt(At, hh)
        B.tdvp_evolve(At, hh, 0.01, 2)
    prog = R.ShallowFullStateTensor(2, np.zeros(15)).program()
    W = torch.from_numpy(expm(-1j * H * 0.1)).to(dev)
    B.loschmidt_trajectory(prog, torch.from_numpy(rng.normal(size=15)).to(dev), W, 2, n_gen=2, npop=64, n_bfgs=2)
    B.zgemm_i8(rng.normal(size=(2, 64, 128)) + 1j * rng.normal(size=(2, 64, 128)), rng.normal(size=(2, 32, 128)) + 1j * rng.normal(size=(2, 32, 128)))
    torch.cuda.synchronize()
    print("sanitize_driver done")


if __name__ == "__main__":
    main()
