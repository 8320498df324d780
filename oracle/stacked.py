"""Stacked (vectorised) numpy forms of the per-call oracle functions.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Two uses:

* the checker for FULL-SIZE batches (2^20 seeded D = 2 solves, the 4096 x 1000 Loschmidt grid) that
  the per-call oracle would need minutes for -- every function here is asserted equal to its
  per-call counterpart in ``tests/test_oracle.py`` first;
* BASELINE.md section 3's variant **B2 "fair vectorised CPU"**: the same algorithm as the reference's
  per-call path (dense ``eig`` of the transfer matrix, argmax |lambda|, Hermitian rotation,
  ``cholesky``), but issued as stacked ``numpy.linalg`` calls over the whole batch -- what a
  performance-minded maintainer of the reference would write without leaving numpy.

Each function names the reference lines its per-call counterpart follows.
"""
import numpy as np

__all__ = ["haar_unitaries", "tensors_of_unitaries", "stacked_transfer_matrices", "stacked_leading_eig",
           "stacked_env_exact", "stacked_cholesky_env", "stacked_merge", "stacked_loschmidt_costs",
           "stacked_energy_transfer", "parallel_map_chunks"]


def haar_unitaries(n, count, seed):
    """``scipy.stats.unitary_group.rvs(n, size=count, random_state=seed)`` -- the seeded synthetic input
    SURVEY 8(d) names (cfg 2: n = 4, count = 2^20, seed = 1)."""
    from scipy.stats import unitary_group
    return unitary_group.rvs(n, size=count, random_state=seed).reshape(count, n, n)


def tensors_of_unitaries(U):
    """stacked ``unitary_to_tensor`` (qmps/tools.py:151-154): U[N,2D,2D] -> A[N,2,D,D]."""
    N, m, _ = U.shape
    D = m // 2
    return np.ascontiguousarray(U.reshape(N, D, 2, 2, D)[:, :, :, 0, :].transpose(0, 2, 1, 3))


def stacked_transfer_matrices(A, B=None):
    """E[n,(i,k),(j,l)] = sum_s A[n,s,i,j] conj(B[n,s,k,l]) (new_tdvp/EnvironmentParamSensitivity.py:37-38)."""
    B = A if B is None else B
    N, _, D1, _ = A.shape
    D2 = B.shape[2]
    return np.einsum("nsij,nskl->nikjl", A, B.conj()).reshape(N, D1 * D2, D1 * D2)


def stacked_leading_eig(E):
    """stacked ``leading_eig``: dense eig of every matrix, eigenpair of largest modulus."""
    w, v = np.linalg.eig(E)
    k = np.argmax(np.abs(w), axis=1)
    idx = np.arange(E.shape[0])
    return w[idx, k], v[idx, :, k]


def stacked_env_exact(A):
    """stacked ``eigs`` -> (eta[N], r[N,D,D]) with r Hermitian, trace 1
    (``TransferMatrix(A).eigs()`` as qmps/tools.py:181-182 consumes it)."""
    N, _, D, _ = A.shape
    eta, v = stacked_leading_eig(stacked_transfer_matrices(A))
    x = v.reshape(N, D, D)
    t = np.einsum("nii->n", x)
    x = x * (np.conj(t) / np.abs(t))[:, None, None]
    r = (x + x.conj().transpose(0, 2, 1)) / 2
    r = r / np.einsum("nii->n", r).real[:, None, None]
    return eta, r


def stacked_cholesky_env(r):
    """stacked ``cholesky(r).conj().T`` (qmps/tools.py:182) -> lower C with r = C C^dagger, and the
    unique first column vec(C)/|C|_F of ``environment_to_unitary`` (qmps/tools.py:106)."""
    C = np.linalg.cholesky(r)
    v0 = C.reshape(C.shape[0], -1)
    return C, v0 / np.linalg.norm(v0, axis=1, keepdims=True)


def stacked_merge(A, B):
    """stacked ``merge`` (qmps/time_evolve_tools.py:20-23, any bond dimension)."""
    N, d1, D, _ = A.shape
    d2 = B.shape[1]
    return np.einsum("naik,nbkj->nabij", A, B).reshape(N, d1 * d2, D, D)


def stacked_loschmidt_costs(A0, Bs, Ws):
    """cost[p, k] = -sqrt|eta_2| with eta_2 the leading eigenvalue of
    ``Map(W_k . merge(A0,A0), merge(B_p,B_p))`` (qmps/loschmidts/time_evo.py:75-116) for every
    parameter set p (tensors Bs[NP,2,D,D]) and time k (gates Ws[NT,4,4]).  Returns (cost, echo)."""
    MA = stacked_merge(A0[None], A0[None])[0]                    # [4, D, D]
    WMA = np.einsum("kab,bij->kaij", Ws, MA)                     # [NT, 4, D, D]
    MB = stacked_merge(Bs, Bs)                                   # [NP, 4, D, D]
    NP, NT, D = Bs.shape[0], Ws.shape[0], A0.shape[1]
    E = np.einsum("kaij,pars->pkirjs", WMA, MB.conj()).reshape(NP * NT, D * D, D * D)
    w = np.linalg.eigvals(E)
    a = np.abs(w).max(axis=1).reshape(NP, NT)
    return -np.sqrt(a), -np.log(a * a)


def stacked_energy_transfer(A, H):
    """stacked ``energy_transfer`` (qmps/ground_state.py:251-266 / scripts/ground_state_finding.py:119-128
    as a transfer-matrix expression): e[n] = Re sum_ab H[a,b] tr(M_a^dagger M_b r)."""
    _, r = stacked_env_exact(A)
    M = stacked_merge(A, A)
    return np.einsum("ab,naji,nbjk,nki->n", H, M.conj(), M, r).real


def _call(args):
    fn, chunk = args
    return fn(*chunk)


def parallel_map_chunks(fn, arrays, nproc=None, chunk=8192):
    """Run ``fn(*[a[i:j] for a in arrays])`` over row chunks on a fork pool with one BLAS thread per
    process; returns the per-chunk results in order.  ``fn`` must be a module-level function."""
    import multiprocessing as mp
    import os
    nproc = nproc or os.cpu_count() or 1
    N = arrays[0].shape[0]
    jobs = [(fn, [a[i:i + chunk] for a in arrays]) for i in range(0, N, chunk)]
    if nproc == 1 or len(jobs) == 1:
        return [_call(j) for j in jobs]
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    with mp.get_context("fork").Pool(min(nproc, len(jobs))) as pool:
        return pool.map(_call, jobs)
