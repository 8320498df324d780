"""Classical iTDVP for a single-site uniform MPS.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  **Parity unpinned**: the reference delegates all of
this to the un-vendored ``xmps`` package -- ``iMPS.dA_dt([H])`` (scripts/classical_time_evolution.py:22-26,
scripts/mixed_environment.py:41, Time Evo.ipynb), ``xmps.iTDVP.Trajectory(mps_0, H).eulerint(T).loschmidts()``
(qmps/loschmidts/mps_loschmidts.py:20-22).  What is restated here is the published algorithm those names refer to
(Haegeman et al., PRL 107, 070601 (2011), eqs. for the gauge-fixed tangent vector of a uniform MPS with a
nearest-neighbour Hamiltonian), anchored on what the reference's call sites and plots force:

* the drivers are exactly the reference's: classical RK4 with ``left_canonicalise`` after each step
  (classical_time_evolution.py:21-27) and explicit Euler (mps_loschmidts.py:22);
* ``loschmidts()`` is plotted against the analytic TFIM rate ``loschmidts(T, g0, g1)`` of
  qmps/loschmidts/exact_loschmidt.py (mps_loschmidts.py:25-26), so it is the per-site rate
  -log|eta(E_{A_t A_0})|^2 -- and the trajectory must reproduce that analytic curve at short times
  (tests/test_oracle_tdvp.py);
* energy and norm are conserved by the flow, single-site Hamiltonians are integrated exactly.

Left-canonical gauge (l = 1, r = trace-1 right fixed point), C^{st} = sum h[(s,t),(s',t')] A_s' A_t':

    H_l = sum_st (A_s A_t)^dagger C^{st},   e = tr(H_l r)                      (energy per site)
    K - sum_s A_s^dagger K A_s = H_l - e 1,  tr(K r) = 0                        (everything to the left)
    G^s = sum_t C^{st} r A_t^dagger r^{-1} + sum_t A_t^dagger C^{ts} + K A_s
    dA^s/dt = -i (G^s - A_s sum_u A_u^dagger G^u)                               (project out the gauge part)

For a tensor in any gauge: canonicalise (A_L = L A L^{-1}/sqrt(eta)), take the tangent there and transform it
back, dA = sqrt(eta) L^{-1} dA_L L -- the left gauge condition is covariant, so this is THE tangent vector.
"""
import numpy as np
import scipy.linalg as sla

from .tensors import transfer_matrix, leading_eig, _hermitian_gauge, eigs, right_fixed_point

__all__ = ["tdvp_canonical_parts", "tdvp_tangent_left_canonical", "dA_dt", "tdvp_euler_step", "tdvp_rk4_step",
           "tdvp_trajectory", "loschmidt_rates", "energy_density"]


def tdvp_canonical_parts(A):
    """(A_L, eta, L): A_L = L A L^{-1} / sqrt|eta| left-canonical, L upper triangular with L^dagger L = l D / tr l
    (the Cholesky gauge of ``oracle.left_canonicalise``)."""
    D = A.shape[1]
    eta, v = leading_eig(transfer_matrix(A).conj().T)
    l = _hermitian_gauge(v.reshape(D, D))
    l = l / np.trace(l).real * D
    L = sla.cholesky(l)
    AL = np.einsum("ab,sbc,cd->sad", L, A, np.linalg.inv(L)) / np.sqrt(abs(eta))
    return AL, eta, L


def energy_density(AL, h, r=None):
    """e = sum_st tr((A_s A_t)^dagger C^{st} r) for a left-canonical tensor: <h> per bond."""
    d, D, _ = AL.shape
    if r is None:
        _, _, r = eigs(AL)
    AA = np.einsum("sij,tjk->stik", AL, AL)
    C = np.einsum("abcd,cdik->abik", np.asarray(h, dtype=complex).reshape(d, d, d, d), AA)
    return float(np.einsum("stji,stjk,ki->", AA.conj(), C, r).real)


def tdvp_tangent_left_canonical(AL, h, imaginary=False, r=None):
    """(dA_L/dt, e) for a LEFT-CANONICAL tensor A_L[d, D, D] and a two-site Hamiltonian h[d*d, d*d];
    ``imaginary``: the imaginary-time flow (-1 instead of -i).  ``r``: the Hermitian trace-1 right fixed point if the
    caller already has it (large D: a dense D^2 x D^2 eig is slow, an iterative eigen-solver supplies it)."""
    d, D, _ = AL.shape
    if r is None:
        _, _, r = eigs(AL)                                         # Hermitian, trace 1
    AA = np.einsum("sij,tjk->stik", AL, AL)
    C = np.einsum("abcd,cdik->abik", np.asarray(h, dtype=complex).reshape(d, d, d, d), AA)
    Hl = np.einsum("stji,stjk->ik", AA.conj(), C)
    e = np.einsum("ik,ki->", Hl, r).real
    # (1 - E_left) K = Hl - e, with the redundant first equation replaced by tr(K r) = 0
    EL = np.einsum("sji,slk->ikjl", AL.conj(), AL).reshape(D * D, D * D)       # (E_left K)_ik = sum conj(A[s,j,i]) K[j,l] A[s,l,k]
    M = np.eye(D * D, dtype=complex) - EL
    rhs = (Hl - e * np.eye(D)).reshape(-1)
    M[0, :] = r.T.reshape(-1)
    rhs[0] = 0.0
    K = np.linalg.solve(M, rhs).reshape(D, D)
    rinv = np.linalg.inv(r)
    G = (np.einsum("stik,kl,tml,mj->sij", C, r, AL.conj(), rinv)               # h on (n, n+1), tangent on n
         + np.einsum("tki,tskj->sij", AL.conj(), C)                           # h on (n-1, n), tangent on n
         + np.einsum("ik,skj->sij", K, AL))                                   # h further left
    P = np.einsum("ski,skj->ij", AL.conj(), G)
    dA = (-1.0 if imaginary else -1j) * (G - np.einsum("sik,kj->sij", AL, P))
    return dA, float(e)


def dA_dt(A, h, imaginary=False):
    """``iMPS([A]).dA_dt([h])``: the TDVP tangent vector of a uniform tensor in ANY gauge."""
    AL, eta, L = tdvp_canonical_parts(A)
    dAL, e = tdvp_tangent_left_canonical(AL, h, imaginary)
    Li = np.linalg.inv(L)
    return np.sqrt(abs(eta)) * np.einsum("ab,sbc,cd->sad", Li, dAL, L), e


def _canon(A):
    return tdvp_canonical_parts(A)[0]


def tdvp_euler_step(A, h, dt, imaginary=False):
    """``Trajectory.eulerint`` step (qmps/loschmidts/mps_loschmidts.py:22): A <- canon(A + dt dA/dt)."""
    return _canon(A + dt * dA_dt(A, h, imaginary)[0])


def tdvp_rk4_step(A, h, dt, imaginary=False):
    """The reference's own RK4 loop body (scripts/classical_time_evolution.py:22-26)."""
    k1 = dA_dt(A, h, imaginary)[0] * dt
    k2 = dA_dt(A + k1 / 2, h, imaginary)[0] * dt
    k3 = dA_dt(A + k2 / 2, h, imaginary)[0] * dt
    k4 = dA_dt(A + k3, h, imaginary)[0] * dt
    return _canon(A + (k1 + 2 * k2 + 2 * k3 + k4) / 6)


def tdvp_trajectory(A0, h, dt, n_steps, method="rk4", imaginary=False):
    """[A_0, A_1, ..., A_n] (each left-canonical)."""
    step = tdvp_rk4_step if method == "rk4" else tdvp_euler_step
    out = [_canon(A0)]
    for _ in range(n_steps):
        out.append(step(out[-1], h, dt, imaginary))
    return out


def loschmidt_rates(traj, A_ref=None):
    """``Trajectory.loschmidts()``: -log|eta(E_{A_t, A_ref})|^2 per site (A_ref = A_0 by default)."""
    A_ref = traj[0] if A_ref is None else A_ref
    return np.array([-np.log(abs(right_fixed_point(A, A_ref)[0]) ** 2) for A in traj])
