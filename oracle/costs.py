"""Oracle, part 3: Hamiltonians, energy / overlap / Loschmidt costs, rotosolve.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Each cost is given twice:
the reference's own route (state-vector simulation of its circuit, restated
gate by gate) and the transfer-matrix expression the CUDA kernels evaluate
(SURVEY A.4); the CPU tests assert that the two agree.
"""
import numpy as np
import scipy.linalg as sla
from scipy.integrate import quad
from scipy.optimize import minimize_scalar

from .tensors import (unitary_to_tensor, tensor_to_unitary, environment_to_unitary,
                      eigs, merge, right_fixed_point, left_fixed_point,
                      apply_two_site_gate, put_env_on_left_site, put_env_on_right_site)
from .gates import (I2, PX, PY, PZ, simulate, state_circuit, hadamard, cnot,
                    shallow_full_state_tensor)

__all__ = [
    "hamiltonian_strings", "hamiltonian_to_matrix", "tfim_matrix", "heisenberg_matrix",
    "energy_statevector", "energy_transfer", "energy_of_unitary",
    "energy_two_site_statevector", "energy_two_site_transfer",
    "loschmidt_cost_circuit", "loschmidt_cost", "get_overlap_exact",
    "rotosolve_theta3", "rotosolve_step3", "double_rotosolve_fit", "double_rotosolve",
    "wrap_angle", "exact_loschmidt_f", "exact_loschmidt", "exact_loschmidts",
    "tfim_e0_exact", "tfim_evolution_gate",
]

_S = {"I": I2, "X": PX, "Y": PY, "Z": PZ}


# ---- a10: Hamiltonian (qmps/ground_state.py:66-88) --------------------------
def hamiltonian_strings(strings):
    """Key expansion of ``Hamiltonian.__init__`` (ground_state.py:73-80): a
    single-letter key P with coefficient c becomes IP: c/2 and PI: c/2."""
    out = {}
    for key, val in strings.items():
        if len(key) == 1:
            out["I" + key] = out.get("I" + key, 0) + val / 2
            out[key + "I"] = out.get(key + "I", 0) + val / 2
        else:
            out[key] = out.get(key, 0) + val
    return out


def hamiltonian_to_matrix(strings):
    """``Hamiltonian(strings).to_matrix()`` (ground_state.py:82-88); Pauli sigma
    matrices (``paulis(0.5)``), checked against the literal TFIM matrix of
    tests/test_ground_state.py:29-38."""
    h = np.zeros((4, 4), dtype=np.complex128)
    for key, J in hamiltonian_strings(strings).items():
        h += J * np.kron(_S[key[0]], _S[key[1]])
    return h


def tfim_matrix(g, J=-1.0):
    return hamiltonian_to_matrix({"ZZ": J, "X": g})


def heisenberg_matrix():
    return hamiltonian_to_matrix({"XX": 1.0, "YY": 1.0, "ZZ": 1.0})


def tfim_evolution_gate(g, dt):
    """``expm(-1j*Hamiltonian({'ZZ':-1,'X':g}).to_matrix()*dt)``
    (scripts/loschmidt.py:341 uses 2*dt; qmps/new_time_evolve.py:240 uses dt)."""
    return sla.expm(-1j * tfim_matrix(g) * dt)


# ---- a9: energy -----------------------------------------------------------
def energy_statevector(U, V, H):
    """The reference's route (ground_state.py:251-266 / 150-168): simulate
    ``State(U, V, 2)`` on |0...0> and evaluate Re <psi| 1_D (x) H (x) 1_D |psi>."""
    D = U.shape[0] // 2
    ops, n = state_circuit(U, V, 2)
    psi = simulate(ops, n)
    op = np.kron(np.kron(np.eye(D), H), np.eye(D))
    return float(np.real(np.vdot(psi, op @ psi)))


def energy_transfer(A, H, r=None):
    """e = sum_ab H_ab tr(M_a^dagger M_b r), M = merge(A, A), r the trace-1 right
    fixed point of E_AA."""
    if r is None:
        _, _, r = eigs(A)
    M = merge(A, A)
    G = np.einsum("aji,bjk,ki->ab", M.conj(), M, r)       # tr(M_a^dagger M_b r)
    return float(np.real(np.sum(H * G)))


def energy_of_unitary(U, H):
    """Reference call chain get_env_exact -> State -> simulate (ground_state.py:251-266)."""
    from .tensors import get_env_exact
    return energy_statevector(U, get_env_exact(U), H)


def _two_site_env(U1, U2):
    """ground_state.py:291-297."""
    A12 = merge(unitary_to_tensor(U1), unitary_to_tensor(U2))
    _, _, r = eigs(A12)
    return environment_to_unitary(sla.cholesky(r).conj().T)


def energy_two_site_statevector(U1, U2, H):
    """ground_state.py:299-331 (D=2 only in the reference)."""
    out = []
    for Ua, Ub in ((U1, U2), (U2, U1)):
        V = _two_site_env(Ua, Ub)
        ops = [(V, (2, 3)), (Ub, (1, 2)), (Ua, (0, 1))]
        psi = simulate(ops, 4)
        op = np.kron(np.kron(np.eye(2), H), np.eye(2))
        out.append(np.real(np.vdot(psi, op @ psi)))
    return float(sum(out) / 2)


def energy_two_site_transfer(A1, A2, H):
    """Two-site unit cell: r1 from E of merge(A1,A2), r2 from merge(A2,A1)
    (SURVEY A.4)."""
    out = []
    for Aa, Ab in ((A1, A2), (A2, A1)):
        M = merge(Aa, Ab)
        _, _, r = eigs(M)
        G = np.einsum("aji,bjk,ki->ab", M.conj(), M, r)
        out.append(np.real(np.sum(H * G)))
    return float(sum(out) / 2)


# ---- a11: Loschmidt / TDVP-step cost ---------------------------------------
def loschmidt_cost_circuit(A, B, W):
    """The reference's 6-qubit circuit amplitude (loschmidts/time_evo.py:75-116 =
    scripts/loschmidt.py:209-239): returns -sqrt(2 |<0|C|0>|).  D=2 only."""
    E_A = apply_two_site_gate(W, merge(A, A))
    x, r = right_fixed_point(E_A, merge(B, B))
    l = r                                              # scripts/loschmidt.py:216
    U = tensor_to_unitary(A)
    Ub = tensor_to_unitary(B)
    R = put_env_on_left_site(r)
    L = put_env_on_right_site(l.conj().T)
    Ubd = Ub.conj().T
    ops = [(hadamard(), (3,)), (cnot(), (3, 4)),
           (U, (2, 3)), (U, (1, 2)), (W, (2, 3)),
           (L, (0, 1)), (R, (4, 5)),
           (Ubd, (1, 2)), (Ubd, (2, 3)),
           (cnot(), (3, 4)), (hadamard(), (3,))]
    amp = simulate(ops, 6)[0]
    return float(-np.sqrt(2 * abs(amp)))


def loschmidt_cost(A, B, W):
    """-sqrt(|eta_2|), eta_2 the leading eigenvalue of
    Map(W . merge(A,A), merge(B,B)) (the script's own commented check
    ``np.sqrt(np.abs(x[0]))`` at loschmidts/time_evo.py:113)."""
    x, _ = right_fixed_point(apply_two_site_gate(W, merge(A, A)), merge(B, B))
    return float(-np.sqrt(abs(x)))


def get_overlap_exact(p1, p2, gate=shallow_full_state_tensor, testing=True):
    """qmps/time_evolve_tools.py:84-91.  ``left_canonicalise`` is a pure gauge
    change on tensors that come from a unitary and is skipped."""
    A = unitary_to_tensor(gate(p1))
    B = unitary_to_tensor(gate(p2))
    x, r = right_fixed_point(A, B)
    return (abs(x) ** 2, r) if testing else abs(x) ** 2


# ---- a12: rotosolve --------------------------------------------------------
def wrap_angle(x):
    return np.arctan2(np.sin(x), np.cos(x))


def rotosolve_theta3(e0, ep, em):
    """qmps/rotosolve.py:175: theta* from the cost at shifts 0, +pi/2, -pi/2."""
    return -np.pi / 2 - np.arctan2(2 * e0 - ep - em, ep - em)


def rotosolve_step3(theta_i, e0, ep, em):
    """qmps/rotosolve.py:175-177: updated (wrapped) coordinate."""
    return wrap_angle(theta_i + wrap_angle(rotosolve_theta3(e0, ep, em)))


def double_rotosolve_fit(M0, Mpi, Mp2, Mm2, Mp4, Mm4):
    """qmps/tools.py:434-447: two-frequency fit from six shifted costs.
    Returns (a, b, c, d, P, u, Q, v) of f(x) = P sin(2x+u) + Q sin(x+v)."""
    A, B = M0 + Mpi, M0 - Mpi
    C, Dd = Mp2 + Mm2, Mp2 - Mm2
    E = Mp4 - Mm4
    a, b, c, d = (2 * E - np.sqrt(2) * Dd) / 4, (A - C) / 4, Dd / 2, B / 2
    return a, b, c, d, np.hypot(a, b), np.arctan2(b, a), np.hypot(c, d), np.arctan2(d, c)


def double_rotosolve(eps, initial_parameters, N_iters=100):
    """qmps/tools.py:422-457 without printing: in-place coordinate sweeps.
    Returns (history, params)."""
    params = initial_parameters
    eye = np.eye(len(params))
    history = []
    for _ in range(N_iters):
        for i in range(len(params)):
            def M(x):
                return np.sum(eps(params + eye[i] * x))
            _, _, _, _, P, u, Q, v = double_rotosolve_fit(
                M(0), M(np.pi), M(np.pi / 2), M(-np.pi / 2), M(np.pi / 4), M(-np.pi / 4))
            th = minimize_scalar(lambda x: P * np.sin(2 * x + u) + Q * np.sin(x + v),
                                 bounds=[-np.pi, np.pi]).x
            params[i] += wrap_angle(th)
        history.append(eps(params))
    return history, params


# ---- a13: analytic TFIM anchors --------------------------------------------
def exact_loschmidt_f(z, g0, g1):
    """qmps/loschmidts/exact_loschmidt.py:6-17."""
    def theta(k, g):
        return np.arctan2(np.sin(k), g - np.cos(k)) / 2

    def integrand(k):
        phi = theta(k, g0) - theta(k, g1)
        eps_k = -2 * np.sqrt((g1 - np.cos(k)) ** 2 + np.sin(k) ** 2)
        return -np.log(np.cos(phi) ** 2 + np.sin(phi) ** 2 * np.exp(-2 * z * eps_k)) / (2 * np.pi)

    return quad(integrand, 0, np.pi)[0]


def exact_loschmidt(t, g0, g1):
    """qmps/loschmidts/exact_loschmidt.py:19-20."""
    return exact_loschmidt_f(1j * t, g0, g1) + exact_loschmidt_f(-1j * t, g0, g1)


def exact_loschmidts(T, g0, g1):
    return np.array([exact_loschmidt(t, g0, g1) for t in T])


def tfim_e0_exact(g):
    """tests/test_ground_state.py:101-102."""
    return quad(lambda k: -2 * np.sqrt(1 + g * g - 2 * g * np.cos(k)) / np.pi / 2, 0, np.pi)[0]
