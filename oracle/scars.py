"""CPU restatement of the PXP scar-dynamics driver of the reference (``scars.py`` at the reference root):
a caller of the same ``Map(merge(A1,A2), merge(A1',A2')).right_fixed_point()`` eigen-solve.
TEST INFRASTRUCTURE ONLY -- the product path never imports this.

Pinned by ``tests/golden/ref_scars.npz`` (``oracle/make_golden_scars.py`` runs the reference's own
``scars_time_evolve_cost_function`` / ``scars_cost_fun_alternate`` / ``A`` / ``H`` / ``W`` / ``func_list`` under stubs).
"""
import numpy as np
import scipy.linalg as sla

from .tensors import merge, right_fixed_point, tensor_to_unitary, put_env_on_left_site, put_env_on_right_site
from .gates import hadamard, cnot, simulate

__all__ = ["scars_tensor", "scars_hamiltonian", "scars_evolution_gate", "scar_ansatz_unitary", "scar_gate_unitary",
           "scars_cost_circuit", "scars_cost", "scars_tdvp_rhs"]

_P = np.array([[0, 0], [0, 1]], dtype=complex)
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_n = np.array([[1, 0], [0, 0]], dtype=complex)
_I = np.eye(2, dtype=complex)


def _kron(*ops):
    out = np.eye(1, dtype=complex)
    for o in ops:
        out = np.kron(out, o)
    return out


def scars_tensor(theta, phi):
    """``A(theta, phi)`` (scars.py:70-73): the D = 2 PXP ansatz tensor [s][i][j]."""
    return np.array([[[0, 1j * np.exp(-1j * phi)], [0, 0]],
                     [[np.cos(theta), 0], [np.sin(theta), 0]]], dtype=complex)


def scars_hamiltonian(mu):
    """``H(mu)`` (scars.py:23-27): PXP on the middle two of four sites + chemical potential, 16 x 16."""
    return 0.5 * (_kron(_I, _P, _X, _P) + _kron(_P, _X, _P, _I)) + (mu / 4) * (
        _kron(_I, _I, _I, _n) + _kron(_I, _I, _n, _I) + _kron(_I, _n, _I, _I) + _kron(_n, _I, _I, _I))


def scars_evolution_gate(mu, dt):
    """``W(mu, dt)`` (scars.py:29): expm(+1j dt H) -- the sign is the reference's."""
    return sla.expm(1j * dt * scars_hamiltonian(mu))


def _on(g, qubits, n):
    from .gates import on_qubits
    return on_qubits(g, qubits, n)


def scar_ansatz_unitary(theta, phi):
    """``ScarsAnsatz([theta, phi])`` (scars.py:31-51) as a 4 x 4 matrix; gate matrices in cirq's conventions
    (ZPowGate(t) = diag(1, e^{i pi t}), CNotPowGate(t) = |0><0| x 1 + |1><1| x X^t, X^t = e^{i pi t/2} R_x(pi t))."""
    def zpow(t):
        return np.diag([1, np.exp(1j * np.pi * t)])

    def xpow(t):
        c, s = np.cos(np.pi * t / 2), np.sin(np.pi * t / 2)
        return np.exp(1j * np.pi * t / 2) * np.array([[c, -1j * s], [-1j * s, c]])

    def cxpow(t):                                # control = first qubit of the call
        return np.block([[np.eye(2), np.zeros((2, 2))], [np.zeros((2, 2)), xpow(t)]])
    S = np.diag([1, 1j])
    ops = [(zpow(0.5 - phi / np.pi), (1,)), (_X, (0,)), (cnot(), (0, 1)), (_X, (0,)),
           (cxpow(2 * theta / np.pi), (1, 0)), (S, (0,)), (zpow(-theta / np.pi), (1,))]
    U = np.eye(4, dtype=complex)
    for g, qs in ops:
        U = _on(g, qs, 2) @ U
    return U


def scar_gate_unitary(params):
    """``ScarGate([theta, phi, phi', theta'])`` (scars.py:53-68): 8 x 8."""
    th, ph, ph_, th_ = params
    return _on(scar_ansatz_unitary(th, ph), (0, 1), 3) @ _on(scar_ansatz_unitary(th_, ph_), (1, 2), 3)


def _cell(params):
    th1, ph1, ph2, th2 = params
    return merge(scars_tensor(th1, ph1), scars_tensor(th2, ph2))


def scars_cost_circuit(params, current_params, W, gates="tensor"):
    """The reference's 8-qubit read-out (scars.py:76-111 with ``gates="circuit"``, :113-155 with ``"tensor"``):
    ``-2 |<0|C|0>|`` with r from ``Map(merge(A1,A2), merge(A1',A2')).right_fixed_point()``."""
    M, M_ = _cell(current_params), _cell(params)
    _, r = right_fixed_point(M, M_)
    R = put_env_on_left_site(r)
    L = put_env_on_right_site(r.conj().T)
    if gates == "circuit":
        U, U_ = scar_gate_unitary(current_params), scar_gate_unitary(params)
    else:
        U, U_ = tensor_to_unitary(M), tensor_to_unitary(M_)
    Ud = U_.conj().T
    ops = [(hadamard(), (5,)), (cnot(), (5, 6)), (U, (3, 4, 5)), (U, (1, 2, 3)), (L, (0, 1)), (W, (2, 3, 4, 5)),
           (R, (6, 7)), (Ud, (1, 2, 3)), (Ud, (3, 4, 5)), (cnot(), (5, 6)), (hadamard(), (5,))]
    return float(-np.abs(simulate(ops, 8)[0]) * 2)


def scars_cost(params, current_params, W):
    """The same number as a tensor contraction (what the CUDA kernel computes):
    Phi[i,(s,t),b] = (M^s M^t)[i,b] for two unit cells, T = Phi'^dagger W Phi, and with the unit-norm r
    amplitude = 1/2 sum (r^dagger)[i',i] (r^T)[b',b] T[(i',b'),(i,b)]."""
    M, M_ = _cell(current_params), _cell(params)
    _, r = right_fixed_point(M, M_)
    Phi = np.einsum("sik,tkb->istb", M, M).reshape(2, 16, 2)
    Phi_ = np.einsum("sik,tkb->istb", M_, M_).reshape(2, 16, 2)
    T = np.einsum("xpy,pq,iqb->xyib", Phi_.conj(), W, Phi)
    amp = 0.5 * np.einsum("xi,yb,xyib->", r.conj().T, r.T, T) / np.sum(np.abs(r) ** 2)
    return float(-2 * np.abs(amp))


def scars_tdvp_rhs(angles, mu):
    """``func_list`` (scars.py:175-181): the classical TDVP equations of motion of the PXP ansatz."""
    from numpy import sin, cos, tan

    def dth(t1, p1, p2, t2):
        return tan(t2) * sin(t1) * (cos(t1) ** 2) * cos(p1) + cos(t2) * cos(p2)

    def dph(t1, p1, p2, t2):
        return 2 * tan(t1) * cos(t2) * sin(p2) - 0.5 * tan(t2) * cos(t1) * sin(p1) * (2 * (sin(t2) ** -2) + cos(2 * t1) - 5)
    a = list(angles)
    return np.array([dth(*a), -mu + dph(*a), -mu + dph(*reversed(a)), dth(*reversed(a))])
