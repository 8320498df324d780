"""Drop-in mirror of ``qmps.rotosolve`` (qmps/rotosolve.py:15-241), classical part.

``rotosolve`` / ``double_rotosolve`` keep the reference signatures (Hamiltonian matrix,
state function returning a state vector, parameters mutated in place).  The batched,
all-on-device sweeps are ``qmps_b200.batched.rotosolve_sweeps``.
"""
import numpy as np
from scipy.optimize import minimize_scalar

from .represent import ShallowFullStateTensor

__all__ = ["gate", "rotosolve", "double_rotosolve", "rotosolve_theta", "double_rotosolve_coefficients"]

π = np.pi


def gate(v, symbol="U"):
    """qmps/rotosolve.py:15-18."""
    return ShallowFullStateTensor(2, v, symbol)


def rotosolve_theta(e0, ep, em):
    """Closed-form minimiser shift from the cost at 0, +pi/2, -pi/2 (qmps/rotosolve.py:175)."""
    return -π / 2 - np.arctan2(2 * e0 - ep - em, ep - em)


def double_rotosolve_coefficients(M0, Mpi, Mp2, Mm2, Mp4, Mm4):
    """(a, b, c, d, P, u, Q, v) of qmps/tools.py:434-447."""
    A, B, C, D, E = M0 + Mpi, M0 - Mpi, Mp2 + Mm2, Mp2 - Mm2, Mp4 - Mm4
    a, b, c, d = (2 * E - np.sqrt(2) * D) / 4, (A - C) / 4, D / 2, B / 2
    return a, b, c, d, np.sqrt(a ** 2 + b ** 2), np.arctan2(b, a), np.sqrt(c ** 2 + d ** 2), np.arctan2(d, c)


def rotosolve(H, state_function, initial_parameters, args=(), N_iters=10):
    """qmps/rotosolve.py:154-181 (without the plotting call)."""
    es, S = [], []
    eye = np.eye(len(initial_parameters))
    params = initial_parameters

    def ϵ(x):
        ψ = state_function(x, *args)
        return np.real(ψ.conj().T @ H @ ψ)

    for _ in range(N_iters):
        for i in range(len(params)):
            θ_ = rotosolve_theta(ϵ(params), ϵ(params + eye[i] * π / 2), ϵ(params - eye[i] * π / 2))
            params[i] += np.arctan2(np.sin(θ_), np.cos(θ_))
            params[i] = np.arctan2(np.sin(params[i]), np.cos(params[i]))
        es.append(ϵ(params))
        S.append(params.copy())
    return es, S


def double_rotosolve(H, state_function, initial_parameters, args=(), N_iters=5):
    """qmps/rotosolve.py:183-241."""
    es = []
    eye = np.eye(len(initial_parameters))
    params = initial_parameters

    def ϵ(x):
        ψ = state_function(x, *args)
        return np.real(ψ.conj().T @ H @ ψ)

    for _ in range(N_iters):
        for i in range(len(params)):
            def M(x):
                return np.sum(ϵ(params + eye[i] * x))
            _, _, _, _, P, u, Q, v = double_rotosolve_coefficients(M(0), M(π), M(π / 2), M(-π / 2), M(π / 4), M(-π / 4))
            θ_ = minimize_scalar(lambda x: P * np.sin(2 * x + u) + Q * np.sin(x + v), bounds=[-π, π]).x
            params[i] += np.arctan2(np.sin(θ_), np.cos(θ_))
        es.append(ϵ(params))
    return np.array(es), params
