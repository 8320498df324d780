"""Drop-in mirror of ``qmps.tools`` for the classical hot path.

Same names, argument meaning and error behaviour as the reference functions
(qmps/tools.py:36-186, 195-270, 422-464); numpy in, numpy out.  Each call is a
batch of one through the C ABI into the CUDA kernels -- use ``qmps_b200.batched`` to
amortise launches over many problems.  There is no CPU fallback: without the built
library and a CUDA device these functions raise ``qmps_b200.QmpsError``.
"""
from dataclasses import dataclass, field
from typing import Any

import numpy as np
from numpy.linalg import LinAlgError
from scipy.optimize import minimize, minimize_scalar

__all__ = [
    "random_unitary", "svals", "from_real_vector", "to_real_vector", "eye_like", "cT", "direct_sum",
    "unitary_extension", "environment_to_unitary", "environment_from_unitary", "tensor_to_unitary",
    "unitary_to_tensor", "split_2s", "split_3s", "split_ns", "get_env_exact", "get_env_exact_alternative",
    "env_exact_parts", "Optimizer", "OptimizerCircuit", "double_rotosolve", "RotosolveResult",
]


# ---- bookkeeping helpers (no arithmetic worth a kernel) ------------------------------
def random_unitary(*args):
    """qmps/tools.py:36-37 (real orthogonal from QR of a Gaussian matrix)."""
    return np.linalg.qr(np.random.randn(*args))[0]


def svals(A):
    return np.linalg.svd(A, compute_uv=False)


def from_real_vector(v):
    """(re..., im...) -> complex vector (qmps/tools.py:43-46)."""
    v = np.asarray(v)
    half = len(v) // 2
    return v[:half] + 1j * v[half:]


def to_real_vector(A):
    """qmps/tools.py:49-52."""
    A = np.asarray(A)
    return np.concatenate([A.real.reshape(-1), A.imag.reshape(-1)], axis=0)


def eye_like(A):
    return np.eye(A.shape[0])


def cT(tensor):
    """Hermitian conjugate of the last two indices (qmps/tools.py:61-66)."""
    return np.swapaxes(np.conj(tensor), -1, -2)


def direct_sum(A, B):
    """qmps/tools.py:69-73."""
    out = np.zeros((A.shape[0] + B.shape[0], A.shape[1] + B.shape[1]), dtype=np.result_type(A, B))
    out[:A.shape[0], :A.shape[1]] = A
    out[A.shape[0]:, A.shape[1]:] = B
    return out


def split_2s(x):
    return [x[i:i + 2] for i in range(0, len(x), 2)]


def split_3s(x):
    return [x[i:i + 3] for i in range(0, len(x), 3)]


def split_ns(x, n):
    return [x[i:i + n] for i in range(0, len(x), n)]


# ---- a1 / a2 / a3 --------------------------------------------------------------------
def unitary_to_tensor(U):
    """A[s, i, j] = U[(i, s), (0, j)]  (qmps/tools.py:151-154)."""
    from . import batched
    U = np.asarray(U)
    return batched.unitary_to_tensor(U[None].astype(np.complex128)).cpu().numpy()[0]


def _complete_columns(Q):
    """tall isometry Q (m x k, k | m) -> m x m unitary whose first k columns are Q."""
    from . import batched
    m, k = Q.shape
    if m % k:
        raise ValueError("unitary_extension: the row count must be a multiple of the column count")
    d = m // k
    # pack Q as a tensor A[s][i][j] with iso[(i,s)][j] = Q[i*d+s][j]
    A = np.ascontiguousarray(Q.reshape(k, d, k).transpose(1, 0, 2).astype(np.complex128))
    return batched.tensor_to_unitary(A[None]).cpu().numpy()[0]


def unitary_extension(Q, D=None):
    """Extend an isometry to a unitary (qmps/tools.py:76-94).  Q's own columns (rows for
    a wide Q) are kept exactly; the completion is any orthonormal one, as in the
    reference (its SVD null space is not unique either)."""
    Q = np.asarray(Q)
    rows, cols = Q.shape
    if rows > cols:
        full = _complete_columns(Q)
    elif rows < cols:
        full = _complete_columns(Q.conj().T).conj().T
    else:
        full = Q
    if D is not None and D > full.shape[0]:
        full = direct_sum(full, np.eye(D - full.shape[0]))
    return full


def tensor_to_unitary(A, testing=False):
    """Left-isometric tensor -> unitary with U[:, :D] = iso (qmps/tools.py:123-148)."""
    from . import batched
    A = np.asarray(A, dtype=np.complex128)
    d, D, _ = A.shape
    U = batched.tensor_to_unitary(np.ascontiguousarray(A)[None]).cpu().numpy()[0]
    if testing:                                   # the reference's D=2 checks, generalised
        iso = A.transpose(1, 0, 2).reshape(D * d, D)
        n = D * d
        passed = (np.allclose(cT(iso) @ iso, np.eye(D)) and np.allclose(U @ cT(U), np.eye(n))
                  and np.allclose(cT(U) @ U, np.eye(n)) and np.allclose(U[:, :D], iso)
                  and (np.allclose(unitary_to_tensor(U), A) if d == 2 else True))
        return U, bool(passed)
    return U


def environment_to_unitary(v):
    """vec(v)/|v| as the first column of a unitary (qmps/tools.py:97-108)."""
    from . import batched
    v = np.asarray(v, dtype=np.complex128).reshape(1, -1)
    return batched.environment_to_unitary(v).cpu().numpy()[0]


def environment_from_unitary(u):
    """First column reshaped to a square matrix (qmps/tools.py:111-120; the reference
    hard-codes 2x2)."""
    u = np.asarray(u)
    D = int(round(np.sqrt(u.shape[0])))
    return u[:, 0].reshape(D, D)


# ---- a4 / a5 ---------------------------------------------------------------------------
def env_exact_parts(U):
    """(eta, r, C) for one unitary; raises ``numpy.linalg.LinAlgError`` where the reference's
    ``cholesky`` would (qmps/tools.py:182)."""
    from . import batched, _lib
    U = np.asarray(U, dtype=np.complex128)
    res = batched.env_exact(U=np.ascontiguousarray(U)[None])
    st = int(res.status.cpu()[0])
    if st == _lib.ST_NOT_PD:
        raise LinAlgError("environment is not positive definite (Cholesky failed)")
    if st == _lib.ST_SINGULAR:
        raise LinAlgError("transfer matrix has a degenerate leading eigenvalue")
    return complex(res.eta.cpu()[0]), res.r.cpu().numpy()[0], res.C.cpu().numpy()[0]


def get_env_exact(U):
    """Exact environment unitary V of a state unitary U (qmps/tools.py:176-182): one C-ABI call on host buffers
    (``qmps_get_env_exact_host``: H2D, solve, Cholesky, environment_to_unitary, D2H, a single synchronisation)."""
    from . import _lib
    U = np.ascontiguousarray(np.asarray(U, dtype=np.complex128))
    D = U.shape[0] // 2
    V = np.empty((D * D, D * D), dtype=np.complex128)
    st = np.zeros(1, dtype=np.int32)
    lib = _lib.require_device()
    _lib.check(lib.qmps_get_env_exact_host(D, 1, U.ctypes.data, V.ctypes.data, st.ctypes.data, _lib.C128, _device_index()),
               "get_env_exact_host")
    if st[0] == _lib.ST_NOT_PD:
        raise LinAlgError("environment is not positive definite (Cholesky failed)")
    if st[0] == _lib.ST_SINGULAR:
        raise LinAlgError("transfer matrix has a degenerate leading eigenvalue")
    return V


def _device_index():
    import torch
    return torch.cuda.current_device()


def get_env_exact_alternative(U):
    """qmps/tools.py:184-186: ``AL, AR, C = iMPS([unitary_to_tensor(U)]).mixed()`` ->
    ``environment_to_unitary(C)``.  The tensor of a unitary is already left-canonical; C is taken
    in the lower-triangular (Cholesky) gauge, so this equals ``get_env_exact(U)``."""
    from . import batched, _lib
    A = unitary_to_tensor(np.asarray(U, dtype=np.complex128))
    res = batched.mixed_canonical(np.ascontiguousarray(A)[None], assume_left_canonical=True)
    if int(res.status.cpu()[0]) != _lib.ST_OK:
        raise LinAlgError("environment is not positive definite (Cholesky failed)")
    return environment_to_unitary(res.C.cpu().numpy()[0])


# ---- host-side drivers (the cost they call runs on the GPU; scipy stays on the host) -------------
# Interface contract taken from qmps/tools.py:195-270, 459-464: the attribute names callers read
# (`u`, `v`, `initial_guess`, `iters`, `optimized_result`, `obj_fun_values`, `settings`, `circuit`),
# the settings keys, and the hooks subclasses override (`gate_from_params`, `update_state`,
# `objective_function`).  The bodies are this package's own.
_DEFAULT_SETTINGS = (("maxiter", 10000), ("verbose", True), ("method", "Nelder-Mead"), ("tol", 1e-8),
                     ("store_values", True), ("bayesian", False))


@dataclass
class OptimizerCircuit:
    """Plain record of a circuit and its qubit bookkeeping (qmps/tools.py:195-200)."""
    circuit: Any = None
    total_qubits: Any = None
    aux_qubits: Any = None
    qubits: Any = field(default=None, init=False)


@dataclass
class RotosolveResult:
    """What the rotosolve drivers return: the scipy-result fields callers read (qmps/tools.py:459-464)."""
    history: Any
    fun: Any
    x: Any
    message: str = ""


class Optimizer:
    """Base class of every variational optimiser in the reference (qmps/tools.py:203-270).
    Subclasses supply ``objective_function`` (or pass ``obj_fun``/``args``) and may override the
    ``gate_from_params`` / ``update_state`` hooks; ``optimize`` dispatches on ``settings``."""

    def __init__(self, u=None, v=None, initial_guess=None, obj_fun=None, args=None):
        self.u, self.v = u, v
        self.initial_guess = initial_guess
        self.obj_fun, self.args = obj_fun, args
        self.settings = dict(_DEFAULT_SETTINGS)
        self.is_verbose = self.settings["verbose"]
        self.circuit = OptimizerCircuit()
        self.iters, self.obj_fun_values, self.optimized_result = 0, [], None

    # -- hooks ------------------------------------------------------------------------------
    def gate_from_params(self, params):
        return None

    def update_state(self):
        return None

    def objective_function(self, params):
        if self.obj_fun is None:
            return None
        return self.obj_fun(params, *(self.args or ()))

    # -- driver -------------------------------------------------------------------------------
    def change_settings(self, new_settings):
        return self.settings.update(new_settings)

    def callback_store_values(self, xk):
        """scipy callback: record (and, when verbose, print ``iteration:value``) the cost at ``xk``."""
        value = self.objective_function(xk)
        self.obj_fun_values.append(value)
        if self.settings["verbose"]:
            print("%d:%s" % (self.iters, value))
        self.iters += 1

    def _run_minimiser(self):
        cfg = self.settings
        if cfg["bayesian"]:
            raise NotImplementedError("bayesian optimisation needs skopt, which the reference has commented out")
        if cfg["method"] == "Rotosolve":
            return double_rotosolve(self.objective_function, self.initial_guess, cfg["maxiter"], cfg["verbose"])
        return minimize(self.objective_function, self.initial_guess, method=cfg["method"], tol=cfg["tol"],
                        callback=self.callback_store_values if cfg["store_values"] else None,
                        options={"maxiter": cfg["maxiter"], "disp": cfg["verbose"]})

    def optimize(self):
        self.optimized_result = res = self._run_minimiser()
        self.update_state()
        if self.is_verbose:
            print("Reason for termination is %s \nObjective Function Value is %s" % (res.message, res.fun))
        return res


def double_rotosolve(ϵ, initial_parameters, N_iters=100, disp=True):
    """Two-frequency coordinate sweeps with a Python callable cost (qmps/tools.py:422-457).
    ``initial_parameters`` is updated IN PLACE, as in the reference.  The six shifted
    evaluations per coordinate are issued as one batch when the callable supports it
    (attribute ``batched``: theta[N, P] -> cost[N])."""
    params = initial_parameters
    eye = np.eye(len(params))
    shifts = np.array([0.0, np.pi, np.pi / 2, -np.pi / 2, np.pi / 4, -np.pi / 4])
    history = []
    batched_cost = getattr(ϵ, "batched", None)
    for w in range(N_iters):
        if disp:
            print(w, ", ", sep="", end="", flush=True)
        for i in range(len(params)):
            if batched_cost is not None:
                M0, Mpi, Mp2, Mm2, Mp4, Mm4 = np.asarray(batched_cost(params[None, :] + shifts[:, None] * eye[i][None, :]))
            else:
                M0, Mpi, Mp2, Mm2, Mp4, Mm4 = [np.sum(ϵ(params + eye[i] * x)) for x in shifts]
            A, B, C, D, E = M0 + Mpi, M0 - Mpi, Mp2 + Mm2, Mp2 - Mm2, Mp4 - Mm4
            a, b, c, d = (2 * E - np.sqrt(2) * D) / 4, (A - C) / 4, D / 2, B / 2
            P, u = np.sqrt(a ** 2 + b ** 2), np.arctan2(b, a)
            Q, v = np.sqrt(c ** 2 + d ** 2), np.arctan2(d, c)
            θ_ = minimize_scalar(lambda x: P * np.sin(2 * x + u) + Q * np.sin(x + v), bounds=[-np.pi, np.pi]).x
            params[i] += np.arctan2(np.sin(θ_), np.cos(θ_))
        if disp:
            print("\n", sep="", end="", flush=True)
        history.append(ϵ(params))
    return RotosolveResult(history, history[-1], params, "")
