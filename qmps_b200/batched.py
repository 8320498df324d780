"""Batched device API: torch CUDA tensors in, torch CUDA tensors out.

torch is plumbing here (device memory, streams); every computation goes through the
C ABI of ``include/qmps_b200.h`` into the hand-written sm_100a kernels.  All calls
are stream-ordered on ``torch.cuda.current_stream()`` and do not synchronise.

Shapes: tensors ``A[..., d, D, D]`` with ``A[s, i, j]`` as produced by
``unitary_to_tensor`` (qmps/tools.py:151-154); parameter vectors ``theta[N, P]``
(float64).
"""
from collections import namedtuple

import collections

import numpy as np
import torch

from . import _lib as L

EnvResult = namedtuple("EnvResult", "eta r C status")
FixedPoint = namedtuple("FixedPoint", "eta vec cost echo fid status")
RotoFit = namedtuple("RotoFit", "theta_star fit")
Canonical = namedtuple("Canonical", "AL eta L status")
Mixed = namedtuple("Mixed", "AL AR C eta status")
TdvpRun = namedtuple("TdvpRun", "A traj rates energy status")

_CDT = {torch.complex128: L.C128, torch.complex64: L.C64}
_RDT = {torch.complex128: torch.float64, torch.complex64: torch.float32}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _cdev(x, dtype=torch.complex128, device=None):
    """numpy / torch -> contiguous complex CUDA tensor."""
    L.require_device()
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if device is None:
        device = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
    return x.to(device=device, dtype=dtype).contiguous()


def _rdev(x, device=None):
    L.require_device()
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64)))
    if device is None:
        device = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
    return x.to(device=device, dtype=torch.float64).contiguous()


def _dt(t):
    if t.dtype not in _CDT:
        raise TypeError(f"complex128 or complex64 expected, got {t.dtype}")
    return _CDT[t.dtype]


# ---- a1 / a2 / a3 -------------------------------------------------------------------
def unitary_to_tensor(U):
    """U[N, 2D, 2D] -> A[N, 2, D, D]  (qmps/tools.py:151-154)."""
    U = _cdev(U, U.dtype if isinstance(U, torch.Tensor) and U.dtype in _CDT else torch.complex128)
    N, m, _ = U.shape
    D = m // 2
    A = torch.empty((N, 2, D, D), dtype=U.dtype, device=U.device)
    with torch.cuda.device(U.device):
        L.check(L.load().qmps_unitary_to_tensor(D, N, _p(U), _p(A), _dt(U), _stream()), "unitary_to_tensor")
    return A


def tensor_to_unitary(A):
    """A[N, d, D, D] (left-canonical) -> U[N, dD, dD], U[:, :, :D] = iso  (qmps/tools.py:123-148)."""
    A = _cdev(A, A.dtype if isinstance(A, torch.Tensor) and A.dtype in _CDT else torch.complex128)
    N, d, D, _ = A.shape
    U = torch.empty((N, d * D, d * D), dtype=A.dtype, device=A.device)
    with torch.cuda.device(A.device):
        L.check(L.load().qmps_tensor_to_unitary(d, D, N, _p(A), _p(U), _dt(A), _stream()), "tensor_to_unitary")
    return U


def environment_to_unitary(v):
    """v[N, n] -> V[N, n, n] unitary with V[:, :, 0] = v/|v|  (qmps/tools.py:97-108)."""
    v = _cdev(v, v.dtype if isinstance(v, torch.Tensor) and v.dtype in _CDT else torch.complex128)
    N, n = v.shape
    V = torch.empty((N, n, n), dtype=v.dtype, device=v.device)
    with torch.cuda.device(v.device):
        L.check(L.load().qmps_environment_to_unitary(n, N, _p(v), _p(V), _dt(v), _stream()), "environment_to_unitary")
    return V


# ---- a4 / a5 -------------------------------------------------------------------------
def env_exact(A=None, U=None, assume_left_canonical=True, want_eta=True, want_r=True, want_C=True,
              want_status=True):
    """Exact right environment of E_AA for a batch (qmps/tools.py:176-182).

    Pass either tensors ``A[N, d, D, D]`` or unitaries ``U[N, 2D, 2D]``.  Returns
    ``EnvResult(eta[N], r[N,D,D], C[N,D,D], status[N])`` (None for outputs not asked for).
    """
    if (A is None) == (U is None):
        raise ValueError("pass exactly one of A, U")
    x = A if A is not None else U
    x = _cdev(x, x.dtype if isinstance(x, torch.Tensor) and x.dtype in _CDT else torch.complex128)
    N = x.shape[0]
    if A is not None:
        d, D = x.shape[1], x.shape[2]
    else:
        d, D = 2, x.shape[1] // 2
    dev, cd = x.device, x.dtype
    eta = torch.empty((N,), dtype=cd, device=dev) if want_eta else None
    r = torch.empty((N, D, D), dtype=cd, device=dev) if want_r else None
    C = torch.empty((N, D, D), dtype=cd, device=dev) if want_C else None
    st = torch.empty((N,), dtype=torch.int32, device=dev) if want_status else None
    with torch.cuda.device(dev):
        L.check(L.load().qmps_env_exact(d, D, N, _p(x), int(U is not None), int(bool(assume_left_canonical)),
                                        _p(eta), _p(r), _p(C), _p(st), _dt(x), _stream()), "env_exact")
    return EnvResult(eta, r, C, st)


def env_exact_packed(A=None, U=None):
    """D = 2 complex128 left-canonical environments as 64-byte records (``qmps_env_exact_packed``):
    ``packed[N, 8] = [r00, Re r01, Im r01, c00, Re c10, Im c10, c11, status]``; ``unpack_env`` expands them."""
    if (A is None) == (U is None):
        raise ValueError("pass exactly one of A, U")
    x = _cdev(A if A is not None else U, torch.complex128)
    N = x.shape[0]
    out = torch.empty((N, 8), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.load().qmps_env_exact_packed(N, _p(x), int(U is not None), _p(out), _stream()), "env_exact_packed")
    return out


def env_exact_packed_host(x, is_unitary=False, out=None, device=0):
    """Host-buffer form (numpy in, numpy out; pinned buffers make it link-bound): x = A[N,2,2,2] or U[N,4,4]."""
    x = np.ascontiguousarray(x, dtype=np.complex128)
    N = x.shape[0]
    if out is None:
        out = np.empty((N, 8), dtype=np.float64)
    L.check(L.require_device().qmps_env_exact_packed_host(N, x.ctypes.data, int(bool(is_unitary)), out.ctypes.data, int(device)),
            "env_exact_packed_host")
    return out


def unpack_env(packed):
    """packed[N, 8] (numpy) -> (eta[N] = 1, r[N,2,2], C[N,2,2], status[N]): the outputs of ``env_exact``."""
    p = np.asarray(packed)
    N = p.shape[0]
    r = np.empty((N, 2, 2), dtype=np.complex128)
    C = np.zeros((N, 2, 2), dtype=np.complex128)
    r[:, 0, 0] = p[:, 0]; r[:, 1, 1] = 1.0 - p[:, 0]
    r[:, 0, 1] = p[:, 1] + 1j * p[:, 2]; r[:, 1, 0] = p[:, 1] - 1j * p[:, 2]
    C[:, 0, 0] = p[:, 3]; C[:, 1, 0] = p[:, 4] + 1j * p[:, 5]; C[:, 1, 1] = p[:, 6]
    return np.ones(N, dtype=np.complex128), r, C, p[:, 7].astype(np.int32)


# ---- a6 / a8 / a11 ---------------------------------------------------------------------
def fixed_point(A, B, pair="elementwise", left=False, want_vec=True, want_costs=True, want_status=True,
                gauge="zgeev"):
    """Leading eigenpair of the mixed transfer matrix E_AB (``Map(A,B).right_fixed_point()``).

    ``pair='elementwise'``: A[NA], B[NB] broadcast (NA == NB or one of them 1).
    ``pair='outer'``: every (A[ia], B[ib]); outputs are shaped [NA, NB, ...].
    ``gauge``: phase of the unit-norm eigenvector -- ``'zgeev'`` (largest component real positive: what
    the recorded xmps output in the reference's ``Time Evo.ipynb`` cells 22-24 shows) or ``'trace'``
    (tr >= 0, Hermitian-compatible).
    """
    A = _cdev(A, A.dtype if isinstance(A, torch.Tensor) and A.dtype in _CDT else torch.complex128)
    B = _cdev(B, A.dtype, A.device)
    NA, d, D, _ = A.shape
    NB = B.shape[0]
    if B.shape[1:] != A.shape[1:]:
        raise ValueError("A and B must share (d, D, D)")
    outer = pair == "outer"
    shape = (NA, NB) if outer else (max(NA, NB),)
    dev, cd, rd = A.device, A.dtype, _RDT[A.dtype]
    eta = torch.empty(shape, dtype=cd, device=dev)
    vec = torch.empty(shape + (D, D), dtype=cd, device=dev) if want_vec else None
    cost = torch.empty(shape, dtype=rd, device=dev) if want_costs else None
    echo = torch.empty(shape, dtype=rd, device=dev) if want_costs else None
    fid = torch.empty(shape, dtype=rd, device=dev) if want_costs else None
    st = torch.empty(shape, dtype=torch.int32, device=dev) if want_status else None
    with torch.cuda.device(dev):
        L.check(L.load().qmps_fixed_point_ex(d, D, NA, _p(A), NB, _p(B), int(outer), int(bool(left)),
                                             {"trace": L.GAUGE_TRACE, "zgeev": L.GAUGE_ZGEEV}[gauge], _p(eta), _p(vec),
                                             _p(cost), _p(echo), _p(fid), _p(st), _dt(A), _stream()), "fixed_point")
    return FixedPoint(eta, vec, cost, echo, fid, st)


# ---- a7 ------------------------------------------------------------------------------
def merge(A, B, W=None):
    """M[(s1,s2), i, j] = (A^s1 B^s2)[i, j], optionally followed by a two-site gate
    ``tensordot(W, M, [1, 0])`` (qmps/time_evolve_tools.py:20-23, loschmidts/time_evo.py:79)."""
    A = _cdev(A, A.dtype if isinstance(A, torch.Tensor) and A.dtype in _CDT else torch.complex128)
    B = _cdev(B, A.dtype, A.device)
    NA, d1, D, _ = A.shape
    NB, d2 = B.shape[0], B.shape[1]
    NW = 0
    if W is not None:
        W = _cdev(W, A.dtype, A.device)
        NW = W.shape[0]
    N = max(NA, NB, NW)
    if any(n not in (1, N) for n in (NA, NB) + ((NW,) if W is not None else ())):
        raise ValueError(f"merge: batch sizes {NA}, {NB}, {NW} do not broadcast (each must be 1 or {N})")
    M = torch.empty((N, d1 * d2, D, D), dtype=A.dtype, device=A.device)
    with torch.cuda.device(A.device):
        L.check(L.load().qmps_merge(d1, d2, D, NA, _p(A), NB, _p(B), NW, _p(W), _p(M), _dt(A), _stream()), "merge")
    return M


# ---- a14 -----------------------------------------------------------------------------
def ansatz_tensors(program, theta, full_unitary=False, dtype=torch.complex128):
    """theta[N, P] -> A[N, 2, D, D] (or U[N, 2D, 2D]) for a gate program (qmps/represent.py:268-423)."""
    theta = _rdev(theta)
    N, P = theta.shape
    R = 2 ** program.nq
    shape = (N, R, R) if full_unitary else (N, 2, R // 2, R // 2)
    out = torch.empty(shape, dtype=dtype, device=theta.device)
    ops = program.c_ops()
    with torch.cuda.device(theta.device):
        L.check(L.load().qmps_ansatz(ops, len(program), program.nq, N, P, _p(theta), int(full_unitary), _p(out),
                                     _CDT[dtype], _stream()), "ansatz")
    return out


def ansatz_unitaries_host(program, theta_np):
    """numpy theta[N, P] -> numpy U[N, 2D, 2D] (one H2D, one launch, one D2H)."""
    U = ansatz_tensors(program, np.asarray(theta_np, dtype=np.float64), full_unitary=True)
    return U.cpu().numpy()


# ---- a9 / a12 --------------------------------------------------------------------------
def _hdev(H, dtype, device):
    H = _cdev(H, dtype, device)
    if H.shape != (4, 4):
        raise ValueError("H must be a 4x4 two-site Hamiltonian")
    return H


def energy_theta(program, theta, H, coord=None, shifts=None, want_status=False, dtype=torch.complex128):
    """Energy cost for a batch of parameter vectors (qmps/ground_state.py:150-168, 251-266).

    With ``coord``/``shifts`` the rotosolve fan-out is fused into the launch:
    ``energy[n, s] = e(theta[n] + shifts[s] * e_coord)`` (qmps/rotosolve.py:175).
    """
    theta = _rdev(theta)
    N, P = theta.shape
    Hd = _hdev(H, dtype, theta.device)
    sh = None
    ns = 0
    if shifts is not None:
        shifts = np.ascontiguousarray(np.asarray(shifts, dtype=np.float64))
        ns = len(shifts)
        sh = shifts.ctypes.data
    shape = (N, ns) if ns else (N,)
    e = torch.empty(shape, dtype=_RDT[dtype], device=theta.device)
    st = torch.empty(shape, dtype=torch.int32, device=theta.device) if want_status else None
    ops = program.c_ops()
    with torch.cuda.device(theta.device):
        L.check(L.load().qmps_energy_theta(ops, len(program), program.nq, N, P, _p(theta), _p(Hd),
                                           -1 if coord is None else int(coord), sh, ns, _p(e), _p(st),
                                           _CDT[dtype], _stream()), "energy_theta")
    return (e, st) if want_status else e


def energy_tensor(A, H, two_site=False, want_status=False):
    """Energy from tensors: ``A[N, 2, D, D]`` (single-site unit cell) or the two-site block
    ``M[N, 4, D, D] = merge(A1, A2)`` (qmps/ground_state.py:291-331)."""
    A = _cdev(A, A.dtype if isinstance(A, torch.Tensor) and A.dtype in _CDT else torch.complex128)
    N, D = A.shape[0], A.shape[2]
    Hd = _hdev(H, A.dtype, A.device)
    e = torch.empty((N,), dtype=_RDT[A.dtype], device=A.device)
    st = torch.empty((N,), dtype=torch.int32, device=A.device) if want_status else None
    with torch.cuda.device(A.device):
        L.check(L.load().qmps_energy_tensor(D, N, _p(A), int(bool(two_site)), _p(Hd), _p(e), _p(st), _dt(A),
                                            _stream()), "energy_tensor")
    return (e, st) if want_status else e


ROTO3_SHIFTS = (0.0, np.pi / 2, -np.pi / 2)                                   # qmps/rotosolve.py:175
ROTO6_SHIFTS = (0.0, np.pi, np.pi / 2, -np.pi / 2, np.pi / 4, -np.pi / 4)     # qmps/tools.py:434-438


def rotosolve_fit(costs, theta=None, coord=None):
    """costs[N, 3|6] -> RotoFit(theta_star[N], fit[N, 8] or None).  With ``theta``/``coord``
    the coordinate is updated in place on the device (qmps/rotosolve.py:176-177, tools.py:452)."""
    costs = _rdev(costs)
    N, ns = costs.shape
    ts = torch.empty((N,), dtype=torch.float64, device=costs.device)
    fit = torch.empty((N, 8), dtype=torch.float64, device=costs.device) if ns == 6 else None
    P = 0
    if theta is not None:
        if not (theta.is_cuda and theta.dtype == torch.float64 and theta.is_contiguous()):
            raise ValueError("theta must be a contiguous float64 CUDA tensor (updated in place)")
        P = theta.shape[1]
    with torch.cuda.device(costs.device):
        L.check(L.load().qmps_rotosolve_fit(N, ns, _p(costs), _p(ts), _p(fit), _p(theta), P,
                                            0 if coord is None else int(coord), _stream()), "rotosolve_fit")
    return RotoFit(ts, fit)


def rotosolve_sweeps(program, theta, H, n_sweeps=1, double=False, dtype=torch.complex128):
    """Whole coordinate sweeps on the device for a batch of parameter vectors, ONE C-ABI call
    (``qmps_rotosolve_sweep``): for each coordinate a fused (shift fan-out + energy) launch and a
    closed-form update launch -- no host round trip.  ``theta`` is updated in place.
    Returns energy[N] after the last sweep."""
    if not (isinstance(theta, torch.Tensor) and theta.is_cuda and theta.dtype == torch.float64 and theta.is_contiguous()):
        raise ValueError("theta must be a contiguous float64 CUDA tensor (updated in place)")
    N, P = theta.shape
    Hd = _hdev(H, dtype, theta.device)
    e = torch.empty((N,), dtype=_RDT[dtype], device=theta.device)
    ops = program.c_ops()
    with torch.cuda.device(theta.device):
        L.check(L.load().qmps_rotosolve_sweep(ops, len(program), program.nq, N, P, _p(theta), _p(Hd), int(n_sweeps),
                                              int(bool(double)), _p(e), _CDT[dtype], _stream()), "rotosolve_sweep")
    return e


# ---- a11 pipeline ------------------------------------------------------------------------
def loschmidt_costs(program, theta, A0, W, dtype=torch.complex128, want_status=False):
    """cost[p, k] = -sqrt|eta_2|, echo[p, k] = -log|eta_2|^2 with eta_2 the leading eigenvalue
    of Map(W_k . merge(A0, A0), merge(B_p, B_p))  (qmps/loschmidts/time_evo.py:75-116).

    theta[NP, P] parameter sets, A0[2, D, D] the state being evolved, W[NT, 4, 4] two-site
    gates (one per time).  One C-ABI call (``qmps_loschmidt_batched``: ansatz, merge, gate-merge,
    fixed points on one stream); outputs are [NP, NT]."""
    theta = _rdev(theta)
    NP, P = theta.shape
    A0 = _cdev(A0, dtype, theta.device).reshape(*A0.shape[-3:]).contiguous()
    W = _cdev(W, dtype, theta.device)
    if W.dim() == 2:
        W = W[None]
    W = W.contiguous()
    NT = W.shape[0]
    rd = _RDT[dtype]
    cost = torch.empty((NP, NT), dtype=rd, device=theta.device)
    echo = torch.empty((NP, NT), dtype=rd, device=theta.device)
    eta = torch.empty((NP, NT), dtype=dtype, device=theta.device)
    st = torch.empty((NP, NT), dtype=torch.int32, device=theta.device) if want_status else None
    ops = program.c_ops()
    with torch.cuda.device(theta.device):
        L.check(L.load().qmps_loschmidt_batched(ops, len(program), program.nq, NP, P, _p(theta), _p(A0), NT, _p(W),
                                                _p(cost), _p(echo), _p(eta), _p(st), _CDT[dtype], _stream()),
                "loschmidt_batched")
    return (cost, echo, eta, st) if want_status else (cost, echo, eta)


def loschmidt_costs_host(program, theta, A0, W, dtype=np.complex128, device=0, want_echo=True, out_cost=None, out_echo=None):
    """``loschmidt_costs`` on HOST arrays through ``qmps_loschmidt_batched_host`` (numpy in, numpy out; the
    copies are part of the call): theta[NP, P] float64, A0[2, D, D], W[NT, 4, 4] -> cost[NP, NT] (, echo).
    ``out_cost`` / ``out_echo``: caller-owned result arrays, e.g. views of PINNED memory
    (``torch.empty(..., pin_memory=True).numpy()``) -- the 16 NT bytes per parameter set that come back then move at
    the link rate instead of the pageable-copy rate."""
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    A0 = np.ascontiguousarray(A0, dtype=dtype)
    W = np.ascontiguousarray(W, dtype=dtype).reshape(-1, 4, 4)
    NP, P = theta.shape
    NT = W.shape[0]
    rd = np.float64 if np.dtype(dtype) == np.complex128 else np.float32
    cost = np.empty((NP, NT), dtype=rd) if out_cost is None else out_cost
    echo = (np.empty((NP, NT), dtype=rd) if out_echo is None else out_echo) if want_echo else None
    for o in (cost, echo):
        if o is not None and (o.shape != (NP, NT) or o.dtype != rd or not o.flags.c_contiguous):
            raise ValueError("out_cost / out_echo must be C-contiguous [NP, NT] arrays of the real type of dtype")
    ops = program.c_ops()
    L.check(L.require_device().qmps_loschmidt_batched_host(
        ops, len(program), program.nq, NP, P, theta.ctypes.data, A0.ctypes.data, NT, W.ctypes.data, cost.ctypes.data,
        echo.ctypes.data if want_echo else None, None, None, L.C128 if np.dtype(dtype) == np.complex128 else L.C64, int(device)),
        "loschmidt_batched_host")
    return (cost, echo) if want_echo else cost


def energy_theta_host(program, theta, H, coord=None, shifts=None, dtype=np.complex128, device=0, want_status=False):
    """``energy_theta`` on HOST arrays through ``qmps_energy_theta_host``: 8 P bytes in, 8 bytes per shift out, one
    synchronisation (this is also the scalar path of the optimiser classes: one parameter vector per call)."""
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    Hh = np.ascontiguousarray(H, dtype=dtype)
    N, P = theta.shape
    ns = 0 if shifts is None else len(shifts)
    sh = np.ascontiguousarray(shifts, dtype=np.float64) if ns else None
    rd = np.float64 if np.dtype(dtype) == np.complex128 else np.float32
    e = np.empty((N, ns) if ns else (N,), dtype=rd)
    st = np.zeros((N, ns) if ns else (N,), dtype=np.int32) if want_status else None
    ops = program.c_ops()
    L.check(L.require_device().qmps_energy_theta_host(
        ops, len(program), program.nq, N, P, theta.ctypes.data, Hh.ctypes.data, -1 if coord is None else int(coord),
        sh.ctypes.data if ns else None, ns, e.ctypes.data, st.ctypes.data if want_status else None,
        L.C128 if np.dtype(dtype) == np.complex128 else L.C64, int(device)), "energy_theta_host")
    return (e, st) if want_status else e


def overlap_theta(program, theta1, theta2, dtype=torch.complex128):
    """Per-site fidelity |eta(E_AB)|^2 for pairs of parameter vectors
    (``get_overlap_exact``, qmps/time_evolve_tools.py:84-91)."""
    A = ansatz_tensors(program, theta1, dtype=dtype)
    B = ansatz_tensors(program, theta2, dtype=dtype)
    return fixed_point(A, B, want_vec=True)


# ---- canonical forms and local expectation values ----------------------------
def left_canonicalise(A, want_L=False, want_status=True):
    """``iMPS([A]).left_canonicalise()`` for a batch A[N, d, D, D] of normalisable tensors
    (call sites qmps/time_evolve_tools.py:85-86, qmps/loschmidts/time_evo.py:76,143):
    AL = L A L^-1 / sqrt(eta), sum_s AL_s^dagger AL_s = 1.  Returns ``Canonical(AL, eta, L, status)``."""
    A = _cdev(A, A.dtype if isinstance(A, torch.Tensor) and A.dtype in _CDT else torch.complex128)
    N, d, D, _ = A.shape
    AL = torch.empty_like(A)
    eta = torch.empty((N,), dtype=A.dtype, device=A.device)
    Lm = torch.empty((N, D, D), dtype=A.dtype, device=A.device) if want_L else None
    st = torch.empty((N,), dtype=torch.int32, device=A.device) if want_status else None
    with torch.cuda.device(A.device):
        L.check(L.load().qmps_left_canonicalise(d, D, N, _p(A), _p(AL), _p(eta), _p(Lm), _p(st), _dt(A), _stream()),
                "left_canonicalise")
    return Canonical(AL, eta, Lm, st)


def mixed_canonical(A, assume_left_canonical=False, want_status=True):
    """``iMPS([A]).mixed() -> (AL, AR, C)`` for a batch (qmps/tools.py:184-186,
    tests/test_represent.py:18-31).  Returns ``Mixed(AL, AR, C, eta, status)``."""
    A = _cdev(A, A.dtype if isinstance(A, torch.Tensor) and A.dtype in _CDT else torch.complex128)
    N, d, D, _ = A.shape
    AL = A if assume_left_canonical else torch.empty_like(A)
    AR = torch.empty_like(A)
    C = torch.empty((N, D, D), dtype=A.dtype, device=A.device)
    eta = torch.empty((N,), dtype=A.dtype, device=A.device)
    st = torch.empty((N,), dtype=torch.int32, device=A.device) if want_status else None
    with torch.cuda.device(A.device):
        L.check(L.load().qmps_mixed_canonical(d, D, N, _p(A), int(bool(assume_left_canonical)),
                                              None if assume_left_canonical else _p(AL), _p(AR), _p(C), _p(eta), _p(st),
                                              _dt(A), _stream()), "mixed_canonical")
    return Mixed(AL, AR, C, eta, st)


def gauge_transform(A, X, kind, eta=None):
    """``kind='left'``: X = l (Hermitian PD), A' = L A L^-1 / sqrt|eta| with l = L^dagger L;
    ``kind='right'``: X = C lower triangular, A' = C^-1 A C."""
    A = _cdev(A, A.dtype if isinstance(A, torch.Tensor) and A.dtype in _CDT else torch.complex128)
    X = _cdev(X, A.dtype, A.device)
    N, d, D, _ = A.shape
    eta = None if eta is None else _cdev(eta, A.dtype, A.device)
    out = torch.empty_like(A)
    st = torch.empty((N,), dtype=torch.int32, device=A.device)
    with torch.cuda.device(A.device):
        L.check(L.load().qmps_gauge_transform(d, D, N, _p(A), _p(X), {"left": 0, "right": 1}[kind], _p(eta), _p(out),
                                              None, _p(st), _dt(A), _stream()), "gauge_transform")
    return out, st


def expectation_values(A, ops, r=None, lvec=None, eta=None, assume_left_canonical=True):
    """``iMPS([A]).Es(ops)`` for a batch (qmps/loschmidts/time_evo.py:144, tests/test_represent.py:37):
    out[n, o] = <ops[o]> on site tensors A[N, d, D, D]; ops [nops, d, d].

    ``assume_left_canonical`` (tensors that come from a unitary): the trace-1 right environment
    is solved here unless ``r`` is given.  Otherwise the right / left leading eigenvectors of
    E_AA are computed (or taken from ``r``, ``lvec``, ``eta``) and the general formula is used."""
    A = _cdev(A, A.dtype if isinstance(A, torch.Tensor) and A.dtype in _CDT else torch.complex128)
    ops = _cdev(np.asarray(ops) if not isinstance(ops, torch.Tensor) else ops, A.dtype, A.device)
    N, d, D, _ = A.shape
    if ops.dim() == 2:
        ops = ops[None]
    if assume_left_canonical:
        if r is None:
            r = env_exact(A=A, want_eta=False, want_C=False, want_status=False).r
        lvec = eta = None
    else:
        if r is None or eta is None:
            fp = fixed_point(A, A, want_costs=False, want_status=False)
            r, eta = fp.vec, fp.eta
        if lvec is None:
            lvec = fixed_point(A, A, left=True, want_costs=False, want_status=False).vec
        lvec, eta = _cdev(lvec, A.dtype, A.device), _cdev(eta, A.dtype, A.device)
    r = _cdev(r, A.dtype, A.device)
    out = torch.empty((N, ops.shape[0]), dtype=A.dtype, device=A.device)
    with torch.cuda.device(A.device):
        L.check(L.load().qmps_expectation(d, D, N, _p(A), _p(r), _p(lvec), _p(eta), ops.shape[0], _p(ops), _p(out),
                                          _dt(A), _stream()), "expectation")
    return out


def overlap(A, B):
    """``iMPS.overlap`` as the reference plots it: per-site fidelity |eta(E_AB)|^2
    for batches of tensors (broadcast like ``fixed_point``).  D <= 16: dense eigen-solve of the mixed transfer matrix;
    larger D: its leading eigenvalue by the power method on the tensor-core contraction (``overlap_power``)."""
    D = A.shape[-1]
    if D > 16:
        return overlap_power(A, B).fid
    return fixed_point(A, B, want_vec=False, want_status=False).fid


PowerOverlap = collections.namedtuple("PowerOverlap", "eta fid rate r iterations converged")


def overlap_power(A, B, tol=1e-10, chunk=32, max_iter=8192):
    """Leading eigenvalue eta of the mixed transfer matrix E_AB at LARGE bond dimension (D >= 64 runs on tcgen05) by
    the power method r <- sum_s A_s r B_s^dagger / |.|: the Loschmidt-echo / overlap step of a classical iMPS
    (``A_.overlap(A)``, ``Trajectory.loschmidts()``; qmps/loschmidts/time_evo.py:145, mps_loschmidts.py:20-22) where a
    dense D^2 x D^2 eigen-solve is out of reach.  Runs ``chunk`` applications per C-ABI call (``qmps_tm_power``) and
    stops when every problem's Rayleigh quotient moved by less than ``tol`` (relative) over a chunk (the exact-integer
    complex128 contraction is good to ~4e-12 per application, so tolerances below ~1e-11 cannot be met).
    Returns ``PowerOverlap(eta[N], fid = |eta|^2, rate = -log|eta|^2, r[N, D, D], iterations, converged[N])``."""
    A = _cdev(A, A.dtype if isinstance(A, torch.Tensor) and A.dtype in _CDT else torch.complex128)
    B = _cdev(B, A.dtype, A.device)
    if A.shape != B.shape:
        raise ValueError("overlap_power needs A and B of the same shape [N, d, D, D]")
    r, prev, it = None, None, 0
    conv = None
    while it < max_iter:
        r, ray = tm_power(A, B, chunk, r0=r)
        it += chunk
        if prev is not None:
            conv = (ray - prev).abs() <= tol * ray.abs().clamp_min(1e-300)
            if bool(conv.all()):
                break
        prev = ray
    if conv is None:
        conv = torch.zeros(ray.shape, dtype=torch.bool, device=ray.device)
    a2 = (ray.real ** 2 + ray.imag ** 2)
    return PowerOverlap(ray, a2, -torch.log(a2), r, it, conv)


# ---- a13 -----------------------------------------------------------------------------
def loschmidt_rate(t, g0, g1):
    """Exact TFIM Loschmidt rate function for a batch of times (qmps/loschmidts/exact_loschmidt.py)."""
    t = _rdev(t).reshape(-1)
    out = torch.empty_like(t)
    with torch.cuda.device(t.device):
        L.check(L.load().qmps_loschmidt_rate(t.numel(), _p(t), float(g0), float(g1), _p(out), _stream()),
                "loschmidt_rate")
    return out


# ---- cfg 5 -------------------------------------------------------------------------------
def tm_power(A, B, K, r0=None):
    """K normalised applications r <- sum_s A_s r B_s^dagger / |.|_F from r0 (default
    1/sqrt(D)).  A, B [N, d, D, D].  Returns (r_K [N, D, D], rayleigh [N])."""
    A = _cdev(A, A.dtype if isinstance(A, torch.Tensor) and A.dtype in _CDT else torch.complex128)
    B = _cdev(B, A.dtype, A.device)
    N, d, D, _ = A.shape
    if r0 is None:
        r = (torch.eye(D, dtype=A.dtype, device=A.device) / np.sqrt(D)).repeat(N, 1, 1).contiguous()
    else:
        r = _cdev(r0, A.dtype, A.device).clone()
    ray = torch.empty((N,), dtype=A.dtype, device=A.device)
    with torch.cuda.device(A.device):
        L.check(L.load().qmps_tm_power(d, D, N, _p(A), _p(B), _p(r), int(K), _p(ray), _dt(A), _stream()), "tm_power")
    return r, ray



def tm_apply(A, B, X):
    """One unnormalised application of the (mixed) transfer map, Y[N, D, D] = sum_s A_s X B_s^dagger (``qmps_tm_apply``)."""
    A = _cdev(A, A.dtype if isinstance(A, torch.Tensor) and A.dtype in _CDT else torch.complex128)
    B = _cdev(B, A.dtype, A.device)
    X = _cdev(X, A.dtype, A.device)
    N, d, D, _ = A.shape
    Y = torch.empty_like(X)
    with torch.cuda.device(A.device):
        L.check(L.load().qmps_tm_apply(d, D, N, _p(A), _p(B), _p(X), _p(Y), _dt(A), _stream()), "tm_apply")
    return Y


LoschmidtTrajectory = namedtuple("LoschmidtTrajectory", "theta step_cost echo")


def loschmidt_trajectory(program, theta0, W, n_steps, n_gen=8, npop=2048, sigma0=0.05, seed=0, n_bfgs=30, dtype=torch.complex128):
    """The reference's time-evolution loop (scripts/loschmidt.py:367-375) on the device in ONE C-ABI call
    (``qmps_loschmidt_trajectory``): theta_{t+1} = argmin_p obj(p, A(theta_t), W) by ``n_gen`` generations of a
    population search followed by ``n_bfgs`` BFGS iterations (batched finite-difference gradient and line search),
    everything fed back on the device, then the echo series |eta(E_{A_t A_0})|^2.
    Returns ``LoschmidtTrajectory(theta[n_steps+1, P], step_cost[n_steps], echo[n_steps+1])`` (CUDA tensors)."""
    theta0 = _rdev(torch.as_tensor(np.asarray(theta0, dtype=np.float64)) if not isinstance(theta0, torch.Tensor) else theta0).reshape(-1).contiguous()
    P = theta0.numel()
    dev = theta0.device
    Wd = _cdev(W, dtype, dev).reshape(4, 4).contiguous()
    rd = _RDT[dtype]
    traj = torch.empty((n_steps + 1, P), dtype=torch.float64, device=dev)
    cost = torch.empty((n_steps,), dtype=rd, device=dev)
    echo = torch.empty((n_steps + 1,), dtype=rd, device=dev)
    ops = program.c_ops()
    with torch.cuda.device(dev):
        L.check(L.load().qmps_loschmidt_trajectory(ops, len(program), program.nq, P, _p(theta0), _p(Wd), int(n_steps), int(n_gen),
                                                   int(npop), float(sigma0), int(seed), int(n_bfgs), _p(traj), _p(cost), _p(echo), _CDT[dtype],
                                                   _stream()), "loschmidt_trajectory")
    return LoschmidtTrajectory(traj, cost, echo)


# ---- classical iTDVP ------------------------------------------------------------
def tdvp_dadt(A, h, imaginary=False, assume_left_canonical=False, want_status=False):
    """``iMPS([A]).dA_dt([h])`` for a batch A[N, d, D, D] (xmps; call sites scripts/classical_time_evolution.py:22-26,
    scripts/mixed_environment.py:41): the gauge-fixed TDVP tangent vector for the two-site Hamiltonian h[d*d, d*d].
    Returns (dA[N, d, D, D], energy[N]) (+ status)."""
    A = _cdev(A, A.dtype if isinstance(A, torch.Tensor) and A.dtype in _CDT else torch.complex128)
    N, d, D, _ = A.shape
    if D > 16:                                   # large bond dimension: composed from the tensor-core contraction kernels
        if assume_left_canonical:
            dA, e, _ = tdvp_tangent_large(A, h, imaginary=imaginary)
        else:
            dA, e = tdvp_dadt_large(A, h, imaginary=imaginary)
        st = torch.zeros((N,), dtype=torch.int32, device=A.device)
        return (dA, e, st) if want_status else (dA, e)
    hd = _cdev(h, A.dtype, A.device).reshape(d * d, d * d).contiguous()
    dA = torch.empty_like(A)
    e = torch.empty((N,), dtype=_RDT[A.dtype], device=A.device)
    st = torch.empty((N,), dtype=torch.int32, device=A.device) if want_status else None
    fn = L.load().qmps_tdvp_tangent if assume_left_canonical else L.load().qmps_tdvp_dadt
    with torch.cuda.device(A.device):
        L.check(fn(d, D, N, _p(A), _p(hd), int(bool(imaginary)), _p(dA), _p(e), _p(st), _dt(A), _stream()), "tdvp_dadt")
    return (dA, e, st) if want_status else (dA, e)


def tdvp_tangent_large(AL, h, imaginary=False, tol=1e-11, chunk=32, max_iter=4096):
    """The TDVP tangent vector of LEFT-CANONICAL tensors at LARGE bond dimension (D a multiple of 64, complex128):
    ``iMPS([A]).dA_dt([h])`` (scripts/classical_time_evolution.py:22-26) where the D^2 x D^2 systems of
    ``qmps_tdvp_tangent`` (D <= 16) are out of reach.  Same formulas as ``csrc/tdvp.cuh`` / ``oracle/tdvp.py``, with every
    O(D^3) contraction on the tcgen05 ``kind::i8`` kernels (``qmps_tm_power`` for the fixed point r, ``qmps_zgemm_c128_i8``
    for the products) and the two linear solves as iterations of the transfer map:
        r          power method (chunks of ``chunk`` applications until the Rayleigh quotient is stationary to ``tol``)
        K          the Neumann series K = sum_n E_L^n (H_l - e), E_L(K) = sum_s A_s^dagger K A_s (the right-hand side has no
                   component along the fixed point, so the series converges like |lambda_2|^n); stopped at ``tol``.
    This is a HOST-LEVEL COMPOSITION: the O(D^2) glue (transposes, axpy, traces, the d^2 x d^2 contraction with h) and the
    D x D inverse of r (``torch.linalg.inv``, a library call) go through torch; the tensor-core kernels do the rest.
    Returns (dA[N, d, D, D], energy[N], info) with info = dict(r_iterations, k_iterations, r)."""
    AL = _cdev(AL, torch.complex128)
    N, d, D, _ = AL.shape
    if D % 64:
        raise ValueError("tdvp_tangent_large needs D % 64 == 0 (use tdvp_dadt for D <= 16)")
    dev = AL.device
    hd = _cdev(h, torch.complex128, dev).reshape(d, d, d, d)
    T_ = lambda X: X.transpose(-1, -2).contiguous()                         # noqa: E731
    H_ = lambda X: X.conj().transpose(-1, -2).contiguous()                   # noqa: E731

    def mm(P, Q):                                                            # P . Q, batched over the leading dimensions
        shp = P.shape[:-2]
        return zgemm_i8(P.reshape(-1, D, D).contiguous(), T_(Q).reshape(-1, D, D)).reshape(*shp, D, D)
    # fixed point r of r -> sum_s A_s r A_s^dagger (Hermitian, trace 1)
    r, prev, it_r = None, None, 0
    while it_r < max_iter:
        r, ray = tm_power(AL, AL, chunk, r0=r)
        it_r += chunk
        if prev is not None and bool(((ray - prev).abs() <= tol).all()):
            break
        prev = ray
    r = 0.5 * (r + H_(r))
    tr = torch.einsum("nii->n", r)
    r = r / tr[:, None, None]
    r = 0.5 * (r + H_(r))
    # two-site blocks and the left Hamiltonian
    As = AL[:, :, None].expand(N, d, d, D, D)
    At = AL[:, None, :].expand(N, d, d, D, D)
    AA = mm(As, At)                                                          # AA[s, t] = A_s A_t
    C = torch.einsum("abcd,ncdik->nabik", hd, AA).contiguous()              # C[a, b] = sum h[(a,b),(c,d)] AA[c, d]
    Hl = mm(H_(AA), C).sum(dim=(1, 2))                                       # sum_st AA_st^dagger C_st
    e = torch.einsum("nik,nki->n", Hl, r).real
    eye = torch.eye(D, dtype=torch.complex128, device=dev)
    Bm = Hl - e[:, None, None].to(torch.complex128) * eye
    AH = H_(AL)                                                              # A_s^dagger
    K, it_k = Bm.clone(), 0
    while it_k < max_iter:
        Kn = Bm + tm_apply(AH, AH, K)                                        # sum_s A_s^dagger K A_s: one C-ABI call
        it_k += 1
        if it_k % 4:                                                         # the stopping test costs a device synchronisation:
            K = Kn                                                           # every fourth term only (extra terms only converge further)
            continue
        delta = (Kn - K).abs().amax(dim=(1, 2))
        K = Kn
        if bool((delta <= tol * K.abs().amax(dim=(1, 2)).clamp_min(1e-300)).all()):
            break
    K = K - torch.einsum("nik,nki->n", K, r)[:, None, None] * eye            # tr(K r) = 0
    rinv = torch.linalg.inv(r)
    # G^s = sum_t C^{st} r A_t^dagger r^-1 + sum_t A_t^dagger C^{ts} + K A_s
    Cr = mm(C, r[:, None, None].expand(N, d, d, D, D))                       # C^{st} r
    Ar = mm(AH, rinv[:, None].expand(N, d, D, D))                            # A_t^dagger r^-1
    G = mm(Cr, Ar[:, None].expand(N, d, d, D, D)).sum(dim=2)                 # sum_t (C^{st} r)(A_t^dagger r^-1)
    G = G + mm(AH[:, :, None].expand(N, d, d, D, D), C).sum(dim=1)           # sum_t A_t^dagger C^{ts}  (index [t, s] summed over t)
    G = G + mm(K[:, None].expand(N, d, D, D), AL)
    P = mm(AH, G).sum(dim=1)                                                 # sum_u A_u^dagger G^u
    f = -1.0 if imaginary else -1j
    dA = f * (G - mm(AL, P[:, None].expand(N, d, D, D)))
    return dA, e, dict(r_iterations=it_r, k_iterations=it_k, r=r)


def left_canonicalise_large(A, tol=1e-11, chunk=32, max_iter=4096):
    """``iMPS([A]).left_canonicalise()`` at large bond dimension (D % 64 == 0, complex128): l from the power method of the
    left transfer map on the tensor-core contraction, L = chol(l) (``torch.linalg.cholesky``), A_L = L A L^-1 / sqrt(eta)
    with the two products on ``qmps_zgemm_c128_i8``.  Returns (A_L, eta[N], L) as ``tdvp_canonical_parts`` of the oracle."""
    A = _cdev(A, torch.complex128)
    N, d, D, _ = A.shape
    AH = A.conj().transpose(-1, -2).contiguous()
    l, prev, it = None, None, 0
    while it < max_iter:
        l, ray = tm_power(AH, AH, chunk, r0=l)                              # l <- sum_s A_s^dagger l A_s
        it += chunk
        if prev is not None and bool(((ray - prev).abs() <= tol * ray.abs()).all()):
            break
        prev = ray
    eta = ray.real
    l = 0.5 * (l + l.conj().transpose(-1, -2))
    l = l / torch.einsum("nii->n", l).real[:, None, None] * D
    Lu = torch.linalg.cholesky(l).conj().transpose(-1, -2).contiguous()      # upper factor: l = L^dagger L
    Li = torch.linalg.inv(Lu)

    def mm(P, Q):
        shp = P.shape[:-2]
        return zgemm_i8(P.reshape(-1, D, D).contiguous(), Q.transpose(-1, -2).contiguous().reshape(-1, D, D)).reshape(*shp, D, D)
    AL = mm(mm(Lu[:, None].expand(N, d, D, D), A), Li[:, None].expand(N, d, D, D)) / torch.sqrt(eta)[:, None, None, None]
    return AL, eta, Lu


def tdvp_dadt_large(A, h, imaginary=False):
    """``iMPS([A]).dA_dt([h])`` in ANY gauge at large D: canonicalise, tangent in the left gauge, transform back
    (dA = sqrt(eta) L^-1 dA_L L; the left gauge condition is covariant).  Returns (dA, energy)."""
    AL, eta, Lu = left_canonicalise_large(A)
    dAL, e, _ = tdvp_tangent_large(AL, h, imaginary=imaginary)
    N, d, D, _ = AL.shape
    Li = torch.linalg.inv(Lu)

    def mm(P, Q):
        shp = P.shape[:-2]
        return zgemm_i8(P.reshape(-1, D, D).contiguous(), Q.transpose(-1, -2).contiguous().reshape(-1, D, D)).reshape(*shp, D, D)
    dA = mm(mm(Li[:, None].expand(N, d, D, D), dAL), Lu[:, None].expand(N, d, D, D)) * torch.sqrt(eta)[:, None, None, None]
    return dA, e


def tdvp_evolve_large(A, h, dt, n_steps, imaginary=False):
    """The reference's RK4 loop (scripts/classical_time_evolution.py:21-27: four ``dA_dt``, ``left_canonicalise`` after
    the step) at large bond dimension, with ``Trajectory.loschmidts()`` (-log|eta(E_{A_t A_0})|^2 by ``overlap_power``)
    and the energy per bond at the start of every step.  Returns ``TdvpRun(A_final, traj, rates, energy)``."""
    A0 = left_canonicalise_large(A)[0]
    cur = A0
    traj, energy = [A0], []
    for _ in range(int(n_steps)):
        k1, e = tdvp_dadt_large(cur, h, imaginary)
        k2, _ = tdvp_dadt_large(cur + (dt / 2) * k1, h, imaginary)
        k3, _ = tdvp_dadt_large(cur + (dt / 2) * k2, h, imaginary)
        k4, _ = tdvp_dadt_large(cur + dt * k3, h, imaginary)
        cur = left_canonicalise_large(cur + (dt / 6) * (k1 + 2 * k2 + 2 * k3 + k4))[0]
        traj.append(cur)
        energy.append(e)
    T = torch.stack(traj)
    rates = torch.stack([overlap_power(a, A0).rate for a in traj])
    return TdvpRun(cur, T, rates, torch.stack(energy) if energy else None, None)


def tdvp_evolve(A, h, dt, n_steps, method="rk4", imaginary=False, want_traj=False, want_rates=True, want_energy=True):
    """``n_steps`` TDVP steps of size ``dt`` for a batch of states, one C-ABI call, no host round trip:
    ``method='rk4'`` is the reference's loop (scripts/classical_time_evolution.py:21-27), ``'euler'`` is
    ``Trajectory.eulerint`` (qmps/loschmidts/mps_loschmidts.py:22).  Returns ``TdvpRun(A_final, traj, rates, energy)``:
    rates[t, n] = -log|eta(E_{A_t A_0})|^2 (``Trajectory.loschmidts()``), energy[t, n] at the start of step t."""
    if A.shape[-1] > 16:                         # large bond dimension: the composition on the tensor-core contraction kernels
        if method != "rk4":
            raise NotImplementedError("tdvp_evolve at D > 16 implements the reference's RK4 loop only")
        return tdvp_evolve_large(A, h, dt, n_steps, imaginary=imaginary)
    A = _cdev(A, A.dtype if isinstance(A, torch.Tensor) and A.dtype in _CDT else torch.complex128).clone()
    N, d, D, _ = A.shape
    hd = _cdev(h, A.dtype, A.device).reshape(d * d, d * d).contiguous()
    rd = _RDT[A.dtype]
    traj = torch.empty((n_steps + 1, N, d, D, D), dtype=A.dtype, device=A.device) if want_traj else None
    rates = torch.empty((n_steps + 1, N), dtype=rd, device=A.device) if want_rates else None
    en = torch.empty((n_steps, N), dtype=rd, device=A.device) if want_energy else None
    st = torch.zeros((N,), dtype=torch.int32, device=A.device)
    with torch.cuda.device(A.device):
        L.check(L.load().qmps_tdvp_evolve(d, D, N, _p(A), _p(hd), float(dt), int(n_steps), 1 if method == "rk4" else 0,
                                          int(bool(imaginary)), _p(traj), _p(rates), _p(en), _p(st), _dt(A), _stream()),
                "tdvp_evolve")
    return TdvpRun(A, traj, rates, en, st)



def zgemm_i8(X, Y, conj_y=False):
    """C[b] = X[b] . Y[b]^T (or ^H) in complex128 on tcgen05 kind::i8 (exact int8 slice products, FP64 recombination):
    X[batch, M, K], Y[batch, N, K] -> C[batch, M, N]; M % 64 == 0, N % 32 == 0, K % 64 == 0 (``qmps_zgemm_c128_i8``)."""
    X = _cdev(X, torch.complex128)
    Y = _cdev(Y, torch.complex128, X.device)
    batch, M, K = X.shape
    N = Y.shape[1]
    C = torch.empty((batch, M, N), dtype=torch.complex128, device=X.device)
    with torch.cuda.device(X.device):
        L.check(L.load().qmps_zgemm_c128_i8(batch, M, N, K, _p(X), _p(Y), int(bool(conj_y)), _p(C), _stream()), "zgemm_c128_i8")
    return C


# ---- (e) -----------------------------------------------------------------------------
def argmin(cost, index_offset=0):
    """(min, argmin + index_offset) of a float64 CUDA vector, as 1-element CUDA tensors."""
    cost = _rdev(cost).reshape(-1)
    bc = torch.empty((1,), dtype=torch.float64, device=cost.device)
    bi = torch.empty((1,), dtype=torch.int64, device=cost.device)
    with torch.cuda.device(cost.device):
        L.check(L.load().qmps_argmin(cost.numel(), _p(cost), int(index_offset), _p(bc), _p(bi), _stream()), "argmin")
    return bc, bi
