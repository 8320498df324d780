"""Two-layer brick-wall iMPS family: device mirror of the contraction classes of
``new_tdvp/ClassicalTDVPStripped.py`` -- ``RightEnvironment`` / ``LeftEnvironment`` (:312-426),
``OverlapCalculator`` (:428-586), ``ManifoldOverlap`` (:228-309), ``Represent.exact_env`` (:657-660)
and the body of ``Evolve.exact_cost_function`` (:777-790).

Batched functions (``bw_*``) take CUDA tensors / numpy arrays of 4x4 unitaries ``[N, 4, 4]`` (or a
single ``[4, 4]`` shared by the batch) and go through the C ABI (``qmps_bw_*``); the classes keep the
reference's names and call signatures -- ``(2,2,2,2)`` views, ``U1_``/``U2_`` already daggered by the
caller -- as a batch of one.  ``paramU`` (:159-180) needs ``xmps.spin.U4`` (not vendored), so the
parameter -> unitary step is outside this module: candidates are passed as unitaries.  No CPU fallback.
"""
from collections import namedtuple

import numpy as np
import torch

from . import _lib as L
from .batched import _cdev, _p, _stream, _dt, _RDT

BwEnvironment = namedtuple("BwEnvironment", "mat eta vec status")
BwCost = namedtuple("BwCost", "cost overlap eta Mr status")


def _mats(x, n, dtype, device=None):
    """[..., n, n] (or the reference's (2,)*k views) -> (contiguous [count, n, n] CUDA tensor, count)."""
    if not isinstance(x, torch.Tensor):
        x = np.asarray(x)
    x = x.reshape(-1, n, n)
    x = _cdev(x, dtype, device)
    return x, x.shape[0]


def _batch(*counts):
    if 0 in counts:                      # an empty per-problem array: empty batch
        return 0
    N = max(counts)
    for c in counts:
        if c not in (1, N):
            raise ValueError(f"batch sizes must be 1 or {N}, got {counts}")
    return N


def _ctype(*xs):
    for x in xs:
        if isinstance(x, torch.Tensor) and x.dtype == torch.complex64:
            return torch.complex64
    return torch.complex128


def bw_environment(U1, U2, U1_, U2_, side="right", bra_undaggered=False, want_mat=True, want_vec=True):
    """``exact_environment_circuit`` + ``exact_environment`` for a batch: the 4x4 map, its eigenvalue
    chosen by numpy's complex argmax and the eigenvector ``[N, 2, 2]`` in ``scipy.linalg.eig``'s gauge."""
    dt = _ctype(U1, U2, U1_, U2_)
    U1, nk = _mats(U1, 4, dt)
    U2, nk2 = _mats(U2, 4, dt, U1.device)
    B1, nb = _mats(U1_, 4, dt, U1.device)
    B2, nb2 = _mats(U2_, 4, dt, U1.device)
    if nk != nk2 or nb != nb2:
        raise ValueError("U1/U2 (and U1_/U2_) must have the same batch size")
    N = _batch(nk, nb)
    mat = torch.empty((N, 4, 4), dtype=dt, device=U1.device) if want_mat else None
    eta = torch.empty((N,), dtype=dt, device=U1.device)
    vec = torch.empty((N, 2, 2), dtype=dt, device=U1.device) if want_vec else None
    st = torch.empty((N,), dtype=torch.int32, device=U1.device)
    with torch.cuda.device(U1.device):
        L.check(L.load().qmps_bw_environment({"right": 0, "left": 1}[side], N, nk, _p(U1), _p(U2), nb, _p(B1), _p(B2),
                                             int(bool(bra_undaggered)), _p(mat), _p(eta), _p(vec), _p(st), _dt(U1),
                                             _stream()), "bw_environment")
    return BwEnvironment(mat, eta, vec, st)


def bw_env_apply(U1, U2, U1_, U2_, M, bra_undaggered=False):
    """``RightEnvironment.circuit``: one application of the right map to ``M [.., 2, 2]``."""
    dt = _ctype(U1, U2, U1_, U2_, M)
    U1, nk = _mats(U1, 4, dt)
    U2, _ = _mats(U2, 4, dt, U1.device)
    B1, nb = _mats(U1_, 4, dt, U1.device)
    B2, _ = _mats(U2_, 4, dt, U1.device)
    M, nm = _mats(M, 2, dt, U1.device)
    N = _batch(nk, nb, nm)
    out = torch.empty((N, 2, 2), dtype=dt, device=U1.device)
    with torch.cuda.device(U1.device):
        L.check(L.load().qmps_bw_env_apply(N, nk, _p(U1), _p(U2), nb, _p(B1), _p(B2), int(bool(bra_undaggered)), nm,
                                           _p(M), _p(out), _dt(U1), _stream()), "bw_env_apply")
    return out


def bw_expectation(U1, U2, O):
    """``OverlapCalculator.expectation_value``: ``O`` is ``[.., 4, 4]`` (2-qubit) or ``[.., 16, 16]``
    (4-qubit), also accepted as the reference's ``(2,)*4`` / ``(2,)*8`` views of ONE operator."""
    dt = _ctype(U1, U2, O)
    shape = tuple(O.shape)
    if shape in ((2,) * 4, (2,) * 8):
        n = 4 if len(shape) == 4 else 16
    else:
        n = shape[-1]
    if n not in (4, 16):
        raise ValueError("operator must act on 2 or 4 qubits")
    U1, nk = _mats(U1, 4, dt)
    U2, _ = _mats(U2, 4, dt, U1.device)
    O, no = _mats(O, n, dt, U1.device)
    N = _batch(nk, no)
    out = torch.empty((N,), dtype=_RDT[dt], device=U1.device)
    with torch.cuda.device(U1.device):
        L.check(L.load().qmps_bw_expectation(N, nk, _p(U1), _p(U2), 2 if n == 4 else 4, no, _p(O), _p(out), _dt(U1),
                                             _stream()), "bw_expectation")
    return out


def bw_overlap(U1, U2, U1_, U2_, Mr, Ml, W, bra_undaggered=False):
    """``ManifoldOverlap.circuit``: <phi| Ml (x) W (x) Mr |psi>, complex ``[N]``."""
    dt = _ctype(U1, U2, U1_, U2_, Mr, Ml, W)
    U1, nk = _mats(U1, 4, dt)
    U2, _ = _mats(U2, 4, dt, U1.device)
    B1, nb = _mats(U1_, 4, dt, U1.device)
    B2, _ = _mats(U2_, 4, dt, U1.device)
    Mr, nm = _mats(Mr, 2, dt, U1.device)
    Ml, nm2 = _mats(Ml, 2, dt, U1.device)
    if nm != nm2:
        raise ValueError("Mr and Ml must have the same batch size")
    W, nw = _mats(W, 16, dt, U1.device)
    N = _batch(nk, nb, nm, nw)
    out = torch.empty((N,), dtype=dt, device=U1.device)
    with torch.cuda.device(U1.device):
        L.check(L.load().qmps_bw_overlap(N, nk, _p(U1), _p(U2), nb, _p(B1), _p(B2), int(bool(bra_undaggered)), nm,
                                         _p(Mr), _p(Ml), nw, _p(W), _p(out), _dt(U1), _stream()), "bw_overlap")
    return out


def bw_evolve_cost(U1, U2, V1, V2, W, want_all=False):
    """Body of ``Evolve.exact_cost_function`` for candidate unitaries ``V1, V2 [N, 4, 4]`` (what
    ``paramU`` returns, undaggered): exact right environment of the mixed map -> overlap -> ``-|.|^2``
    in one launch.  Returns the cost ``[N]`` or ``BwCost(cost, overlap, eta, Mr, status)``."""
    dt = _ctype(U1, U2, V1, V2, W)
    U1, nk = _mats(U1, 4, dt)
    U2, _ = _mats(U2, 4, dt, U1.device)
    V1, nb = _mats(V1, 4, dt, U1.device)
    V2, _ = _mats(V2, 4, dt, U1.device)
    W, nw = _mats(W, 16, dt, U1.device)
    N = _batch(nk, nb, nw)
    dev = U1.device
    cost = torch.empty((N,), dtype=_RDT[dt], device=dev)
    ov = eta = Mr = st = None
    if want_all:
        ov = torch.empty((N,), dtype=dt, device=dev)
        eta = torch.empty((N,), dtype=dt, device=dev)
        Mr = torch.empty((N, 2, 2), dtype=dt, device=dev)
        st = torch.empty((N,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().qmps_bw_evolve_cost(N, nk, _p(U1), _p(U2), nb, _p(V1), _p(V2), nw, _p(W), _p(cost), _p(ov),
                                             _p(eta), _p(Mr), _p(st), _dt(U1), _stream()), "bw_evolve_cost")
    return BwCost(cost, ov, eta, Mr, st) if want_all else cost


# ---- the reference's classes, batch of one ------------------------------------------------------
def _np(t):
    return t.detach().cpu().numpy()


class RightEnvironment:
    """``ClassicalTDVPStripped.RightEnvironment`` (:355-426)."""
    _side = "right"

    def exact_environment_circuit(self, U1, U2, U1_, U2_):
        return _np(bw_environment(U1, U2, U1_, U2_, side=self._side, want_vec=False).mat)[0]

    def exact_environment(self, U1, U2, U1_, U2_):
        res = bw_environment(U1, U2, U1_, U2_, side=self._side, want_mat=False)
        return complex(_np(res.eta)[0]), _np(res.vec)[0]

    def circuit(self, U1, U2, U1_, U2_, M, path=None):
        return _np(bw_env_apply(U1, U2, U1_, U2_, M))[0]


class LeftEnvironment(RightEnvironment):
    """``ClassicalTDVPStripped.LeftEnvironment`` (:312-352); the reference defines no ``circuit`` here."""
    _side = "left"

    def circuit(self, *a, **k):
        raise AttributeError("LeftEnvironment has no circuit() in the reference")


class OverlapCalculator:
    """``ClassicalTDVPStripped.OverlapCalculator`` (:428-586)."""

    def expectation_value(self, U1, U2, O, path=None):
        return float(_np(bw_expectation(U1, U2, O))[0])

    def mexpectation_value(self, U1, U2, O):
        return self.expectation_value(U1, U2, O)


class ManifoldOverlap:
    """``ClassicalTDVPStripped.ManifoldOverlap`` (:228-309)."""

    def circuit(self, U1, U2, U1_, U2_, Mr, Ml, W, path=None):
        return complex(_np(bw_overlap(U1, U2, U1_, U2_, Mr, Ml, W))[0])

    def mcircuit(self, U1, U2, U1_, U2_, Mr, Ml, W):
        return self.circuit(U1, U2, U1_, U2_, Mr, Ml, W)


class Represent:
    """``ClassicalTDVPStripped.Represent.exact_env`` (:657-660); the variational search (:609-655) is a
    scipy.optimize driver and stays with the caller."""

    def __init__(self):
        self.RE, self.LE = RightEnvironment(), LeftEnvironment()

    def exact_env(self, U1, U2, U1_, U2_):
        _, Mr = self.RE.exact_environment(U1, U2, U1_, U2_)
        _, Ml = self.LE.exact_environment(U1, U2, U1_, U2_)
        return Mr, Ml


class Evolve:
    """Cost function of ``ClassicalTDVPStripped.Evolve`` (:777-790) on candidate unitaries."""

    def __init__(self, W=None, U1=None, U2=None):
        self.W, self.U1, self.U2 = W, U1, U2

    def exact_cost_function_unitaries(self, V1, V2):
        """``exact_cost_function`` after ``paramU``: V1, V2 = ``paramU(params)`` (one pair -> float, a
        batch ``[N, 4, 4]`` -> numpy array)."""
        c = _np(bw_evolve_cost(self.U1, self.U2, V1, V2, self.W))
        shape = tuple(V1.shape)
        return float(c[0]) if len(shape) == 2 or shape == (2, 2, 2, 2) else c
