"""The slice of xmps' ``iMPS`` / ``Map`` / ``TransferMatrix`` API that the reference's loops
use around the environment solve, on the CUDA path.

xmps (github.com/fergusbarratt/xmps) is an un-vendored dependency of the reference; call sites:

* ``iMPS([A]).left_canonicalise()[0]``   qmps/time_evolve_tools.py:85-86, loschmidts/time_evo.py:76
* ``A_.Es(ops)``, ``A_.overlap(A)``      qmps/loschmidts/time_evo.py:143-145, scripts/loschmidt.py:368-370
* ``iMPS([A]).mixed() -> AL, AR, C``      qmps/tools.py:184-186, tests/test_represent.py:18
* ``TransferMatrix(A).eigs()``            qmps/tools.py:181, qmps/ground_state.py:296
* ``Map(A,B).right_fixed_point()`` etc.   qmps/time_evolve_tools.py:87, loschmidts/time_evo.py:79-82

Single-site unit cells only (every call site above uses one).  Batch-of-1 wrappers over
``qmps_b200.batched``; there is no CPU path.
"""
import numpy as np
from numpy.linalg import LinAlgError

from . import _lib, batched


def _np(t):
    return t.cpu().numpy()


def _raise_for(status, what):
    st = int(status.cpu()[0])
    if st == _lib.ST_NOT_PD:
        raise LinAlgError(f"{what}: fixed point is not positive definite")
    if st == _lib.ST_NO_CONVERGE:
        raise LinAlgError(f"{what}: eigenvalue iteration did not converge")
    if st == _lib.ST_SINGULAR:
        raise LinAlgError(f"{what}: degenerate leading eigenvalue")


class iMPS:
    """Uniform MPS with a one-site unit cell: ``iMPS([A])``, ``A`` of shape (d, D, D)."""

    def __init__(self, data=None):
        self.data = [np.ascontiguousarray(np.asarray(a, dtype=np.complex128)) for a in (data or [])]
        if len(self.data) > 1:
            raise NotImplementedError("qmps_b200.imps.iMPS supports one-site unit cells (as every hot-path call site uses)")

    # ---- container protocol used by the reference (``A_[0]``, ``A + dA``) ----
    def __getitem__(self, k):
        return self.data[k]

    def __len__(self):
        return len(self.data)

    def __add__(self, other):
        o = other.data[0] if isinstance(other, iMPS) else np.asarray(other)
        return iMPS([self.data[0] + o])

    def __rmul__(self, c):
        return iMPS([c * self.data[0]])

    @property
    def d(self):
        return self.data[0].shape[0]

    @property
    def D(self):
        return self.data[0].shape[1]

    def random(self, d, D, seed=None):
        rng = np.random.default_rng(seed)
        return iMPS([rng.normal(size=(d, D, D)) + 1j * rng.normal(size=(d, D, D))])

    # ---- SURVEY 8(f)-1 ----
    def left_canonicalise(self):
        res = batched.left_canonicalise(self.data[0][None])
        _raise_for(res.status, "left_canonicalise")
        return iMPS([_np(res.AL)[0]])

    def mixed(self):
        res = batched.mixed_canonical(self.data[0][None])
        _raise_for(res.status, "mixed")
        return iMPS([_np(res.AL)[0]]), iMPS([_np(res.AR)[0]]), _np(res.C)[0]

    def Es(self, ops):
        """Expectation values of single-site operators (real parts, as plotted by the reference)."""
        out = batched.expectation_values(self.data[0][None], np.asarray(ops, dtype=np.complex128),
                                         assume_left_canonical=False)
        return _np(out)[0].real

    def E(self, op):
        return float(self.Es([op])[0])

    def overlap(self, other):
        o = other.data[0] if isinstance(other, iMPS) else np.asarray(other)
        return float(_np(batched.overlap(self.data[0][None], np.ascontiguousarray(o, dtype=np.complex128)[None]))[0])

    # ---- TDVP (scripts/classical_time_evolution.py:22-26, scripts/mixed_environment.py:41) ----
    def __sub__(self, other):
        o = other.data[0] if isinstance(other, iMPS) else np.asarray(other)
        return iMPS([self.data[0] - o])

    def __mul__(self, c):
        return iMPS([c * self.data[0]])

    def __truediv__(self, c):
        return iMPS([self.data[0] / c])

    def dA_dt(self, H, imaginary=False):
        """``mps.dA_dt([H])``: the TDVP tangent vector as an ``iMPS`` (so that ``mps + k1/2`` works as in the
        reference's RK4 loop).  ``H`` is a list with one two-site Hamiltonian matrix."""
        h = H[0] if isinstance(H, (list, tuple)) else H
        dA, _, st = batched.tdvp_dadt(self.data[0][None], np.asarray(h, dtype=np.complex128), imaginary=imaginary,
                                      want_status=True)
        _raise_for(st, "dA_dt")
        return iMPS([_np(dA)[0]])

    def energy(self, H):
        h = H[0] if isinstance(H, (list, tuple)) else H
        return float(_np(batched.tdvp_dadt(self.data[0][None], np.asarray(h, dtype=np.complex128))[1])[0])


class Trajectory:
    """``xmps.iTDVP.Trajectory(mps_0=A, H=[h])`` as qmps/loschmidts/mps_loschmidts.py:20-22 uses it:
    ``eulerint(T)`` / ``rk4int(T)`` integrate over the time grid ``T`` on the device in one call, ``loschmidts()``
    returns -log|overlap with the initial state|^2 per site, ``mps_list()`` the stored states."""

    def __init__(self, mps_0, H):
        self.mps_0 = mps_0 if isinstance(mps_0, iMPS) else iMPS([mps_0])
        self.H = H if isinstance(H, (list, tuple)) else [H]
        self.run = None

    def _integrate(self, T, method):
        T = np.asarray(T, dtype=np.float64)
        if len(T) < 2:
            raise ValueError("need at least two time points")
        dt = float(T[1] - T[0])
        if not np.allclose(np.diff(T), dt):
            raise ValueError("uniform time grids only")
        self.run = batched.tdvp_evolve(self.mps_0.data[0][None], np.asarray(self.H[0], dtype=np.complex128), dt, len(T) - 1,
                                       method=method, want_traj=True)
        _raise_for(self.run.status, "Trajectory")
        return self

    def eulerint(self, T):
        return self._integrate(T, "euler")

    def rk4int(self, T):
        return self._integrate(T, "rk4")

    def loschmidts(self):
        return _np(self.run.rates)[:, 0]

    def mps_list(self):
        return [iMPS([a]) for a in _np(self.run.traj)[:, 0]]


class TransferMatrix:
    """``TransferMatrix(A).eigs() -> (eta, l, r)`` (qmps/tools.py:181): r Hermitian trace 1,
    l Hermitian with tr(l r) = 1."""

    def __init__(self, A):
        self.A = np.ascontiguousarray(np.asarray(A, dtype=np.complex128))

    def eigs(self):
        A = self.A[None]
        res = batched.env_exact(A=A, assume_left_canonical=False, want_C=False)
        _raise_for(res.status, "TransferMatrix.eigs")
        r = _np(res.r)[0]
        l = _np(batched.fixed_point(A, A, left=True, want_costs=False, gauge="trace").vec)[0]
        l = (l + l.conj().T) / 2
        l = l / np.trace(l @ r).real
        eta = complex(_np(res.eta)[0])
        return (eta.real if abs(eta.imag) < 1e-12 * max(1.0, abs(eta)) else eta), l, r


class Map:
    """``Map(A, B)``: the mixed transfer matrix E_AB (qmps/time_evolve_tools.py:87)."""

    def __init__(self, A, B):
        self.A = np.ascontiguousarray(np.asarray(A, dtype=np.complex128))
        self.B = np.ascontiguousarray(np.asarray(B, dtype=np.complex128))

    def _fp(self, left):
        fp = batched.fixed_point(self.A[None], self.B[None], left=left, want_costs=False)
        _raise_for(fp.status, "Map fixed point")
        return complex(_np(fp.eta)[0]), _np(fp.vec)[0]

    def right_fixed_point(self):
        return self._fp(False)

    def left_fixed_point(self):
        return self._fp(True)

    def asmatrix(self):
        D1, D2 = self.A.shape[1], self.B.shape[1]
        return np.einsum("sij,skl->ikjl", self.A, self.B.conj()).reshape(D1 * D2, D1 * D2)

    def is_right_eigenvector(self, r, tol=1e-8):
        v = np.asarray(r).reshape(-1)
        w = self.asmatrix() @ v
        lam = np.vdot(v, w) / np.vdot(v, v)
        return bool(np.allclose(w, lam * v, atol=tol))

    def is_left_eigenvector(self, l, tol=1e-8):
        v = np.asarray(l).reshape(-1)
        w = self.asmatrix().conj().T @ v
        lam = np.vdot(v, w) / np.vdot(v, v)
        return bool(np.allclose(w, lam * v, atol=tol))
