"""Drop-in mirror of ``qmps.time_evolve_tools`` (classical part).

Reference: qmps/time_evolve_tools.py:20-91.  numpy in / numpy out, batch of one
through the C ABI; see ``qmps_b200.batched`` for the batched forms.
"""
import numpy as np

from . import batched
from .represent import ShallowFullStateTensor, StateGate, unitary
from .tools import unitary_to_tensor

__all__ = ["merge", "put_env_on_left_site", "get_env_off_left_site", "put_env_on_right_site",
           "get_env_off_right_site", "gate", "egate", "get_overlap_exact", "right_fixed_point",
           "left_fixed_point"]

_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


def merge(A, B):
    """-A- -B- -> -AB-: M[(s1,s2), i, j] = (A^s1 B^s2)[i, j] (qmps/time_evolve_tools.py:20-23).
    The reference reshapes to bond dimension 2 unconditionally; this keeps the general
    (d1*d2, D, D) shape, which is identical at D = 2."""
    A = np.ascontiguousarray(np.asarray(A, dtype=np.complex128))
    B = np.ascontiguousarray(np.asarray(B, dtype=np.complex128))
    return batched.merge(A[None], B[None]).cpu().numpy()[0]


def _rows_to_unitary(rows):
    """two orthonormal rows (2 x 4) -> 4 x 4 unitary with those first two rows."""
    Q = rows.conj().T                                  # 4 x 2 isometry
    A = np.ascontiguousarray(Q.reshape(2, 2, 2).transpose(1, 0, 2))
    return batched.tensor_to_unitary(A[None]).cpu().numpy()[0].conj().T


def put_env_on_left_site(q, ret_n=False):
    """Embed q^T/|q| in a 4x4 unitary acting on the left site (qmps/time_evolve_tools.py:38-53)."""
    a, b, c, d = np.asarray(q, dtype=np.complex128).T.reshape(-1)
    n = np.sqrt(abs(a) ** 2 + abs(b) ** 2 + abs(c) ** 2 + abs(d) ** 2)
    rows = np.array([[a, np.conj(c), b, np.conj(d)], [c, -np.conj(a), d, -np.conj(b)]]) / n
    A = _SWAP @ _rows_to_unitary(rows)
    return (A, n) if ret_n else A


def get_env_off_left_site(A):
    """qmps/time_evolve_tools.py:55-57."""
    return np.asarray(A).reshape(2, 2, 2, 2)[:, 0, :, 0].T


def put_env_on_right_site(q, ret_n=False):
    """qmps/time_evolve_tools.py:59-70."""
    a, b, c, d = np.asarray(q, dtype=np.complex128).reshape(-1)
    n = np.sqrt(abs(a) ** 2 + abs(b) ** 2 + abs(c) ** 2 + abs(d) ** 2)
    rows = np.array([[a, b, np.conj(d), -np.conj(c)], [c, d, -np.conj(b), np.conj(a)]]) / n
    A = _rows_to_unitary(rows)
    return (A, n) if ret_n else A


def get_env_off_right_site(A):
    """qmps/time_evolve_tools.py:72-74."""
    return np.asarray(A).reshape(2, 2, 2, 2)[0, :, 0, :]


def gate(v, symbol="U"):
    """qmps/time_evolve_tools.py:76-80."""
    return ShallowFullStateTensor(2, v, symbol)


def egate(v, symbol="R"):
    return StateGate(v, symbol)


def right_fixed_point(A, B):
    """``Map(A, B).right_fixed_point()`` -> (x, r), r of unit Frobenius norm (xmps; call site
    qmps/time_evolve_tools.py:87)."""
    A = np.ascontiguousarray(np.asarray(A, dtype=np.complex128))
    B = np.ascontiguousarray(np.asarray(B, dtype=np.complex128))
    fp = batched.fixed_point(A[None], B[None])
    return complex(fp.eta.cpu()[0]), fp.vec.cpu().numpy()[0]


def left_fixed_point(A, B):
    """``Map(A, B).left_fixed_point()`` -> (x, l) for the action l -> sum_s A_s^dagger l B_s."""
    A = np.ascontiguousarray(np.asarray(A, dtype=np.complex128))
    B = np.ascontiguousarray(np.asarray(B, dtype=np.complex128))
    fp = batched.fixed_point(A[None], B[None], left=True)
    return complex(fp.eta.cpu()[0]), fp.vec.cpu().numpy()[0]


def get_overlap_exact(p1, p2, gate=gate, testing=True):
    """|eta(E_AB)|^2 (and r) for two parameter vectors (qmps/time_evolve_tools.py:84-91).
    ``left_canonicalise`` of the reference is a pure gauge change on tensors that come
    from a unitary and leaves |eta| unchanged."""
    A = unitary_to_tensor(unitary(gate(p1)))
    B = unitary_to_tensor(unitary(gate(p2)))
    x, r = right_fixed_point(A, B)
    if testing:
        return np.abs(x) ** 2, r
    return np.abs(x) ** 2
