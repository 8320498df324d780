"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the two-layer brick-wall iMPS family of
``new_tdvp/ClassicalTDVPStripped.py``: environments, expectation values and the
manifold overlap behind ``Evolve.exact_cost_function``.

The reference writes every quantity as an ``np.einsum`` over ``(2,2,2,2)`` views of the 4x4
unitaries.  This restatement uses state vectors and 4x4 matrix algebra instead (the closed forms are
derived in the docstrings), so agreement with the reference's own outputs
(``tests/golden/ref_brickwall.npz``, made by ``oracle/make_golden_bw.py`` from the unmodified
reference code) is a real check.  PINNED: every function here against that file.

Conventions: ``U[(a,b),(c,d)] = U.reshape(2,2,2,2)[a,b,c,d]``; ``u2 = U2[:,0].reshape(2,2)`` is the
column the circuit reaches from |00>, ``w2 = U2_[0,:].reshape(2,2)`` the matching bra row;
``U1_``, ``U2_`` are what the reference passes as ``U1_``/``U2_`` -- already daggered by the caller.
"""
import numpy as np
from scipy.linalg import eig

__all__ = ["bw_right_env_matrix", "bw_left_env_matrix", "bw_exact_environment", "bw_right_env_circuit",
           "bw_state", "bw_bra", "bw_expectation", "bw_manifold_overlap", "bw_exact_cost", "lapack_gauge",
           "numpy_argmax_complex"]


def _m(U):
    return np.asarray(U, dtype=complex).reshape(4, 4)


def bw_right_env_matrix(U1, U2, U1_, U2_):
    """``RightEnvironment.exact_environment_circuit`` (ClassicalTDVPStripped.py:394-417).

    ``M[(a,b),(c,e)] = sum_{x,y} P[(b,y),(a,x)] u2[x,c] w2[y,e]``,  ``P = U1_ @ U1``."""
    P = (_m(U1_) @ _m(U1)).reshape(2, 2, 2, 2)            # [b, y, a, x]
    u2 = _m(U2)[:, 0].reshape(2, 2)
    w2 = _m(U2_)[0, :].reshape(2, 2)
    return np.einsum("byax,xc,ye->abce", P, u2, w2).reshape(4, 4)


def bw_left_env_matrix(U1, U2, U1_, U2_):
    """``LeftEnvironment.exact_environment_circuit`` (:322-345).

    ``M[(a,b),(c,e)] = sum_{x,y} u2[c,x] w2[e,y] P[(y,b),(x,a)]``."""
    P = (_m(U1_) @ _m(U1)).reshape(2, 2, 2, 2)            # [y, b, x, a]
    u2 = _m(U2)[:, 0].reshape(2, 2)
    w2 = _m(U2_)[0, :].reshape(2, 2)
    return np.einsum("ybxa,cx,ey->abce", P, u2, w2).reshape(4, 4)


def numpy_argmax_complex(w):
    """``np.argmax`` on a complex array orders lexicographically by (real, imag) -- the selection
    rule the reference applies to the eigenvalues (:351, :423), NOT the largest modulus."""
    best = 0
    for k in range(1, len(w)):
        if (w[k].real, w[k].imag) > (w[best].real, w[best].imag):
            best = k
    return best


def lapack_gauge(v):
    """zgeev's eigenvector normalisation (what ``scipy.linalg.eig`` returns): unit 2-norm and the
    component of largest modulus real positive."""
    v = np.asarray(v, dtype=complex).ravel()
    v = v / np.linalg.norm(v)
    k = int(np.argmax(v.real ** 2 + v.imag ** 2))
    return v * (np.conj(v[k]) / abs(v[k]))


def bw_exact_environment(M):
    """``exact_environment`` (:347-352, :419-426): ``eig`` of the 4x4 map, ``argmax`` eigenvalue,
    eigenvector reshaped (2,2)."""
    w, v = eig(np.asarray(M, dtype=complex))
    k = numpy_argmax_complex(w)
    return w[k], lapack_gauge(v[:, k]).reshape(2, 2)


def bw_right_env_circuit(U1, U2, U1_, U2_, M):
    """``RightEnvironment.circuit`` (:360-384): one application of the right map to ``M``.

    ``out[i,j] = sum_{x,y,z,w} w2[y,z] P[(i,y),(j,x)] M[z,w] u2[x,w]``."""
    P = (_m(U1_) @ _m(U1)).reshape(2, 2, 2, 2)
    u2 = _m(U2)[:, 0].reshape(2, 2)
    w2 = _m(U2_)[0, :].reshape(2, 2)
    return np.einsum("yz,iyjx,zw,xw->ij", w2, P, np.asarray(M, dtype=complex), u2)


def _kron(*ms):
    out = np.array([[1.0 + 0j]])
    for m in ms:
        out = np.kron(out, m)
    return out


def bw_state(U1, U2, cells):
    """|psi> = (1 (x) U1^{(x)(cells-1)} (x) 1) (U2^{(x)cells}) |0...0> on 2*cells qubits, big-endian
    (the ket half of :248-266, :462-484, :515-533)."""
    col = _m(U2)[:, 0]
    psi = np.array([1.0 + 0j])
    for _ in range(cells):
        psi = np.kron(psi, col)
    mid = _kron(np.eye(2), *([_m(U1)] * (cells - 1)), np.eye(2))
    return mid @ psi


def bw_bra(U1_, U2_, cells):
    """<phi| = <0...0| (U2_^{(x)cells}) (1 (x) U1_^{(x)(cells-1)} (x) 1) as a row vector."""
    row = _m(U2_)[0, :]
    phi = np.array([1.0 + 0j])
    for _ in range(cells):
        phi = np.kron(phi, row)
    mid = _kron(np.eye(2), *([_m(U1_)] * (cells - 1)), np.eye(2))
    return phi @ mid


def bw_expectation(U1, U2, O):
    """``OverlapCalculator.expectation_value`` (:428-533): <psi| 1 (x) O (x) 1 |psi>.real with a
    2-qubit (4x4) operator on 4 qubits or a 4-qubit (16x16) operator on 6 qubits."""
    O = np.asarray(O, dtype=complex)
    n = int(round(np.sqrt(O.size)))
    O = O.reshape(n, n)
    cells = 2 if n == 4 else 3
    psi = bw_state(U1, U2, cells)
    return float(np.real(np.vdot(psi, _kron(np.eye(2), O, np.eye(2)) @ psi)))


def bw_manifold_overlap(U1, U2, U1_, U2_, Mr, Ml, W):
    """``ManifoldOverlap.circuit`` (:228-268): <phi| Ml (x) W (x) Mr |psi> on 6 qubits."""
    psi = bw_state(U1, U2, 3)
    phi = bw_bra(U1_, U2_, 3)
    mid = _kron(np.asarray(Ml, dtype=complex), np.asarray(W, dtype=complex).reshape(16, 16), np.asarray(Mr, dtype=complex))
    return complex(phi @ (mid @ psi))


def bw_exact_cost(U1, U2, V1, V2, W):
    """Body of ``Evolve.exact_cost_function`` (:777-790) for a candidate ``(V1, V2)`` (the
    UNdaggered unitaries ``paramU`` returns): right environment of the mixed map, overlap with
    ``(Mr, Mr^dagger)``, ``-|overlap|^2``.  Returns (cost, overlap, eta, Mr)."""
    U1_, U2_ = _m(V1).conj().T, _m(V2).conj().T
    eta, Mr = bw_exact_environment(bw_right_env_matrix(U1, U2, U1_, U2_))
    ov = bw_manifold_overlap(U1, U2, U1_, U2_, Mr, Mr.conj().T, W)
    return -abs(ov) ** 2, ov, eta, Mr
