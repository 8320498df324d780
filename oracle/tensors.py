"""Oracle, part 1: tensor <-> unitary bookkeeping and transfer-matrix fixed points.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  numpy/scipy, complex128.
Citations are into ``/root/reference``.
"""
import numpy as np
import scipy.linalg as sla

__all__ = [
    "cT", "direct_sum", "from_real_vector", "to_real_vector",
    "unitary_to_tensor", "unitary_extension", "tensor_to_unitary",
    "environment_to_unitary", "environment_from_unitary",
    "transfer_matrix", "leading_eig", "eigs", "right_fixed_point",
    "left_fixed_point", "env_exact_parts", "get_env_exact", "merge",
    "apply_two_site_gate", "left_canonicalise", "overlap", "LinAlgError",
    "power_method", "put_env_on_left_site", "put_env_on_right_site",
    "get_env_off_left_site", "get_env_off_right_site",
]

LinAlgError = np.linalg.LinAlgError


# --------------------------------------------------------------------------
# small helpers (qmps/tools.py:43-73)
# --------------------------------------------------------------------------
def cT(tensor):
    """Hermitian conjugate over the last two axes (qmps/tools.py:61-66)."""
    return np.conj(np.swapaxes(tensor, -1, -2))


def direct_sum(A, B):
    """Block-diagonal sum of two matrices (qmps/tools.py:69-73)."""
    out = np.zeros((A.shape[0] + B.shape[0], A.shape[1] + B.shape[1]),
                   dtype=np.result_type(A, B))
    out[:A.shape[0], :A.shape[1]] = A
    out[A.shape[0]:, A.shape[1]:] = B
    return out


def from_real_vector(v):
    """First half real parts, second half imaginary parts (qmps/tools.py:43-46)."""
    v = np.asarray(v)
    h = v.shape[0] // 2
    return v[:h] + 1j * v[h:]


def to_real_vector(A):
    """Inverse of :func:`from_real_vector` on the flattened array (qmps/tools.py:49-52)."""
    A = np.asarray(A)
    return np.concatenate([A.real.reshape(-1), A.imag.reshape(-1)])


# --------------------------------------------------------------------------
# a1 / a2 / a3: embeddings (qmps/tools.py:76-154)
# --------------------------------------------------------------------------
def unitary_to_tensor(U):
    """A[s,i,j] = U[(i,s),(0,j)]  (qmps/tools.py:151-154).

    Row index of U is (left bond i, physical s) big-endian, column index is
    (input qubit q, right bond j); the first input qubit is projected on |0>.
    """
    U = np.asarray(U)
    D = U.shape[0] // 2
    return np.ascontiguousarray(U.reshape(D, 2, 2, D)[:, :, 0, :].transpose(1, 0, 2))


def unitary_extension(Q, D=None):
    """Complete an isometry to a unitary with scipy's SVD null space
    (qmps/tools.py:76-94).  Tall Q: append null_space(Q^dagger) columns; wide Q:
    the mirrored construction; optional identity padding to dimension D."""
    rows, cols = Q.shape
    if rows > cols:
        full = np.concatenate([Q, sla.null_space(Q.conj().T)], axis=1)
    elif rows < cols:
        full = np.concatenate([Q.conj().T, sla.null_space(Q)], axis=1).conj().T
    else:
        full = Q
    if D is not None and D > full.shape[0]:
        full = direct_sum(full, np.eye(D - full.shape[0]))
    return full


def tensor_to_unitary(A, testing=False):
    """iso[(i,s),j] = A[s,i,j]; U = [iso | completion] (qmps/tools.py:123-148).

    Only the first D columns are unique.  ``testing`` mirrors the reference's
    D=2-only internal checks (qmps/tools.py:131-136)."""
    d, D, _ = A.shape
    iso = A.transpose(1, 0, 2).reshape(D * d, D)
    U = unitary_extension(iso)
    if testing:
        ok = (np.allclose(cT(iso) @ iso, np.eye(2))
              and np.allclose(U @ cT(U), np.eye(4))
              and np.allclose(cT(U) @ U, np.eye(4))
              and np.allclose(U[:iso.shape[0], :iso.shape[1]], iso)
              and np.allclose(U.reshape(2, 2, 2, 2)[:, :, 0, :].reshape(4, 2), iso))
        return U, ok
    return U


def environment_to_unitary(v):
    """vec(v)/|v| as column 0 of a unitary, null-space completion
    (qmps/tools.py:97-108)."""
    row = np.asarray(v).reshape(1, -1) / sla.norm(v)
    rest = sla.null_space(row).conj().T
    return np.concatenate([row, rest], axis=0).T


def environment_from_unitary(u):
    """(u e0).reshape(2,2): hard-coded D=2 in the reference (qmps/tools.py:111-120);
    the oracle generalises to any square environment."""
    n = u.shape[0]
    D = int(round(np.sqrt(n)))
    return u[:, 0].reshape(D, D)


# --------------------------------------------------------------------------
# a4 / a6: transfer matrices and their leading eigen-data (xmps)
# --------------------------------------------------------------------------
def transfer_matrix(A, B=None):
    """E[(i,k),(j,l)] = sum_s A[s,i,j] conj(B[s,k,l])
    (in-repo definition new_tdvp/EnvironmentParamSensitivity.py:37-38)."""
    B = A if B is None else B
    D1, D2 = A.shape[1], B.shape[1]
    return np.einsum("sij,skl->ikjl", A, B.conj()).reshape(D1 * D2, A.shape[2] * B.shape[2])


def leading_eig(M):
    """Dense ``eig``; eigenpair of largest modulus (ties: first in LAPACK order)."""
    w, v = np.linalg.eig(M)
    k = int(np.argmax(np.abs(w)))
    return w[k], v[:, k]


def _hermitian_gauge(x):
    """Rotate the phase of a (numerically) Hermitian-up-to-phase matrix so that
    its trace is real positive, then symmetrise (xmps.tensor.rotate_to_hermitian,
    call site qmps/time_evolve_tools.py:107)."""
    t = np.trace(x)
    if abs(t) > 0:
        x = x * (np.conj(t) / abs(t))
    return (x + x.conj().T) / 2


def eigs(A):
    """``TransferMatrix(A).eigs() -> (eta, l, r)`` as the call sites force it
    (qmps/tools.py:181-182, ground_state.py:296-297): r Hermitian
    with tr r = 1, l Hermitian with tr(l r) = 1.  [xmps: parity unpinned]"""
    D = A.shape[1]
    E = transfer_matrix(A)
    eta, vr = leading_eig(E)
    r = _hermitian_gauge(vr.reshape(D, D))
    r = r / np.trace(r).real
    # left action l -> sum_s A_s^dagger l A_s is the Hilbert-Schmidt adjoint: matrix E^dagger
    _, vl = leading_eig(E.conj().T)
    l = _hermitian_gauge(vl.reshape(D, D))
    l = l / np.trace(l @ r).real
    if abs(eta.imag) < 1e-12 * max(1.0, abs(eta)):
        eta = eta.real
    return eta, l, r


def _fix_phase_unit(x):
    """Unit Frobenius norm, phase chosen so that tr x is real non-negative (or,
    for a traceless x, so that its largest entry is real positive).  The phase of
    a mixed fixed point is a gauge freedom; this is the build's
    documented convention, shared with the CUDA kernels."""
    x = x / np.linalg.norm(x)
    t = np.trace(x)
    if abs(t) > 1e-8:
        return x * (np.conj(t) / abs(t))
    k = np.argmax(np.abs(x))
    z = x.reshape(-1)[k]
    return x * (np.conj(z) / abs(z))


def _fix_phase_zgeev(x):
    """Unit Frobenius norm, the entry of largest modulus real positive: LAPACK zgeev's
    eigenvector normalisation, which is what the one xmps output recorded in the reference shows
    (``Time Evo.ipynb`` cells 22-24 print ``Map(A,B).left_fixed_point()[1]`` with unit norm and its
    largest entry 0.76069374+0j; tests/golden/ref_notebook_outputs.json)."""
    x = x / np.linalg.norm(x)
    z = x.reshape(-1)[np.argmax(np.abs(x))]
    return x * (np.conj(z) / abs(z))


_GAUGES = {"zgeev": _fix_phase_zgeev, "trace": _fix_phase_unit}


def right_fixed_point(A, B, gauge="zgeev"):
    """``Map(A,B).right_fixed_point() -> (x, r)``: leading eigenpair of E_AB,
    r with unit Frobenius norm (call sites time_evolve_tools.py:87,
    loschmidts/time_evo.py:81).  [xmps: pinned only by the gauge and norm of the
    notebook-recorded output]"""
    E = transfer_matrix(A, B)
    x, v = leading_eig(E)
    return x, _GAUGES[gauge](v.reshape(A.shape[1], B.shape[1]))


def left_fixed_point(A, B, gauge="zgeev"):
    """``Map(A,B).left_fixed_point() -> (x, l)`` with the left action
    l -> sum_s A_s^dagger l B_s, i.e. the leading eigenpair of
    E_AB^dagger; x is the complex conjugate of the right eigenvalue."""
    E = transfer_matrix(A, B)
    x, v = leading_eig(E.conj().T)
    return x, _GAUGES[gauge](v.reshape(A.shape[1], B.shape[1]))


def env_exact_parts(A):
    """(eta, r, C, v0): trace-1 right fixed point of E_AA, its lower Cholesky
    factor C (r = C C^dagger, positive diagonal; ``cholesky(r).conj().T`` at
    qmps/tools.py:182) and the unique column vec(C)/|C|_F of the environment
    unitary.  Raises ``LinAlgError`` exactly where the reference does."""
    eta, _, r = eigs(A)
    C = sla.cholesky(r).conj().T
    return eta, r, C, C.reshape(-1) / np.linalg.norm(C)


def get_env_exact(U):
    """qmps/tools.py:176-182."""
    _, _, r = eigs(unitary_to_tensor(U))
    return environment_to_unitary(sla.cholesky(r).conj().T)


def merge(A, B):
    """Two-site blocking M[(s1,s2),i,j] = (A^s1 B^s2)[i,j]
    (qmps/time_evolve_tools.py:20-23; the reference hard-codes bond 2, the
    oracle keeps the general (d^2, D, D) shape -- SURVEY A.7)."""
    d1, Dl, _ = A.shape
    d2, _, Dr = B.shape
    return np.einsum("sik,tkj->stij", A, B).reshape(d1 * d2, Dl, Dr)


def apply_two_site_gate(W, M):
    """``tensordot(WW, merge(A,A), [1,0])`` (qmps/loschmidts/time_evo.py:79)."""
    return np.tensordot(W, M, [1, 0])


def left_canonicalise(A):
    """``iMPS([A]).left_canonicalise()[0]`` (call sites time_evolve_tools.py:85-86):
    gauge-transform + rescale a single-site uniform tensor so that
    sum_s A_s^dagger A_s = 1.  With l the left fixed point (l = L^dagger L),
    A_s -> L A_s L^{-1} / sqrt(eta).  [xmps: parity unpinned; gauge not unique]"""
    D = A.shape[1]
    E = transfer_matrix(A)
    eta, v = leading_eig(E.conj().T)
    l = _hermitian_gauge(v.reshape(D, D))
    l = l / np.trace(l).real * D
    L = sla.cholesky(l)                       # upper: l = L^dagger L
    Li = np.linalg.inv(L)
    return np.einsum("ab,sbc,cd->sad", L, A, Li) / np.sqrt(abs(eta))


def overlap(A, B):
    """Per-site fidelity |eta(E_AB)|^2 (qmps/time_evolve_tools.py:84-91;
    ``iMPS.overlap`` is defined here to match ``get_overlap_exact``)."""
    x, _ = right_fixed_point(A, B)
    return abs(x) ** 2


def power_method(A, B, K, r0=None):
    """K normalised applications r <- sum_s A_s r B_s^dagger / |.|_F starting from
    I/sqrt(D) (``qmps.ipynb`` cells 29-32; SURVEY 8(d) cfg 5).  Returns
    (r_K, rayleigh) with rayleigh = <r_K, E r_K> for the unit-norm r_K."""
    D = A.shape[1]
    r = np.eye(D, dtype=np.complex128) / np.sqrt(D) if r0 is None else r0.astype(np.complex128)
    Bh = np.conj(np.swapaxes(B, -1, -2))
    apply = lambda x: np.sum(A @ x @ Bh, axis=0)      # sum_s A_s x B_s^dagger
    for _ in range(K):
        r = apply(r)
        r = r / np.linalg.norm(r)
    Er = apply(r)
    return r, np.vdot(r, Er)


# --------------------------------------------------------------------------
# 2x2 environments embedded on a site (qmps/time_evolve_tools.py:38-74)
# --------------------------------------------------------------------------
_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


def put_env_on_left_site(q, ret_n=False):
    """qmps/time_evolve_tools.py:38-53: two explicit orthonormal rows built from
    q^T/|q|, null-space completion, then a SWAP from the left."""
    a, b, c, d = np.asarray(q).T.reshape(-1)
    n = np.sqrt(abs(a) ** 2 + abs(b) ** 2 + abs(c) ** 2 + abs(d) ** 2)
    rows = np.array([[a, np.conj(c), b, np.conj(d)],
                     [c, -np.conj(a), d, -np.conj(b)]]) / n
    full = np.concatenate([rows, sla.null_space(rows).conj().T], axis=0)
    full = _SWAP @ full
    return (full, n) if ret_n else full


def get_env_off_left_site(A):
    """qmps/time_evolve_tools.py:55-57."""
    return A.reshape(2, 2, 2, 2)[:, 0, :, 0].T


def put_env_on_right_site(q, ret_n=False):
    """qmps/time_evolve_tools.py:59-70."""
    a, b, c, d = np.asarray(q).reshape(-1)
    n = np.sqrt(abs(a) ** 2 + abs(b) ** 2 + abs(c) ** 2 + abs(d) ** 2)
    rows = np.array([[a, b, np.conj(d), -np.conj(c)],
                     [c, d, -np.conj(b), np.conj(a)]]) / n
    full = np.concatenate([rows, sla.null_space(rows).conj().T], axis=0)
    return (full, n) if ret_n else full


def get_env_off_right_site(A):
    """qmps/time_evolve_tools.py:72-74."""
    return A.reshape(2, 2, 2, 2)[0, :, 0, :]
