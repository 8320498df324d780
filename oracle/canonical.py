"""Oracle, part 4: canonical forms and local expectation values.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  numpy/scipy, complex128.

These are xmps' ``iMPS`` methods as the reference's loops call them.  xmps
(github.com/fergusbarratt/xmps, un-vendored, unpinned) is absent, so this is a
restatement of the published definitions anchored on what the call sites and
the reference's own property tests force -- **parity unpinned**:

* ``iMPS([A]).left_canonicalise()``   call sites qmps/time_evolve_tools.py:85-86,
  qmps/loschmidts/time_evo.py:76,143 -- ``tensors.left_canonicalise``
* ``iMPS([A]).mixed() -> (AL, AR, C)`` qmps/tools.py:184-186, qmps/ground_state.py:287;
  properties asserted by tests/test_represent.py:23-31
* ``iMPS([A]).Es(ops)`` / ``.E(op)``   qmps/loschmidts/time_evo.py:144,
  scripts/loschmidt.py:369; pinned by tests/test_represent.py:33-48 to the Bloch
  vector of the embedded circuit state (re-derived in tests/test_oracle.py with
  the gate-by-gate simulator)
* ``iMPS.overlap(other)``              qmps/loschmidts/time_evo.py:145 --
  ``tensors.overlap``

The gauge (which L with l = L^dagger L, which C with r = C C^dagger) is not
unique; this build fixes the Cholesky gauge.  Everything the call sites consume
(eta, spectra, expectation values, overlaps, costs) is gauge invariant.
"""
import numpy as np
import scipy.linalg as sla

from .tensors import eigs, left_canonicalise, transfer_matrix

__all__ = ["mixed", "expectation_values", "expectation_values_left_canonical", "is_left_canonical",
           "is_right_canonical"]


def is_left_canonical(A, tol=1e-10):
    D = A.shape[1]
    return np.allclose(np.einsum("sij,sik->jk", A.conj(), A), np.eye(D), atol=tol)


def is_right_canonical(A, tol=1e-10):
    D = A.shape[1]
    return np.allclose(np.einsum("sij,skj->ik", A, A.conj()), np.eye(D), atol=tol)


def mixed(A, assume_left_canonical=False):
    """``iMPS([A]).mixed() -> (AL, AR, C)`` with r = C C^dagger the trace-1 right fixed
    point of E_ALAL (C lower triangular, positive diagonal -- the same C as
    ``cholesky(r).conj().T`` of qmps/tools.py:182) and AR = C^-1 AL C, so that
    (tests/test_represent.py:23-31) Map(AL,AL) has right eigenvector r and left 1,
    Map(AR,AR) has right eigenvector 1 and left C^dagger C."""
    AL = A if assume_left_canonical else left_canonicalise(A)
    _, _, r = eigs(AL)
    C = sla.cholesky(r).conj().T
    Ci = np.linalg.inv(C)
    AR = np.einsum("ab,sbc,cd->sad", Ci, AL, C)
    return AL, AR, C


def expectation_values(A, ops):
    """``iMPS([A]).Es(ops)``: <O> = l^T E_O r / (eta l^T r) for ANY normalisable A, with
    E_O[(i,k),(j,l)] = sum_st O[s,t] A[t,i,j] conj(A[s,k,l]) and (l, r) the left/right
    leading eigenvectors of E_AA.  Returns complex values (real for Hermitian ops)."""
    D = A.shape[1]
    E = transfer_matrix(A)
    w, vr = np.linalg.eig(E)
    k = int(np.argmax(np.abs(w)))
    eta, r = w[k], vr[:, k]
    wl, vl = np.linalg.eig(E.T)
    kl = int(np.argmin(np.abs(wl - eta)))
    lv = vl[:, kl]
    out = []
    for O in ops:
        EO = np.einsum("st,tij,skl->ikjl", np.asarray(O, dtype=complex), A, A.conj()).reshape(D * D, D * D)
        out.append((lv @ EO @ r) / (eta * (lv @ r)))
    return np.array(out)


def expectation_values_left_canonical(AL, r, ops):
    """Same for a left-canonical tensor and its trace-1 right fixed point:
    <O> = sum_st O[s,t] tr(A_t r A_s^dagger)."""
    return np.array([np.einsum("st,tij,jl,sil->", np.asarray(O, dtype=complex), AL, r, AL.conj()) for O in ops])
