// qmps_b200 D = 2 fast path: one thread per problem, everything in registers.
//
// For a left-canonical A (every tensor that comes from a unitary,
// qmps/tools.py:151-154) the leading eigenvalue of E_AA is 1, so the environment
// (TransferMatrix(A).eigs() at qmps/tools.py:181) is the fixed point of the
// trace-preserving channel r -> sum_s A_s r A_s^dagger.  In the Pauli basis
// r = (1 + x.sigma)/2 that is the REAL 3x3 affine system (1 - T) x = c:
// ~64 products to build T and c, a 3x3 solve, no 4x4 complex elimination.
#pragma once
#include "core.cuh"
#include "ansatz.cuh"

namespace qmps {

// a[s*4 + i*2 + j] = A[s][i][j].  Outputs: r[i*2+j] (Hermitian, trace 1), eta,
// optionally the lower Cholesky factor C (r = C C^dagger).  Returns a status code.
template <typename T, bool WANT_C>
QMPS_HD int env_d2_solve(const cx<T>* a, cx<T>* r, T* eta, cx<T>* C) {
  T pp = 0, qq = 0, uu = 0, vv = 0;
  cx<T> al = mk<T>(0, 0), be = al, ga = al, de = al, ep = al, ze = al;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const cx<T> p = a[s * 4 + 0], q = a[s * 4 + 1], u = a[s * 4 + 2], v = a[s * 4 + 3];
    pp += norm2(p); qq += norm2(q); uu += norm2(u); vv += norm2(v);
    cmad_c(al, p, q); cmad_c(be, u, v); cmad_c(ga, q, u);
    cmad_c(de, p, v); cmad_c(ep, p, u); cmad_c(ze, q, v);
  }
  // c_a = tr(sigma_a Phi(1))/2 ; T_ab = tr(sigma_a Phi(sigma_b))/2   (a,b in x,y,z)
  const T cxv = ep.re + ze.re, cyv = -(ep.im + ze.im), czv = T(0.5) * (pp + qq - uu - vv);
  // m = 1 - T
  const T m00 = T(1) - (ga.re + de.re), m01 = (ga.im - de.im), m02 = -(ep.re - ze.re);
  const T m10 = (ga.im + de.im), m11 = T(1) + (ga.re - de.re), m12 = (ep.im - ze.im);
  const T m20 = -(al.re - be.re), m21 = -(al.im - be.im), m22 = T(1) - T(0.5) * (pp - qq - uu + vv);
  // adjugate / determinant
  const T a00 = m11 * m22 - m12 * m21, a01 = m02 * m21 - m01 * m22, a02 = m01 * m12 - m02 * m11;
  const T a10 = m12 * m20 - m10 * m22, a11 = m00 * m22 - m02 * m20, a12 = m02 * m10 - m00 * m12;
  const T a20 = m10 * m21 - m11 * m20, a21 = m01 * m20 - m00 * m21, a22 = m00 * m11 - m01 * m10;
  const T det = m00 * a00 + m01 * a10 + m02 * a20;
  int status = ST_OK;
  if (!(fabs(det) > T(64) * eps_of<T>::v())) status = ST_SINGULAR;
  const T idet = T(1) / det;
  T x0 = (a00 * cxv + a01 * cyv + a02 * czv) * idet;
  T x1 = (a10 * cxv + a11 * cyv + a12 * czv) * idet;
  T x2 = (a20 * cxv + a21 * cyv + a22 * czv) * idet;
  {  // one step of iterative refinement: cheap, removes the adjugate's extra rounding
    const T r0 = cxv - (m00 * x0 + m01 * x1 + m02 * x2);
    const T r1 = cyv - (m10 * x0 + m11 * x1 + m12 * x2);
    const T r2 = czv - (m20 * x0 + m21 * x1 + m22 * x2);
    x0 += (a00 * r0 + a01 * r1 + a02 * r2) * idet;
    x1 += (a10 * r0 + a11 * r1 + a12 * r2) * idet;
    x2 += (a20 * r0 + a21 * r1 + a22 * r2) * idet;
  }
  const T r00 = T(0.5) * (T(1) + x2), r11 = T(0.5) * (T(1) - x2);
  const T r01re = T(0.5) * x0, r01im = -T(0.5) * x1;
  r[0] = mk<T>(r00, 0); r[1] = mk<T>(r01re, r01im);
  r[2] = mk<T>(r01re, -r01im); r[3] = mk<T>(r11, 0);
  // eta = tr(Phi(r)) / tr(r) = tr(r sum_s A_s^dagger A_s): 1 for an exact isometry
  *eta = r00 * (pp + uu) + r11 * (qq + vv) + T(2) * (r01re * (al.re + be.re) - r01im * (al.im + be.im));
  if (WANT_C) {
    T c00 = T(1), c11 = T(1);
    if (!(r00 > T(0))) { if (status == ST_OK) status = ST_NOT_PD; } else c00 = sqrt(r00);
    const T ic = T(1) / c00;
    const T c10re = r01re * ic, c10im = -r01im * ic;          // C10 = r10 / C00
    const T d11 = r11 - (c10re * c10re + c10im * c10im);
    if (!(d11 > T(0))) { if (status == ST_OK) status = ST_NOT_PD; } else c11 = sqrt(d11);
    C[0] = mk<T>(c00, 0); C[1] = mk<T>(0, 0); C[2] = mk<T>(c10re, c10im); C[3] = mk<T>(c11, 0);
  } else {
    if (status == ST_OK && !(x0 * x0 + x1 * x1 + x2 * x2 < T(1))) status = ST_NOT_PD;
  }
  return status;
}

// 2x2 complex product c = a b
template <typename T> QMPS_HD void mm2(const cx<T>* a, const cx<T>* b, cx<T>* c) {
  c[0] = a[0] * b[0]; cmad(c[0], a[1], b[2]);
  c[1] = a[0] * b[1]; cmad(c[1], a[1], b[3]);
  c[2] = a[2] * b[0]; cmad(c[2], a[3], b[2]);
  c[3] = a[2] * b[1]; cmad(c[3], a[3], b[3]);
}

// e = sum_ab H_ab tr(M_a^dagger M_b r), M_(s1 s2) = A_s1 A_s2   (the
// reference evaluates the same number by simulating State(U,V,2) on 4 qubits,
// qmps/ground_state.py:150-168).  hmat: 16 entries H[a*4+b], any address space.
template <typename T>
QMPS_HD T energy_d2(const cx<T>* a, const cx<T>* r, const cx<T>* hmat) {
  cx<T> M[16], P[4];
#pragma unroll
  for (int s1 = 0; s1 < 2; ++s1)
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) mm2(a + 4 * s1, a + 4 * s2, M + 4 * (2 * s1 + s2));
  T e = T(0);
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    mm2(M + 4 * b, r, P);
#pragma unroll
    for (int ij = 0; ij < 4; ++ij) {
      cx<T> q = mk<T>(0, 0);                      // Q_b = sum_a H[b][a] M_a  (H Hermitian)
#pragma unroll
      for (int aa = 0; aa < 4; ++aa) cmad(q, hmat[b * 4 + aa], M[4 * aa + ij]);
      e += q.re * P[ij].re + q.im * P[ij].im;    // Re(conj(Q) P)
    }
  }
  return e;
}

}  // namespace qmps
