// qmps_b200 leading eigenvalue of a D = 2 mixed transfer matrix, ONE THREAD per problem, everything
// in registers (Map(A,B).right_fixed_point()/.left_fixed_point() eigenvalue at D = 2: call sites
// qmps/time_evolve_tools.py:87, qmps/loschmidts/time_evo.py:79-82 -- the Loschmidt / TDVP-step cost
// of every D = 2 script in the reference).
//
// E[(i,k),(j,l)] = sum_s A[s,i,j] conj(B[s,k,l]) is 4 x 4 complex.  Same algorithm as the generic
// path (core.cuh: Householder reduction to Hessenberg form + complex single-shift QR for ALL
// eigenvalues, arg-max modulus), but written for a compile-time size so that every register index
// is static: the deflation stage EN is a template parameter, the window start l a predicate.
#pragma once
#include "core.cuh"

namespace qmps {

// E (+)= A_s (x) conj(B_s) for one physical index; a, b: 2 x 2 row-major
template <typename T> QMPS_HD void fpd2_accumulate(cx<T> (&E)[4][4], const cx<T>* a, const cx<T>* b, int adjoint) {
#pragma unroll
  for (int row = 0; row < 4; ++row)
#pragma unroll
    for (int col = 0; col < 4; ++col) {
      const int i = row >> 1, k = row & 1, j = col >> 1, l = col & 1;
      // adjoint: E^dagger[(row),(col)] = conj(E[(col),(row)]) = conj(A[j,i]) B[l,k]
      if (!adjoint) cmad_c(E[row][col], a[i * 2 + j], b[k * 2 + l]);
      else cmad_c(E[row][col], b[l * 2 + k], a[j * 2 + i]);
    }
}

// Householder reduction of a 4 x 4 matrix to upper Hessenberg form (core.cuh::hessenberg, unrolled)
template <typename T> QMPS_HD void fpd2_hessenberg(cx<T> (&H)[4][4]) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const cx<T> alpha = H[k + 1][k];
    T xn2 = T(0);
#pragma unroll
    for (int i = k + 2; i < 4; ++i) xn2 += norm2(H[i][k]);
    if (xn2 == T(0) && alpha.im == T(0)) continue;
    T beta = sqrt(norm2(alpha) + xn2);
    if (alpha.re > T(0)) beta = -beta;
    const cx<T> tau = mk<T>((beta - alpha.re) / beta, -alpha.im / beta);
    const cx<T> scal = cinv(alpha - mk<T>(beta, 0));
    cx<T> v[4];
#pragma unroll
    for (int i = k + 1; i < 4; ++i) {
      if (i == k + 1) { v[i] = mk<T>(1, 0); H[i][k] = mk<T>(beta, 0); }
      else { v[i] = H[i][k] * scal; H[i][k] = mk<T>(0, 0); }
    }
    const cx<T> ctau = conj(tau);
#pragma unroll
    for (int j = k + 1; j < 4; ++j) {                  // left:  H[k+1:, k+1:] -= conj(tau) v (v^H H)
      cx<T> s = mk<T>(0, 0);
#pragma unroll
      for (int i = k + 1; i < 4; ++i) cmad(s, conj(v[i]), H[i][j]);
      s = s * ctau;
#pragma unroll
      for (int i = k + 1; i < 4; ++i) cmsub(H[i][j], v[i], s);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {                      // right: H[:, k+1:] -= tau (H v) v^H
      cx<T> s = mk<T>(0, 0);
#pragma unroll
      for (int j = k + 1; j < 4; ++j) cmad(s, H[i][j], v[j]);
      s = s * tau;
#pragma unroll
      for (int j = k + 1; j < 4; ++j) cmsub(H[i][j], s, conj(v[j]));
    }
  }
}

// one explicit shifted QR sweep on the active window l .. EN (core.cuh::hqr_eigenvalues, unrolled)
template <typename T, int EN> QMPS_HD void fpd2_sweep(cx<T> (&H)[4][4], int l, cx<T> sh) {
  cx<T> rc[4], rs[4];
#pragma unroll
  for (int i = 0; i <= EN; ++i) if (i >= l) H[i][i] = H[i][i] - sh;
#pragma unroll
  for (int i = 1; i <= EN; ++i) {
    rc[i] = mk<T>(1, 0); rs[i] = mk<T>(0, 0);
    if (i > l) {
      const cx<T> f = H[i - 1][i - 1], gg = H[i][i - 1];
      const T nr2 = norm2(f) + norm2(gg);
      T nr = T(0);
      cx<T> c = mk<T>(1, 0), s = mk<T>(0, 0);
      if (nr2 > nb_floor<T>::v()) { const T inr = rsq_nb(nr2); nr = nr2 * inr; c = f * inr; s = gg * inr; }   // rsq_nb: no slow-path call
      rc[i] = c; rs[i] = s;
#pragma unroll
      for (int j = i; j <= EN; ++j) {
        const cx<T> p = H[i - 1][j], q = H[i][j];
        H[i - 1][j] = conj(c) * p + conj(s) * q;
        H[i][j] = c * q - s * p;
      }
      H[i - 1][i - 1] = mk<T>(nr, 0);
      H[i][i - 1] = mk<T>(0, 0);
    }
  }
#pragma unroll
  for (int i = 0; i <= EN; ++i) {
    if (i < l) continue;
#pragma unroll
    for (int j = 1; j <= EN; ++j) {
      if (j < i || j <= l) continue;
      const cx<T> c = rc[j], s = rs[j];
      const cx<T> xx = H[i][j - 1], yy = H[i][j];
      H[i][j - 1] = xx * c + yy * s;
      H[i][j] = yy * conj(c) - xx * conj(s);
    }
  }
#pragma unroll
  for (int i = 0; i <= EN; ++i) if (i >= l) H[i][i] = H[i][i] + sh;
}

// deflate eigenvalue EN of the Hessenberg matrix; returns 1 if the sweep limit was hit
template <typename T, int EN> QMPS_HD int fpd2_stage(cx<T> (&H)[4][4], cx<T> (&w)[4]) {
  const T eps = eps_of<T>::v();
  const int maxit = 60;
  for (int its = 0;; ++its) {
    int l = 0;
    bool found = false;
#pragma unroll
    for (int t = EN; t >= 1; --t) {
      if (!found) {
        T s = cabs1(H[t - 1][t - 1]) + cabs1(H[t][t]);
        if (s == T(0)) s = T(1);
        if (cabs1(H[t][t - 1]) <= eps * s) { l = t; found = true; }
      }
    }
    if (l == EN || its >= maxit) {
      w[EN] = H[EN][EN];
      H[EN][EN - 1] = mk<T>(0, 0);
      return l != EN;
    }
    cx<T> sh;
    if (its == 10 || its == 20 || its == 30 || its == 40) {
      T t = fabs(H[EN][EN - 1].re);
      if (EN >= 2) t += fabs(H[EN - 1][EN >= 2 ? EN - 2 : 0].re);
      sh = H[EN][EN] + mk<T>(t, 0);
    } else {
      const cx<T> a = H[EN - 1][EN - 1], b = H[EN - 1][EN], c = H[EN][EN - 1], d = H[EN][EN];
      sh = d;
      const cx<T> bc = b * c;
      if (bc.re != T(0) || bc.im != T(0)) {
        const cx<T> y = (a - d) * T(0.5);
        cx<T> z = csqrt_nb(y * y + bc);
        if (y.re * z.re + y.im * z.im < T(0)) z = -z;
        sh = d - cdiv_nb(bc, y + z);
      }
    }
#pragma unroll
    for (int t = 1; t <= EN; ++t) if (t == l) H[t][t - 1] = mk<T>(0, 0);
    fpd2_sweep<T, EN>(H, l, sh);
  }
}

// all eigenvalues of a 4 x 4 complex matrix (destroyed); returns ST_OK / ST_NO_CONVERGE
template <typename T> QMPS_HD int fpd2_eigenvalues(cx<T> (&H)[4][4], cx<T> (&w)[4]) {
  fpd2_hessenberg<T>(H);
  int fail = fpd2_stage<T, 3>(H, w);
  fail |= fpd2_stage<T, 2>(H, w);
  fail |= fpd2_stage<T, 1>(H, w);
  w[0] = H[0][0];
  return fail ? ST_NO_CONVERGE : ST_OK;
}

// One inverse iteration on (M - lambda) for the eigenvector, in registers: Gaussian elimination with
// partial pivoting on the 4 x 5 augmented system (physical row swaps by predicated moves; same pivot
// choice, tiny-pivot replacement and right-hand side as core.cuh::lu_solve_aug / generic.cuh::
// leading_eigenpair).  M: the map (destroyed).  x: un-normalised solution.
template <typename T> QMPS_HD void fpd2_inverse_iteration(cx<T> (&M)[4][4], cx<T> lam, cx<T> (&x)[4]) {
  cx<T> rhs[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    M[i][i] = M[i][i] - lam;
    T t = T(0.61803398874989485) * T(i + 1);
    t -= floor(t);
    rhs[i] = mk<T>(T(0.5) + t, T(0.25) - T(0.5) * t);
  }
  T scale = cabs(lam);
  if (!(scale > T(1e-30))) scale = T(1);
  const T tiny = eps_of<T>::v() * scale;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int p = k;
    T best = norm2(M[k][k]);
#pragma unroll
    for (int i = k + 1; i < 4; ++i) { const T a = norm2(M[i][k]); if (a > best) { best = a; p = i; } }
#pragma unroll
    for (int i = k + 1; i < 4; ++i) {
      if (p == i) {
#pragma unroll
        for (int j = k; j < 4; ++j) { const cx<T> t = M[k][j]; M[k][j] = M[i][j]; M[i][j] = t; }
        const cx<T> t = rhs[k]; rhs[k] = rhs[i]; rhs[i] = t;
      }
    }
    cx<T> pv = M[k][k];
    if (!(best >= tiny * tiny)) pv = mk<T>(tiny, 0);
    M[k][k] = pv;
    const cx<T> inv = cinv(pv);
#pragma unroll
    for (int i = k + 1; i < 4; ++i) {
      const cx<T> f = M[i][k] * inv;
#pragma unroll
      for (int j = k + 1; j < 4; ++j) cmsub(M[i][j], f, M[k][j]);
      cmsub(rhs[i], f, rhs[k]);
    }
  }
#pragma unroll
  for (int k = 3; k >= 0; --k) {
    cx<T> acc = rhs[k];
#pragma unroll
    for (int j = k + 1; j < 4; ++j) cmsub(acc, M[k][j], x[j]);
    x[k] = cdiv(acc, M[k][k]);
  }
}

// unit 2-norm and a phase convention: gauge 0 = generic.cuh::leading_eigenpair (trace of the 2 x 2
// reshape real non-negative; traceless: largest entry real positive), gauge 1 = zgeev /
// scipy.linalg.eig (component of largest modulus real positive)
template <typename T> QMPS_HD void fpd2_fix_gauge(cx<T> (&x)[4], int gauge) {
  T nrm2 = T(0), bigv = T(-1);
  cx<T> big = mk<T>(1, 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const T a = norm2(x[i]);
    nrm2 += a;
    if (a > bigv) { bigv = a; big = x[i]; }
  }
  const T inv = T(1) / sqrt(nrm2);
  const cx<T> tr = x[0] + x[3];
  cx<T> ph;
  if (gauge == 0 && cabs(tr) * inv > T(1e-8)) ph = conj(tr) * (T(1) / cabs(tr));
  else ph = conj(big) * (T(1) / sqrt(bigv));
  ph = ph * inv;
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = x[i] * ph;
}

// leading eigenvalue (largest modulus, first on ties -- core.cuh::argmax_abs) of a 4 x 4 map (destroyed)
template <typename T> QMPS_HD int fpd2_leading_of(cx<T> (&E)[4][4], cx<T>* lambda_out) {
  cx<T> w[4];
  const int status = fpd2_eigenvalues<T>(E, w);
  int k = 0;
  T best = norm2(w[0]);
#pragma unroll
  for (int i = 1; i < 4; ++i) { const T a2 = norm2(w[i]); if (a2 > best) { best = a2; k = i; } }
  cx<T> lam = w[0];
#pragma unroll
  for (int i = 1; i < 4; ++i) if (i == k) lam = w[i];
  *lambda_out = lam;
  return status;
}

// the same from tensors A, B [d][2][2] (any address space); adjoint = 1: the left fixed point's map
// E^dagger.  vec (optional, 4 entries): the eigenvector, unit norm, gauge 0.
template <typename T>
QMPS_HD void fpd2_build(const cx<T>* A, const cx<T>* B, int d, int adjoint, cx<T> (&E)[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) E[r][c] = mk<T>(0, 0);
  for (int s = 0; s < d; ++s) {
    cx<T> a[4], b[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { a[q] = A[s * 4 + q]; b[q] = B[s * 4 + q]; }
    fpd2_accumulate<T>(E, a, b, adjoint);
  }
}

template <typename T>
QMPS_HD int fpd2_leading(const cx<T>* A, const cx<T>* B, int d, int adjoint, cx<T>* lambda_out, cx<T>* vec = nullptr) {
  cx<T> E[4][4];
  fpd2_build<T>(A, B, d, adjoint, E);
  const int status = fpd2_leading_of<T>(E, lambda_out);
  if (vec) {
    fpd2_build<T>(A, B, d, adjoint, E);          // the QR iteration destroyed the map: rebuild (operands are cached)
    cx<T> x[4];
    fpd2_inverse_iteration<T>(E, *lambda_out, x);
    fpd2_fix_gauge<T>(x, 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) vec[i] = x[i];
  }
  return status;
}

}  // namespace qmps
