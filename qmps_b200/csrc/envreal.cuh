// qmps_b200 exact environment of a LEFT-CANONICAL tensor as a REAL linear system.
//
// r is the fixed point of the Hermiticity-preserving map Phi(r) = sum_s A_s r A_s^dagger
// (TransferMatrix(A).eigs() at qmps/tools.py:181 with eta = 1 known), so it has only
// n = D^2 REAL degrees of freedom.  Unknowns (index u):
//     u <  D             d_u  = r[u][u]
//     u = D + 2p + 0     x_jl = Re r[j][l]        pair p = (j < l), pairs ordered (0,1),(0,2),...
//     u = D + 2p + 1     y_jl = Im r[j][l]
// Equations (index e, same enumeration): e < D: Re (Phi(r) - r)[e][e] = 0;
// e = D + 2p + part: Re / Im of (Phi(r) - r)[i][k] = 0 for the pair p = (i < k).
// With G(j,l) = E[(i,k),(j,l)] = sum_s A[s,i,j] conj(A[s,k,l]):
//     coefficient of d_j  : G(j,j)
//     coefficient of x_jl : G(j,l) + G(l,j)
//     coefficient of y_jl : i (G(j,l) - G(l,j))
// minus the identity.  Equation 0 is redundant (the diagonal rows sum to zero because Phi
// preserves the trace) and is replaced by  sum_j d_j = 1.
// A real n x n solve costs (2/3) n^3 flops instead of (8/3) n^3 for the complex system --
// and a thread can hold its whole row (n + 1 numbers) in registers.
#pragma once
#include "core.cuh"

namespace qmps {

// equation / unknown index -> (i, k, part); part 0 = real, 1 = imaginary
template <int D> QMPS_HD void herm_index(int e, int* i, int* k, int* part) {
  if (e < D) { *i = e; *k = e; *part = 0; return; }
  int p = (e - D) >> 1;
  *part = (e - D) & 1;
  int a = 0;
  while (p >= D - 1 - a) { p -= D - 1 - a; ++a; }
  *i = a; *k = a + 1 + p;
}

// Row e of the (n x (n+1)) augmented real system into m[0..n] (m[n] = right-hand side).
// Ap: tensor with padded rows, element A[s][i][j] at Ap[(s*D + i)*lda + j].
template <typename T, int D>
QMPS_HD void herm_row(const cx<T>* Ap, int lda, int d, int e, T* m) {
  constexpr int n = D * D;
  int i, k, part;
  herm_index<D>(e, &i, &k, &part);
  const cx<T>* Ai = Ap + i * lda;
  const cx<T>* Ak = Ap + k * lda;
  const int sstride = D * lda;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    cx<T> g = mk<T>(0, 0);
    for (int s = 0; s < d; ++s) cmad_c(g, Ai[s * sstride + j], Ak[s * sstride + j]);
    m[j] = (part ? g.im : g.re) - (e == j ? T(1) : T(0));
  }
  int u = D;
#pragma unroll
  for (int j = 0; j < D; ++j) {
#pragma unroll
    for (int l = j + 1; l < D; ++l) {
      cx<T> g1 = mk<T>(0, 0), g2 = mk<T>(0, 0);
      for (int s = 0; s < d; ++s) {
        const cx<T> aij = Ai[s * sstride + j], ail = Ai[s * sstride + l];
        const cx<T> akj = Ak[s * sstride + j], akl = Ak[s * sstride + l];
        cmad_c(g1, aij, akl);          // G(j,l)
        cmad_c(g2, ail, akj);          // G(l,j)
      }
      const T cx_re = g1.re + g2.re, cx_im = g1.im + g2.im;          // coefficient of x_jl
      const T cy_re = -(g1.im - g2.im), cy_im = g1.re - g2.re;       // i (G(j,l) - G(l,j))
      m[u] = (part ? cx_im : cx_re) - (e == u ? T(1) : T(0));
      m[u + 1] = (part ? cy_im : cy_re) - (e == u + 1 ? T(1) : T(0));
      u += 2;
    }
  }
  m[n] = T(0);
  if (e == 0) {
#pragma unroll
    for (int uu = 0; uu < n; ++uu) m[uu] = uu < D ? T(1) : T(0);
    m[n] = T(1);
  }
}

// Same row, for a compile-time physical dimension DP: the D*DP entries A[s][i][:] of "my" left
// index stay in registers and only A[s][k][:] is streamed from shared memory -- 88 instead of 240
// 16-byte shared loads per row at D = 8, DP = 2 (the row build was LSU-bound:
// profiles/ncu_er8_src_r01f.txt).  Identical arithmetic order per entry as herm_row.
template <typename T, int D, int DP>
QMPS_HD void herm_row_cached(const cx<T>* Ap, int lda, int e, T* m) {
  constexpr int n = D * D;
  int i, k, part;
  herm_index<D>(e, &i, &k, &part);
  const cx<T>* Ai = Ap + i * lda;
  const cx<T>* Ak = Ap + k * lda;
  const int sstride = D * lda;
  cx<T> ai[DP][D];
#pragma unroll
  for (int s = 0; s < DP; ++s)
#pragma unroll
    for (int j = 0; j < D; ++j) ai[s][j] = Ai[s * sstride + j];
#pragma unroll
  for (int l = 0; l < D; ++l) {
    cx<T> bl[DP];
#pragma unroll
    for (int s = 0; s < DP; ++s) bl[s] = Ak[s * sstride + l];
    {
      cx<T> g = mk<T>(0, 0);
#pragma unroll
      for (int s = 0; s < DP; ++s) cmad_c(g, ai[s][l], bl[s]);
      m[l] = (part ? g.im : g.re) - (e == l ? T(1) : T(0));
    }
#pragma unroll
    for (int j = 0; j < l; ++j) {
      const int u = D + 2 * (j * D - (j * (j + 1)) / 2 + (l - j - 1));
      cx<T> g1 = mk<T>(0, 0), g2 = mk<T>(0, 0);
#pragma unroll
      for (int s = 0; s < DP; ++s) {
        const cx<T> akj = Ak[s * sstride + j];
        cmad_c(g1, ai[s][j], bl[s]);        // G(j,l)
        cmad_c(g2, ai[s][l], akj);          // G(l,j)
      }
      const T cx_re = g1.re + g2.re, cx_im = g1.im + g2.im;
      const T cy_re = -(g1.im - g2.im), cy_im = g1.re - g2.re;
      m[u] = (part ? cx_im : cx_re) - (e == u ? T(1) : T(0));
      m[u + 1] = (part ? cy_im : cy_re) - (e == u + 1 ? T(1) : T(0));
    }
  }
  m[n] = T(0);
  if (e == 0) {
#pragma unroll
    for (int uu = 0; uu < n; ++uu) m[uu] = uu < D ? T(1) : T(0);
    m[n] = T(1);
  }
}

// Middle ground for the register-capped complex128 D = 8 kernel: only the l-side operands
// (A[s][i][l], A[s][k][l]) are hoisted out of the pair loop -- 144 instead of 240 shared loads per
// row for 8 extra live doubles.  Same arithmetic order per entry as herm_row.
template <typename T, int D, int DP>
QMPS_HD void herm_row_lhoist(const cx<T>* Ap, int lda, int e, T* m) {
  constexpr int n = D * D;
  int i, k, part;
  herm_index<D>(e, &i, &k, &part);
  const cx<T>* Ai = Ap + i * lda;
  const cx<T>* Ak = Ap + k * lda;
  const int sstride = D * lda;
#pragma unroll
  for (int l = 0; l < D; ++l) {
    cx<T> ail[DP], akl[DP];
#pragma unroll
    for (int s = 0; s < DP; ++s) { ail[s] = Ai[s * sstride + l]; akl[s] = Ak[s * sstride + l]; }
    {
      cx<T> g = mk<T>(0, 0);
#pragma unroll
      for (int s = 0; s < DP; ++s) cmad_c(g, ail[s], akl[s]);
      m[l] = (part ? g.im : g.re) - (e == l ? T(1) : T(0));
    }
#pragma unroll
    for (int j = 0; j < l; ++j) {
      const int u = D + 2 * (j * D - (j * (j + 1)) / 2 + (l - j - 1));
      cx<T> g1 = mk<T>(0, 0), g2 = mk<T>(0, 0);
#pragma unroll
      for (int s = 0; s < DP; ++s) {
        const cx<T> aij = Ai[s * sstride + j], akj = Ak[s * sstride + j];
        cmad_c(g1, aij, akl[s]);            // G(j,l)
        cmad_c(g2, ail[s], akj);            // G(l,j)
      }
      const T cx_re = g1.re + g2.re, cx_im = g1.im + g2.im;
      const T cy_re = -(g1.im - g2.im), cy_im = g1.re - g2.re;
      m[u] = (part ? cx_im : cx_re) - (e == u ? T(1) : T(0));
      m[u + 1] = (part ? cy_im : cy_re) - (e == u + 1 ? T(1) : T(0));
    }
  }
  m[n] = T(0);
  if (e == 0) {
#pragma unroll
    for (int uu = 0; uu < n; ++uu) m[uu] = uu < D ? T(1) : T(0);
    m[n] = T(1);
  }
}

// solution vector x[n] -> Hermitian r (row-major D x D, unpadded); one entry pair per call
template <typename T, int D> QMPS_HD void herm_scatter(const T* x, int e, cx<T>* r) {
  int i, k, part;
  herm_index<D>(e, &i, &k, &part);
  if (e < D) { r[e * D + e] = mk<T>(x[e], 0); return; }
  if (part == 0) {
    const T re = x[e], im = x[e + 1];
    r[i * D + k] = mk<T>(re, im);
    r[k * D + i] = mk<T>(re, -im);
  }
}

}  // namespace qmps
