// qmps_b200 generic-D per-problem solvers (D = 2 .. 16), one GROUP of lanes per
// problem, matrices in shared memory (global workspace for n = D^2 > 64).
#pragma once
#include "core.cuh"
#include "ansatz.cuh"

namespace qmps {

// sum of one value per lane over the group; red[] has g.size entries
template <typename T> QMPS_HDN T group_sum(const Grp& g, T v, T* red) {
  g.sync();
  red[g.lane] = v;
  g.sync();
  T s = T(0);
  for (int i = 0; i < g.size; ++i) s += red[i];
  return s;
}

// x (n = D*D entries, row-major D x D) -> Hermitian part, in place
template <typename T> QMPS_HDN void hermitise(const Grp& g, cx<T>* x, int D) {
  g.sync();
  for (int e = g.lane; e < D * D; e += g.size) {
    int i = e / D, j = e - i * D;
    if (j < i) continue;
    cx<T> a = x[i * D + j], b = x[j * D + i];
    cx<T> h = mk<T>(T(0.5) * (a.re + b.re), T(0.5) * (a.im - b.im));
    if (i == j) h.im = T(0);
    x[i * D + j] = h;
    x[j * D + i] = conj(h);
  }
  g.sync();
}

// ---- exact environment of a LEFT-CANONICAL tensor: direct solve ------------------
// (E_AA - 1) vec(r) = 0 with the row of index (0,0) replaced by tr r = 1.  The rows
// (i,i) of E - 1 sum to zero (left fixed point = identity), so one of them is
// redundant; the system is non-singular iff the eigenvalue 1 is simple.
// A: [d][D][D].  E: n x (n+1) scratch.  x: n entries -> Hermitian trace-1 r.
// eta_out = tr(Phi(r)) (1 for an exact isometry).  Returns ST_OK / ST_SINGULAR.
template <typename T>
QMPS_HDN int env_solve_direct(const Grp& g, const cx<T>* A, int d, int D, cx<T>* E, int ld,
                              cx<T>* x, int* step_row, int* done, T* red, T* eta_out) {
  const int n = D * D;
  build_transfer<T>(g, A, A, d, D, E, ld);
  g.sync();
  for (int e = g.lane; e < n * (n + 1); e += g.size) {
    int row = e / (n + 1), col = e - row * (n + 1);
    cx<T> v;
    if (row == 0) {
      bool diag = (col < n) && (col % (D + 1) == 0);
      v = mk<T>((diag || col == n) ? T(1) : T(0), T(0));
    } else if (col == n) {
      v = mk<T>(0, 0);
    } else {
      v = E[row * ld + col];
      if (row == col) v.re -= T(1);
    }
    E[row * ld + col] = v;
  }
  g.sync();
  int bad = lu_solve_aug<T>(g, E, ld, n, x, step_row, done, T(n) * eps_of<T>::v(), 0);
  hermitise<T>(g, x, D);
  // eta = sum_s tr(A_s r A_s^dagger) = sum_{s,i} (A_s r A_s^dagger)[i][i]
  T part = T(0);
  for (int e = g.lane; e < d * D; e += g.size) {
    int s = e / D, i = e - s * D;
    const cx<T>* row = A + (s * D + i) * D;
    cx<T> acc = mk<T>(0, 0);
    for (int j = 0; j < D; ++j) {
      cx<T> t = mk<T>(0, 0);
      for (int l = 0; l < D; ++l) cmad_c(t, x[j * D + l], row[l]);   // (r A_s^dagger)[j][i]
      cmad(acc, row[j], t);
    }
    part += acc.re;
  }
  *eta_out = group_sum<T>(g, part, red);
  return bad ? ST_SINGULAR : ST_OK;
}

// ---- leading eigenpair of a (mixed) transfer matrix -----------------------------------
// Hessenberg + shifted QR for ALL eigenvalues (what numpy.linalg.eig does), argmax
// |lambda|, then one inverse iteration on E - lambda for the eigenvector.
//   adjoint = 0: right fixed point  r -> sum_s A_s r B_s^dagger         (matrix E)
//   adjoint = 1: left fixed point   l -> sum_s A_s^dagger l B_s         (matrix E^dagger)
// H: n x (n+1) scratch.  x[n]: eigenvector, unit norm; gauge 0: phase fixed so that its trace
// is real non-negative (traceless: largest entry real positive) -- the Hermitian-compatible gauge
// the canonical-form routines need; gauge 1: LAPACK zgeev's convention (the component of largest
// modulus real positive), which is what the one recorded xmps output shows
// (Time Evo.ipynb cells 22-24, tests/golden/ref_notebook_outputs.json).
template <typename T>
QMPS_HDN void build_transfer_adj(const Grp& g, const cx<T>* A, const cx<T>* B, int d, int D,
                                 cx<T>* E, int ld, int adjoint) {
  if (!adjoint) { build_transfer<T>(g, A, B, d, D, E, ld); return; }
  const int n = D * D;
  for (int e = g.lane; e < n * n; e += g.size) {
    int col = e / n, row = e - col * n;            // (row,col) of E; written transposed-conjugated
    int i = row / D, k = row - i * D;
    int j = col / D, l = col - j * D;
    cx<T> acc = mk<T>(0, 0);
    for (int s = 0; s < d; ++s) cmad_c(acc, A[(s * D + i) * D + j], B[(s * D + k) * D + l]);
    E[col * ld + row] = conj(acc);
  }
}

template <typename T>
QMPS_HDN int leading_eigenpair(const Grp& g, const cx<T>* A, const cx<T>* B, int d, int D,
                               int adjoint, int want_vec, cx<T>* H, int ld, cx<T>* w, cx<T>* vv,
                               cx<T>* rc, cx<T>* rs, T* rn, cx<T>* x, int* step_row, int* done,
                               cx<T>* lambda_out, int gauge = 0) {
  const int n = D * D;
  build_transfer_adj<T>(g, A, B, d, D, H, ld, adjoint);
  g.sync();
  hessenberg<T>(g, H, ld, n, vv);
  int fail = hqr_eigenvalues<T>(g, H, ld, n, w, rc, rs, rn);
  g.sync();
  const int kmax = argmax_abs<T>(w, n);
  const cx<T> lam = w[kmax];
  *lambda_out = lam;
  int status = fail ? ST_NO_CONVERGE : ST_OK;
  if (!want_vec) return status;
  g.sync();
  build_transfer_adj<T>(g, A, B, d, D, H, ld, adjoint);
  g.sync();
  for (int i = g.lane; i < n; i += g.size) {
    H[i * ld + i] = H[i * ld + i] - lam;
    // fixed, well-spread right-hand side (never orthogonal to a generic eigenvector)
    T t = T(0.61803398874989485) * T(i + 1);
    t -= floor(t);
    H[i * ld + n] = mk<T>(T(0.5) + t, T(0.25) - T(0.5) * t);
  }
  g.sync();
  T scale = cabs(lam);
  if (!(scale > T(1e-30))) scale = T(1);
  lu_solve_aug<T>(g, H, ld, n, x, step_row, done, eps_of<T>::v() * scale, 1);
  // normalise + phase convention (redundant on all lanes; x is final after the solve's last sync)
  T nrm2 = T(0);
  cx<T> tr = mk<T>(0, 0);
  int big = 0;
  T bigv = T(-1);
  for (int i = 0; i < n; ++i) {
    T a = norm2(x[i]);
    nrm2 += a;
    if (a > bigv) { bigv = a; big = i; }
    if (i % (D + 1) == 0) tr = tr + x[i];
  }
  T inv = T(1) / sqrt(nrm2);
  cx<T> ph;
  T tra = cabs(tr) * inv;
  if (gauge == 0 && tra > T(1e-8)) ph = conj(tr) * (T(1) / cabs(tr));
  else ph = conj(x[big]) * (T(1) / sqrt(bigv));
  ph = ph * inv;
  g.sync();
  for (int i = g.lane; i < n; i += g.size) x[i] = x[i] * ph;
  g.sync();
  return status;
}

// ---- energy  e = Re sum_ab H_ab tr(M_a^dagger M_b r) ------------------------------------
// M: [4][D][D] two-site block (merge(A,A) or merge(A1,A2)), r: D x D (row-major, ld = D).
// tmp: 4*D*D scratch for P_b = M_b r.  hmat[a*4+b].
template <typename T>
QMPS_HDN T energy_from_block(const Grp& g, const cx<T>* M, const cx<T>* r, int D,
                             const cx<T>* hmat, cx<T>* tmp, T* red) {
  const int DD = D * D;
  g.sync();
  for (int e = g.lane; e < 4 * DD; e += g.size) {
    int b = e / DD, ij = e - b * DD, i = ij / D, j = ij - i * D;
    cx<T> acc = mk<T>(0, 0);
    for (int k = 0; k < D; ++k) cmad(acc, M[b * DD + i * D + k], r[k * D + j]);
    tmp[e] = acc;
  }
  g.sync();
  T part = T(0);
  for (int e = g.lane; e < 4 * DD; e += g.size) {
    int b = e / DD, ij = e - b * DD;
    cx<T> q = mk<T>(0, 0);
    for (int a = 0; a < 4; ++a) cmad(q, hmat[b * 4 + a], M[a * DD + ij]);
    part += q.re * tmp[e].re + q.im * tmp[e].im;
  }
  return group_sum<T>(g, part, red);
}

// M[(s1,s2)][i][j] = sum_k A[s1][i][k] B[s2][k][j]   (qmps/time_evolve_tools.py:20-23,
// generalised from the reference's hard-coded bond dimension 2)
template <typename T>
QMPS_HDN void merge_block(const Grp& g, const cx<T>* A, const cx<T>* B, int d1, int d2, int D, cx<T>* M) {
  const int DD = D * D;
  for (int e = g.lane; e < d1 * d2 * DD; e += g.size) {
    int ab = e / DD, ij = e - ab * DD, i = ij / D, j = ij - i * D;
    int s1 = ab / d2, s2 = ab - s1 * d2;
    cx<T> acc = mk<T>(0, 0);
    for (int k = 0; k < D; ++k) cmad(acc, A[(s1 * D + i) * D + k], B[(s2 * D + k) * D + j]);
    M[e] = acc;
  }
}

}  // namespace qmps
