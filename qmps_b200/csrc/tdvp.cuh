// qmps_b200 classical iTDVP for a single-site uniform MPS: the tangent vector
// iMPS([A]).dA_dt([h]) (xmps; call sites scripts/classical_time_evolution.py:22-26,
// scripts/mixed_environment.py:41, Time Evo.ipynb), per problem, executed by a cooperating group like
// the routines of core.cuh / generic.cuh.  xmps is not vendored: this is the published gauge-fixed
// tangent vector of a uniform MPS with a nearest-neighbour Hamiltonian (Haegeman et al., PRL 107,
// 070601), restated in oracle/tdvp.py and checked there against energy / norm conservation, exact
// single-site dynamics and the analytic TFIM Loschmidt rate.
//
// Left-canonical gauge (l = 1, r = trace-1 right fixed point), C^{st} = sum h[(s,t),(s',t')] A_s' A_t':
//     H_l = sum_st (A_s A_t)^dagger C^{st},  e = tr(H_l r)
//     K - sum_s A_s^dagger K A_s = H_l - e,  tr(K r) = 0
//     G^s = sum_t C^{st} r A_t^dagger r^-1 + sum_t A_t^dagger C^{ts} + K A_s
//     dA^s/dt = f (G^s - A_s sum_u A_u^dagger G^u),   f = -i (real time) or -1 (imaginary time)
#pragma once
#include "generic.cuh"
#include "canon.cuh"

namespace qmps {

// Tangent vector of a LEFT-CANONICAL tensor.  A: [d][D][D] in group memory; h: [d*d][d*d] (any space);
// E: n x (n+1) scratch; the other pointers: group scratch of the sizes of tdvp_layout.
// out: [d][D][D] (any space).  Returns ST_OK, ST_SINGULAR (degenerate fixed point) or ST_NOT_PD (r singular).
template <typename T>
QMPS_HDN int tdvp_tangent_problem(const Grp& g, const cx<T>* A, const cx<T>* h, int imaginary, int d, int D,
                                  cx<T>* E, cx<T>* x, cx<T>* r, cx<T>* K, cx<T>* rinv, cx<T>* Cc, cx<T>* Ci,
                                  cx<T>* Hl, cx<T>* AA, cx<T>* C, cx<T>* G, int* step_row, int* done, T* red,
                                  cx<T>* out, T* energy_out) {
  const int n = D * D, ld = n + 1, dd = d * d;
  T eta;
  int status = env_solve_direct<T>(g, A, d, D, E, ld, x, step_row, done, red, &eta);
  for (int e = g.lane; e < n; e += g.size) r[e] = x[e];
  // AA[(s,t)] = A_s A_t ;  C[(a,b)] = sum_(c,d) h[(a,b),(c,d)] AA[(c,d)]
  merge_block<T>(g, A, A, d, d, D, AA);
  g.sync();
  for (int e = g.lane; e < dd * n; e += g.size) {
    const int ab = e / n, ik = e - ab * n;
    cx<T> acc = mk<T>(0, 0);
    for (int cd = 0; cd < dd; ++cd) cmad(acc, h[ab * dd + cd], AA[cd * n + ik]);
    C[e] = acc;
  }
  g.sync();
  // Hl[i][k] = sum_(st) sum_j conj(AA[st][j][i]) C[st][j][k]
  for (int e = g.lane; e < n; e += g.size) {
    const int i = e / D, k = e - i * D;
    cx<T> acc = mk<T>(0, 0);
    for (int st = 0; st < dd; ++st)
      for (int j = 0; j < D; ++j) cmad(acc, conj(AA[st * n + j * D + i]), C[st * n + j * D + k]);
    Hl[e] = acc;
  }
  g.sync();
  T part = T(0);
  for (int e = g.lane; e < n; e += g.size) {
    const int i = e / D, k = e - i * D;
    const cx<T> a = Hl[e], b = r[k * D + i];
    part += a.re * b.re - a.im * b.im;
  }
  const T en = group_sum<T>(g, part, red);
  // (1 - E_left) K = Hl - e, first equation replaced by tr(K r) = 0
  build_transfer_adj<T>(g, A, A, d, D, E, ld, 1);
  g.sync();
  for (int e = g.lane; e < n * (n + 1); e += g.size) {
    const int row = e / (n + 1), col = e - row * (n + 1);
    cx<T> v;
    if (row == 0) {
      if (col == n) v = mk<T>(0, 0);
      else { const int j = col / D, l = col - j * D; v = r[l * D + j]; }
    } else if (col == n) {
      v = Hl[row];
      if (row % (D + 1) == 0) v.re -= en;
    } else {
      v = -E[row * ld + col];
      if (row == col) v.re += T(1);
    }
    E[row * ld + col] = v;
  }
  g.sync();
  if (lu_solve_aug<T>(g, E, ld, n, x, step_row, done, T(n) * eps_of<T>::v(), 0) && status == ST_OK) status = ST_SINGULAR;
  for (int e = g.lane; e < n; e += g.size) K[e] = x[e];
  // r^-1 = C^-dagger C^-1 with r = C C^dagger
  g.sync();
  if (cholesky_lower<T>(g, r, D, Cc, D, D) && status == ST_OK) status = ST_NOT_PD;
  g.sync();
  tri_inverse<T>(g, Cc, Ci, D, 0);
  g.sync();
  for (int e = g.lane; e < n; e += g.size) {
    const int i = e / D, j = e - i * D;
    cx<T> acc = mk<T>(0, 0);
    for (int k = (i > j ? i : j); k < D; ++k) cmad(acc, conj(Ci[k * D + i]), Ci[k * D + j]);   // Ci is lower triangular
    rinv[e] = acc;
  }
  // T1^s = sum_t C^{st} r A_t^dagger  (into AA, which is free now: d*n <= d*d*n);  X^{st} = C^{st} r goes through x
  g.sync();
  for (int s = 0; s < d; ++s) {
    for (int e = g.lane; e < n; e += g.size) AA[s * n + e] = mk<T>(0, 0);
    for (int t = 0; t < d; ++t) {
      g.sync();
      for (int e = g.lane; e < n; e += g.size) {              // x = C^{st} r
        const int i = e / D, l = e - i * D;
        cx<T> acc = mk<T>(0, 0);
        for (int k = 0; k < D; ++k) cmad(acc, C[(s * d + t) * n + i * D + k], r[k * D + l]);
        x[e] = acc;
      }
      g.sync();
      for (int e = g.lane; e < n; e += g.size) {              // += x A_t^dagger
        const int i = e / D, m = e - i * D;
        cx<T> acc = AA[s * n + e];
        for (int l = 0; l < D; ++l) cmad_c(acc, x[i * D + l], A[t * n + m * D + l]);
        AA[s * n + e] = acc;
      }
    }
  }
  g.sync();
  // G^s = T1^s rinv + sum_t A_t^dagger C^{ts} + K A_s
  for (int e = g.lane; e < d * n; e += g.size) {
    const int s = e / n, ij = e - s * n, i = ij / D, j = ij - i * D;
    cx<T> acc = mk<T>(0, 0);
    for (int m = 0; m < D; ++m) cmad(acc, AA[s * n + i * D + m], rinv[m * D + j]);
    for (int t = 0; t < d; ++t)
      for (int k = 0; k < D; ++k) cmad(acc, conj(A[t * n + k * D + i]), C[(t * d + s) * n + k * D + j]);
    for (int k = 0; k < D; ++k) cmad(acc, K[i * D + k], A[s * n + k * D + j]);
    G[e] = acc;
  }
  g.sync();
  // P = sum_s A_s^dagger G^s  (into Hl);  out^s = f (G^s - A_s P)
  for (int e = g.lane; e < n; e += g.size) {
    const int i = e / D, j = e - i * D;
    cx<T> acc = mk<T>(0, 0);
    for (int s = 0; s < d; ++s)
      for (int k = 0; k < D; ++k) cmad(acc, conj(A[s * n + k * D + i]), G[s * n + k * D + j]);
    Hl[e] = acc;
  }
  g.sync();
  for (int e = g.lane; e < d * n; e += g.size) {
    const int s = e / n, ij = e - s * n, i = ij / D, j = ij - i * D;
    cx<T> acc = G[e];
    for (int k = 0; k < D; ++k) cmsub(acc, A[s * n + i * D + k], Hl[k * D + j]);
    out[e] = imaginary ? -acc : mk<T>(acc.im, -acc.re);        // -1 * acc  or  -i * acc
  }
  if (energy_out && g.lane == 0) *energy_out = en;
  g.sync();
  return status;
}

// out_s = scale * X^-1 B_s X for an UPPER triangular X (the L of left_canonicalise): a tangent vector taken in
// the canonical gauge back in the gauge of the tensor it belongs to.  sB (d D^2), sX, sXi, sT (D^2 each): scratch.
template <typename T>
QMPS_HDN void gauge_back_problem(const Grp& g, const cx<T>* b, const cx<T>* x, T scale, int d, int D,
                                 cx<T>* sB, cx<T>* sX, cx<T>* sXi, cx<T>* sT, cx<T>* out) {
  const int DD = D * D;
  for (int e = g.lane; e < d * DD; e += g.size) sB[e] = b[e];
  for (int e = g.lane; e < DD; e += g.size) sX[e] = x[e];
  g.sync();
  tri_inverse<T>(g, sX, sXi, D, 1);
  g.sync();
  for (int s = 0; s < d; ++s) {
    for (int e = g.lane; e < DD; e += g.size) {
      const int i = e / D, j = e - i * D;
      cx<T> acc = mk<T>(0, 0);
      for (int k = i; k < D; ++k) cmad(acc, sXi[i * D + k], sB[s * DD + k * D + j]);
      sT[e] = acc;
    }
    g.sync();
    for (int e = g.lane; e < DD; e += g.size) {
      const int i = e / D, j = e - i * D;
      cx<T> acc = mk<T>(0, 0);
      for (int k = 0; k <= j; ++k) cmad(acc, sT[i * D + k], sX[k * D + j]);
      out[s * DD + e] = acc * scale;
    }
    g.sync();
  }
}

}  // namespace qmps
