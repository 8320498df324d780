// qmps_b200 canonical forms and local expectation values, per-problem algorithms
// (SURVEY 8(f)-1).  __host__ __device__ templates over a cooperating group, like core.cuh:
// the CUDA kernels in kernels_canon.cuh run them with one warp per problem, the CPU test
// harness (tests/host_emu) with a group of one lane.
//   iMPS([A]).left_canonicalise()   (call sites qmps/time_evolve_tools.py:85-86,
//                                    qmps/loschmidts/time_evo.py:76,143)
//   iMPS([A]).mixed() -> AL, AR, C  (qmps/tools.py:184-186, tests/test_represent.py:18-31)
//   iMPS([A]).Es(ops)               (qmps/loschmidts/time_evo.py:144, tests/test_represent.py:37)
// xmps itself is not vendored; the gauge used here is the Cholesky gauge of
// oracle/tensors.py:left_canonicalise / oracle/canonical.py (unique for a PD fixed point).
#pragma once
#include "core.cuh"

namespace qmps {

// inverse of a triangular D x D matrix G (row-major, ld = D) into Gi; a lane owns whole columns
template <typename T>
QMPS_HDN void tri_inverse(const Grp& g, const cx<T>* G, cx<T>* Gi, int D, int upper) {
  for (int j = g.lane; j < D; j += g.size) {
    for (int i = 0; i < D; ++i) Gi[i * D + j] = mk<T>(0, 0);
    Gi[j * D + j] = cinv(G[j * D + j]);
    if (upper) {
      for (int i = j - 1; i >= 0; --i) {
        cx<T> s = mk<T>(0, 0);
        for (int k = i + 1; k <= j; ++k) cmad(s, G[i * D + k], Gi[k * D + j]);
        Gi[i * D + j] = -(s * cinv(G[i * D + i]));
      }
    } else {
      for (int i = j + 1; i < D; ++i) {
        cx<T> s = mk<T>(0, 0);
        for (int k = j; k < i; ++k) cmad(s, G[i * D + k], Gi[k * D + j]);
        Gi[i * D + j] = -(s * cinv(G[i * D + i]));
      }
    }
  }
}

// Gauge transform of one tensor.
//   x_kind 0: x = l Hermitian PD up to a positive scale (left fixed point, sum_s A_s^dagger l A_s = eta l):
//             L upper triangular with L^dagger L = l D / tr(l);  out_s = L A_s L^-1 * scale
//   x_kind 1: x = C lower triangular;                            out_s = C^-1 A_s C * scale
// a, x, out, gout: global (or host) memory; sA (d D^2), sG, sGi, sT (D^2 each): group scratch.
// Returns ST_OK or ST_NOT_PD.
template <typename T>
QMPS_HDN int gauge_problem(const Grp& g, const cx<T>* a, const cx<T>* x, int x_kind, T scale, int d, int D,
                           cx<T>* sA, cx<T>* sG, cx<T>* sGi, cx<T>* sT, cx<T>* out, cx<T>* gout) {
  const int DD = D * D;
  int st = ST_OK;
  for (int e = g.lane; e < d * DD; e += g.size) sA[e] = a[e];
  if (x_kind == 0) {
    T tr = T(0);
    for (int i = 0; i < D; ++i) tr += x[i * D + i].re;
    const T sc = tr > T(0) ? T(D) / tr : T(1);
    for (int e = g.lane; e < DD; e += g.size) {
      const int i = e / D, j = e - i * D;
      const cx<T> u = x[e], v = conj(x[j * D + i]);           // symmetrise
      sT[e] = (u + v) * (T(0.5) * sc);
    }
    g.sync();
    if (cholesky_lower<T>(g, sT, D, sGi, D, D)) st = ST_NOT_PD;   // l = M M^dagger
    g.sync();
    for (int e = g.lane; e < DD; e += g.size) {                  // L = M^dagger
      const int i = e / D, j = e - i * D;
      sG[e] = conj(sGi[j * D + i]);
    }
    g.sync();
    tri_inverse<T>(g, sG, sGi, D, 1);
  } else {
    for (int e = g.lane; e < DD; e += g.size) sG[e] = x[e];
    g.sync();
    tri_inverse<T>(g, sG, sGi, D, 0);
  }
  g.sync();
  const cx<T>* Pm = x_kind == 0 ? sG : sGi;    // left factor
  const cx<T>* Qm = x_kind == 0 ? sGi : sG;    // right factor
  for (int s = 0; s < d; ++s) {
    for (int e = g.lane; e < DD; e += g.size) {
      const int i = e / D, j = e - i * D;
      cx<T> acc = mk<T>(0, 0);
      for (int k = 0; k < D; ++k) cmad(acc, Pm[i * D + k], sA[s * DD + k * D + j]);
      sT[e] = acc;
    }
    g.sync();
    for (int e = g.lane; e < DD; e += g.size) {
      const int i = e / D, j = e - i * D;
      cx<T> acc = mk<T>(0, 0);
      for (int k = 0; k < D; ++k) cmad(acc, sT[i * D + k], Qm[k * D + j]);
      out[s * DD + e] = acc * scale;
    }
    g.sync();
  }
  if (gout) for (int e = g.lane; e < DD; e += g.size) gout[e] = sG[e];
  g.sync();
  return st;
}

// Single-site expectation values of one tensor:
//   out[o] = sum_st O_o[s][t] sum_{ikjl} conj(l[i][k]) A_t[i][j] r[j][l] conj(A_s[k][l]) / (eta sum_ik conj(l[i][k]) r[i][k])
// (l: sum_s A_s^dagger l A_s = eta l, the matrix convention of qmps_fixed_point(left = 1));
// lvec = eta = null: l = 1, eta = 1 (left-canonical A, tr r = 1): out[o] = sum_st O[s][t] tr(A_t r A_s^dagger).
// scratch: sA (d D^2), sR, sL (D^2), sP (d D^2), sQ (d^2 D), sM (d^2).
template <typename T>
QMPS_HDN void expect_problem(const Grp& g, const cx<T>* a, const cx<T>* r, const cx<T>* lvec, const cx<T>* eta,
                             const cx<T>* ops, int nops, int d, int D, cx<T>* sA, cx<T>* sR, cx<T>* sL,
                             cx<T>* sP, cx<T>* sQ, cx<T>* sM, cx<T>* out) {
  const int DD = D * D;
  for (int e = g.lane; e < d * DD; e += g.size) sA[e] = a[e];
  for (int e = g.lane; e < DD; e += g.size) {
    sR[e] = r[e];
    if (lvec) sL[e] = lvec[e];
  }
  g.sync();
  for (int e = g.lane; e < d * DD; e += g.size) {            // P_t = A_t r
    const int t = e / DD, il = e - t * DD, i = il / D, l = il - i * D;
    cx<T> acc = mk<T>(0, 0);
    for (int j = 0; j < D; ++j) cmad(acc, sA[t * DD + i * D + j], sR[j * D + l]);
    sP[e] = acc;
  }
  g.sync();
  for (int e = g.lane; e < d * d * D; e += g.size) {         // partial traces over row i
    const int ts = e / D, i = e - ts * D, t = ts / d, s = ts - t * d;
    cx<T> acc = mk<T>(0, 0);
    if (lvec) {
      for (int k = 0; k < D; ++k) {
        cx<T> q = mk<T>(0, 0);
        for (int l = 0; l < D; ++l) cmad_c(q, sP[t * DD + i * D + l], sA[s * DD + k * D + l]);
        cmad(acc, conj(sL[i * D + k]), q);
      }
    } else {
      for (int l = 0; l < D; ++l) cmad_c(acc, sP[t * DD + i * D + l], sA[s * DD + i * D + l]);
    }
    sQ[e] = acc;
  }
  g.sync();
  for (int ts = g.lane; ts < d * d; ts += g.size) {
    cx<T> acc = mk<T>(0, 0);
    for (int i = 0; i < D; ++i) acc = acc + sQ[ts * D + i];
    sM[ts] = acc;                                             // R[t][s]
  }
  g.sync();
  cx<T> den = mk<T>(1, 0);
  if (lvec) {
    den = mk<T>(0, 0);
    for (int e = 0; e < DD; ++e) cmad(den, conj(sL[e]), sR[e]);
  }
  if (eta) den = den * eta[0];
  const cx<T> iden = cinv(den);
  for (int o = g.lane; o < nops; o += g.size) {
    const cx<T>* O = ops + (size_t)o * d * d;
    cx<T> acc = mk<T>(0, 0);
    for (int s = 0; s < d; ++s)
      for (int t = 0; t < d; ++t) cmad(acc, O[s * d + t], sM[t * d + s]);
    out[o] = acc * iden;
  }
  g.sync();
}

}  // namespace qmps
