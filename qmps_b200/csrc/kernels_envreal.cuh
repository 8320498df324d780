// qmps_b200 register-resident direct environment solver for D = 4 and D = 8
// (a4/a5 TransferMatrix(A).eigs() + cholesky, qmps/tools.py:176-182; a9 energy,
// qmps/ground_state.py:150-168, 251-266; a12 rotosolve shift fan-out).
//
// One thread per EQUATION of the real n x n system of envreal.cuh (n = D^2: 16 or 64
// threads per problem); each thread keeps its row (n + 1 numbers) in registers for the
// whole elimination.  Gauss-Jordan with partial pivoting and an implicit row permutation:
//   step k: arg-max |m[k]| over the not-yet-used rows by warp shuffles, the winner
//           publishes its row through shared memory (double-buffered: ONE barrier per
//           step), everybody else eliminates column k with n - k FMAs on registers.
// The k loop is fully unrolled so every register index is static; the triangular
// saving (only columns > k are touched) falls out of the unrolling.
// theta -> U -> A (ansatz program), the transfer map, the solve, Cholesky and the
// energy epilogue all stay on chip: HBM traffic is theta in, one number out.
#pragma once
#include <cuda_runtime.h>
#include "envreal.cuh"
#include "kernels_generic.cuh"

namespace qmps {

template <typename T, int D> struct ErLayout { size_t A, Ap, rowbuf, cand, x, r, C, tmp, red, trig, total; };
template <typename T, int D>
QMPS_HD ErLayout<T, D> er_layout(int d, int nops, int want_tmp) {
  constexpr int n = D * D, NT = n, NW = (NT + 31) / 32;
  ErLayout<T, D> L;
  Bump b;
  L.A = b.take(sizeof(cx<T>) * (size_t)d * n);                  // unpadded (ansatz state / energy)
  L.Ap = b.take(sizeof(cx<T>) * (size_t)d * D * (D + 1));       // padded rows for the row build
  L.rowbuf = b.take(sizeof(T) * 2 * NW * (n + 2));
  L.cand = b.take(sizeof(T) * 2 * NW * 4);
  L.x = b.take(sizeof(T) * n);
  L.r = b.take(sizeof(cx<T>) * n);
  L.C = b.take(sizeof(cx<T>) * n);
  L.tmp = b.take(want_tmp ? sizeof(cx<T>) * (size_t)8 * n : 0);
  L.red = b.take(sizeof(T) * NT);
  L.trig = b.take(sizeof(T) * 2 * (nops > 0 ? nops : 1));
  L.total = (b.off + 127) & ~size_t(127);
  return L;
}

template <typename T> struct alignas(2 * sizeof(T)) pair_of { T x, y; };

template <typename T> struct tiny_of;
template <> struct tiny_of<double> { static __device__ __forceinline__ double v() { return 1e-13; } };
template <> struct tiny_of<float> { static __device__ __forceinline__ float v() { return 1e-5f; } };

// MODE 0: eta, r, C, status.  MODE 1: energy (+ status).
// WIDE (D = 8 only): 1 = 255 registers / 4 CTAs per SM with the register-cached row builder, instead
// of 168 registers / 6 CTAs per SM (3 warps on one SM sub-partition) with the streaming one (0) or the
// l-hoisted one (2).
template <typename T, int D, int MODE, int WIDE>
__global__ void __launch_bounds__(D == 8 ? 64 : 128, (D == 8 && WIDE != 1) ? 5 : 4)
env_real_kernel(EnvParams p) {
  constexpr int n = D * D, NT = n, NW = (NT + 31) / 32;
  constexpr int G = NT;                                     // lanes per problem
  constexpr int ROWLD = n + 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  int gi, gpc;
  const Grp g = make_group<G>(&gi, &gpc);
  const int d = p.d;
  const ErLayout<T, D> L = er_layout<T, D>(d, p.nops, MODE == 1);
  unsigned char* base = smem_raw + (size_t)gi * L.total;
  cx<T>* A = reinterpret_cast<cx<T>*>(base + L.A);
  cx<T>* Ap = reinterpret_cast<cx<T>*>(base + L.Ap);
  T* rowbuf = reinterpret_cast<T*>(base + L.rowbuf);
  T* candbuf = reinterpret_cast<T*>(base + L.cand);
  T* xs = reinterpret_cast<T*>(base + L.x);
  cx<T>* r = reinterpret_cast<cx<T>*>(base + L.r);
  cx<T>* Cc = reinterpret_cast<cx<T>*>(base + L.C);
  cx<T>* tmp = reinterpret_cast<cx<T>*>(base + L.tmp);
  T* red = reinterpret_cast<T*>(base + L.red);
  T* trig = reinterpret_cast<T*>(base + L.trig);
  const int S = (MODE == 1 && p.nshift > 0) ? p.nshift : 1;
  const int64_t total = p.N * S;
  const cx<T>* hmat = reinterpret_cast<const cx<T>*>(p.hmat);
  const int e = g.lane;                                      // my equation
  const int wig = (NW > 1) ? (e >> 5) : 0;                   // warp within the group
  const unsigned smask = g.cta ? 0xffffffffu : g.mask;

  // every group runs the same number of iterations (barriers inside): idle groups redo the last problem
  const int64_t stride = (int64_t)gridDim.x * gpc;
  for (int64_t pid0 = (int64_t)blockIdx.x * gpc; pid0 < total; pid0 += stride) {
    int64_t pid = pid0 + gi;
    const bool live = pid < total;
    if (!live) pid = total - 1;
    const int64_t pn = pid / S;
    const int sidx = (int)(pid - pn * S);
    // ---- 1. tensor A[d][D][D] into shared memory
    if (p.theta) {
      StateLayout SL; SL.R = 2 * D; SL.ncols = D; SL.a_layout = 1;
      const double sh = p.nshift > 0 ? p.shifts[sidx] : 0.0;
      ansatz_eval<T>(g, p.ops, p.nops, p.theta + pn * p.P, p.nshift > 0 ? p.coord : -1, sh, p.nq, SL, A, trig);
    } else if (p.in_is_U) {
      const cx<T>* U = reinterpret_cast<const cx<T>*>(p.in) + pn * (size_t)(4 * n);
      for (int q = g.lane; q < 2 * n; q += g.size) {
        int s = q / n, ij = q - s * n, i = ij / D, j = ij - i * D;
        A[q] = U[(2 * i + s) * (2 * D) + j];
      }
    } else {
      const cx<T>* src = reinterpret_cast<const cx<T>*>(p.in) + pn * (size_t)(d * n);
      for (int q = g.lane; q < d * n; q += g.size) A[q] = src[q];
    }
    g.sync();
    for (int q = g.lane; q < d * n; q += g.size) {
      const int si = q / D, j = q - si * D;
      Ap[si * (D + 1) + j] = A[q];
    }
    g.sync();
    // ---- 2. my row of the real system, in registers
    T m[n + 1];
    if (d == 2 && (WIDE == 1 || D < 8)) herm_row_cached<T, D, 2>(Ap, D + 1, e, m);
    else if (d == 2 && WIDE == 2) herm_row_lhoist<T, D, 2>(Ap, D + 1, e, m);
    else herm_row<T, D>(Ap, D + 1, d, e, m);
    // ---- 3. Gauss-Jordan, implicit partial pivoting
    bool done = false;
    int bad = 0;
    int mycol = 0;
    T mypiv = T(1);
#pragma unroll
    for (int k = 0; k < n; ++k) {
      // arg-max of |m[k]| over the unused rows in ONE warp reduction (redux.sync): a 32-bit key =
      // the top 26 bits of |m[k]| as a float (monotonic in the magnitude) | (63 - row).  Partial
      // pivoting only needs a pivot within rounding of the largest, not the exact maximum.
      const T cand_exact = done ? T(-1) : fabs(m[k]);
      unsigned key = done ? 0u : ((__float_as_uint((float)cand_exact) & ~63u) | (unsigned)(63 - e));
      if (!done && key < 64u) key = 64u | (unsigned)(63 - e);       // zero column entry: still eligible
      const unsigned best = __reduce_max_sync(smask, key);
      int who = 63 - (int)(best & 63u);
      T cand = __uint_as_float(best & ~63u);
      T* buf = rowbuf + ((k & 1) * NW + wig) * ROWLD;
      T* cb = candbuf + ((k & 1) * NW + wig) * 4;
      if (e == who) {                                         // local winner publishes its row
        {
          int j = k;
          if (j & 1) { buf[j] = m[j]; ++j; }
#pragma unroll
          for (; j + 1 <= n; j += 2) { pair_of<T> v; v.x = m[j]; v.y = m[j + 1]; *reinterpret_cast<pair_of<T>*>(buf + j) = v; }
          if (j <= n) buf[j] = m[j];
        }
        cb[0] = cand;
        cb[1] = T(who);
      }
      if (NW > 1 && best == 0u && (e & 31) == 0) cb[0] = T(-1);   // this warp has no unused row left
      g.sync();
      int gwho = who;
      const T* prow = buf;
      if (NW > 1) {                                           // best of the warps' candidates
        const T* c0 = candbuf + ((k & 1) * NW) * 4;
        T bc = c0[0]; int bw = 0;
#pragma unroll
        for (int w = 1; w < NW; ++w) { const T cw = c0[4 * w]; if (cw > bc) { bc = cw; bw = w; } }
        gwho = (int)c0[4 * bw + 1];
        prow = rowbuf + ((k & 1) * NW + bw) * ROWLD;
        cand = bc;
      }
      if (!(cand > tiny_of<T>::v())) bad = 1;
      const T pv = prow[k];
      const T inv = T(1) / pv;
      T f = m[k] * inv;
      if (e == gwho) { done = true; mypiv = pv; mycol = k; f = T(0); }
      {
        int j = k + 1;
        if (j & 1) { m[j] -= f * prow[j]; ++j; }
#pragma unroll
        for (; j + 1 <= n; j += 2) {
          const pair_of<T> v = *reinterpret_cast<const pair_of<T>*>(prow + j);
          m[j] -= f * v.x; m[j + 1] -= f * v.y;
        }
        if (j <= n) m[j] -= f * prow[j];
      }
    }
    xs[mycol] = m[n] / mypiv;
    g.sync();
    herm_scatter<T, D>(xs, e, r);
    g.sync();
    int status = bad ? ST_SINGULAR : ST_OK;
    // eta = tr Phi(r) = sum_{s,i} (A_s r A_s^dagger)[i][i]   (1 for an exact isometry)
    T part = T(0);
    if (MODE == 0 && p.eta) {
      for (int q = g.lane; q < d * D; q += g.size) {
        const cx<T>* row = A + q * D;
        cx<T> acc = mk<T>(0, 0);
        for (int j = 0; j < D; ++j) {
          cx<T> t = mk<T>(0, 0);
          for (int l = 0; l < D; ++l) cmad_c(t, r[j * D + l], row[l]);
          cmad(acc, row[j], t);
        }
        part += acc.re;
      }
    }
    if (MODE == 0) {
      T eta_r = T(1);
      if (p.eta) eta_r = group_sum<T>(g, part, red);
      if (p.C || p.status) {
        const int cb = cholesky_lower<T>(g, r, D, Cc, D, D);
        if (cb && status == ST_OK) status = ST_NOT_PD;
      }
      if (live) {
        if (g.lane == 0) {
          if (p.eta) reinterpret_cast<cx<T>*>(p.eta)[pid] = mk<T>(eta_r, 0);
          if (p.status) p.status[pid] = status;
        }
        if (p.r) { cx<T>* o = reinterpret_cast<cx<T>*>(p.r) + pid * (size_t)n; for (int q = g.lane; q < n; q += g.size) o[q] = r[q]; }
        if (p.C) { cx<T>* o = reinterpret_cast<cx<T>*>(p.C) + pid * (size_t)n; for (int q = g.lane; q < n; q += g.size) o[q] = Cc[q]; }
      }
    } else {
      const cx<T>* M;
      if (p.two_site) M = A;
      else { merge_block<T>(g, A, A, 2, 2, D, tmp); M = tmp; }
      const T en = energy_from_block<T>(g, M, r, D, hmat, tmp + 4 * n, red);
      if (p.status) {
        const int cb = cholesky_lower<T>(g, r, D, Cc, D, D);
        if (cb && status == ST_OK) status = ST_NOT_PD;
      }
      if (live && g.lane == 0) {
        reinterpret_cast<T*>(p.energy)[pid] = en;
        if (p.status) p.status[pid] = status;
      }
    }
    g.sync();
  }
}

}  // namespace qmps
