// qmps_b200 leading eigenvalue of the 16 x 16 mixed transfer matrix (D = 4, any d) --
// Map(A,B).right_fixed_point()/left_fixed_point() eigenvalue (xmps; call sites
// qmps/time_evolve_tools.py:87, qmps/loschmidts/time_evo.py:79-82), the Loschmidt / TDVP
// step cost -sqrt|eta| (qmps/loschmidts/time_evo.py:75-116) and the fidelity |eta|^2
// (qmps/time_evolve_tools.py:84-91).  This is the kernel behind BASELINE config 3.
//
// Same algorithm as the generic path (what numpy.linalg.eig does: Householder reduction to
// Hessenberg form, shifted complex QR for ALL eigenvalues, arg-max |lambda| -- Loschmidt
// cusps are level crossings, so a power iteration cannot meet 1e-10), re-laid-out for SIMT:
//   * half a warp per problem, lane j owns COLUMN j of the matrix in registers (16 complex);
//   * left Givens sweeps are column-local; the rotation parameters come from lane i-1 by
//     shuffles (width 16), no barriers;
//   * the right (RQ) sweep is row-local after a transposition through a padded shared tile;
//   * deflation is detected by all sub-diagonal entries at once (one ballot);
//   * the two problems of a warp execute one converged instruction stream: every decision
//     that steers control flow is made warp-uniform, per-problem differences are predicated.
// Code size is what decides this kernel (round-1 version: four fully unrolled sweep variants and an
// unrolled Hessenberg reduction, 63 % of the stall samples were instruction-cache misses,
// profiles/ncu_fp16_r02a.txt): there is ONE sweep body whose unrolled steps are skipped by
// warp-uniform branches outside the union of the two active windows, and the Hessenberg loop is a
// real loop (the reflector is zero above the current column, so every step runs the same
// 16-row body).
#pragma once
#include <cuda_runtime.h>
#include "kernels_generic.cuh"

namespace qmps {

constexpr int F16_N = 16, F16_LD = 17;

// diagnostics (qmps_debug_counter): [0] problems solved, [1] QR sweeps, [2] forced deflations
__device__ unsigned long long g_fp16_dbg[4];

template <typename T> struct Fp16Layout { size_t S, rot, vbuf, ubuf, A, B, total; };
template <typename T> QMPS_HD Fp16Layout<T> fp16_layout(int d) {
  Fp16Layout<T> L;
  Bump b;
  // A / B staging is dead once the columns of E are built and shares the tile S when it fits;
  // the rotation table of the QR phase shares the reflector buffers of the Hessenberg phase.
  const size_t s_bytes = sizeof(cx<T>) * F16_N * F16_LD, ab_bytes = sizeof(cx<T>) * (size_t)d * F16_N;
  L.S = b.take(s_bytes);
  L.vbuf = b.take(sizeof(cx<T>) * 2 * F16_N);
  L.rot = L.vbuf;
  L.ubuf = b.take(sizeof(cx<T>) * F16_N);
  if (2 * ab_bytes <= s_bytes) { L.A = L.S; L.B = L.S + ab_bytes; }
  else { L.A = b.take(ab_bytes); L.B = b.take(ab_bytes); }
  L.total = (b.off + 127) & ~size_t(127);
  return L;
}

template <typename T> __device__ __forceinline__ T rsqrt_t(T x);
template <> __device__ __forceinline__ double rsqrt_t<double>(double x) { return rsqrt(x); }
template <> __device__ __forceinline__ float rsqrt_t<float>(float x) { return rsqrtf(x); }

// branch-free 1/sqrt(x) for a normal positive x: MUFU.RSQ64H seed (2^-22.9) + two Newton steps.  CUDA's rsqrt() carries
// a slow-path call whose branches split the sweep body into basic blocks the scheduler cannot interleave.
template <typename T> __device__ __forceinline__ T rsq_fast(T x);
template <> __device__ __forceinline__ double rsq_fast<double>(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  return y;
}
template <> __device__ __forceinline__ float rsq_fast<float>(float x) { return rsqrtf(x); }
// smallest squared modulus rsq_fast is used on: the MUFU seed flushes subnormal inputs (rsqrt of zero = inf, then NaN in the
// Newton step), so a pair below this is treated as a zero pair (identity rotation) -- it is below 1e-140 (1e-17 in
// complex64) in modulus and far beneath the negligibility threshold of any matrix the sweeps see
template <typename T> struct rsq_floor;
template <> struct rsq_floor<double> { static __device__ __forceinline__ double v() { return 1e-280; } };
template <> struct rsq_floor<float> { static __device__ __forceinline__ float v() { return 1e-34f; } };

template <typename T> __device__ __forceinline__ cx<T> shfl16(cx<T> v, int src) {
  cx<T> r;
  r.re = __shfl_sync(0xffffffffu, v.re, src, 16);
  r.im = __shfl_sync(0xffffffffu, v.im, src, 16);
  return r;
}

// One explicit shifted-QR sweep on the window [l, en] of each half-warp's matrix.
// In: the row-major copy S of the matrix.  Out: S again (what the deflation test and the next shift read).
// The registers h[] are live inside the sweep only.  Steps i <= lo or i > hi are skipped by warp-uniform
// branches (lo = min l, hi = max en over the two problems of the warp).  Everything that addresses the
// matrix by a lane-dependent index (the diagonal shift, the exact zero at H[l][l-1]) is done on the
// shared copy -- in registers it would cost a compare-and-select per row.
template <typename T>
__device__ __forceinline__ void fp16_sweep(int ln, int l, int en, int lo, int hi, cx<T> sigma, cx<T>* S,
                                           cx<T>* rot) {
  const bool in_win = (ln >= l) && (ln <= en);
  if (in_win) S[ln * F16_LD + ln] = S[ln * F16_LD + ln] - sigma;          // H - sigma on the window's diagonal
  if (l >= 1 && ln == l - 1) S[l * F16_LD + (l - 1)] = mk<T>(0, 0);        // the negligible entry becomes exact
  __syncwarp();
  cx<T> h[F16_N];
#pragma unroll
  for (int i = 0; i < F16_N; ++i)
    if (i <= hi) h[i] = S[i * F16_LD + ln];                                 // my column
  __syncwarp();
  // left phase: R = G_en ... G_{l+1} (H - sigma), column-local
#pragma unroll
  for (int i = 1; i < F16_N; ++i) {
    if (i > lo && i <= hi) {
      const bool act = (i > l) && (i <= en);
      const cx<T> f = shfl16(h[i - 1], i - 1), g = shfl16(h[i], i - 1);
      const T nr2 = norm2(f) + norm2(g);
      cx<T> c = mk<T>(1, 0), s = mk<T>(0, 0);
      T nr = T(0);
      if (act && nr2 > T(0)) {
        const T inr = rsqrt_t<T>(nr2);
        c = f * inr; s = g * inr; nr = nr2 * inr;
      }
      if (ln == 0) { rot[2 * i] = c; rot[2 * i + 1] = s; }
      const cx<T> p = h[i - 1], q = h[i];
      cx<T> top = conj(c) * p; cmad(top, conj(s), q);
      cx<T> bot = c * q; cmsub(bot, s, p);
      h[i - 1] = top; h[i] = bot;
      if (act && ln == i - 1) { h[i - 1] = mk<T>(nr, 0); h[i] = mk<T>(0, 0); }
    }
  }
  // transpose: columns -> rows (rows / columns beyond hi are never read again)
#pragma unroll
  for (int i = 0; i < F16_N; ++i)
    if (i <= hi) S[i * F16_LD + ln] = h[i];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < F16_N; ++j)
    if (j <= hi) h[j] = S[ln * F16_LD + j];
  // right phase: H' = R G_{l+1}^H ... G_en^H + sigma, row-local (rot[] is uniform per problem)
#pragma unroll
  for (int j = 1; j < F16_N; ++j) {
    if (j > lo && j <= hi) {
      const cx<T> c = rot[2 * j], s = rot[2 * j + 1];
      const cx<T> xx = h[j - 1], yy = h[j];
      cx<T> a = xx * c; cmad(a, yy, s);
      cx<T> b = yy * conj(c); cmsub(b, xx, conj(s));
      h[j - 1] = a; h[j] = b;
    }
  }
  // my row back to the shared copy (same thread reads and writes row ln here), then + sigma on the diagonal
#pragma unroll
  for (int j = 0; j < F16_N; ++j)
    if (j <= hi) S[ln * F16_LD + j] = h[j];
  if (in_win) S[ln * F16_LD + ln] = S[ln * F16_LD + ln] + sigma;
}

template <typename T, int MINB>
__global__ void __launch_bounds__(128, MINB)
fp16_kernel(FpParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int d = p.d;
  const Fp16Layout<T> L = fp16_layout<T>(d);
  const int half = (threadIdx.x >> 4) & 1;
  const int ln = threadIdx.x & 15;                       // lane within the problem = my column
  const int gi = threadIdx.x >> 4, gpc = blockDim.x >> 4;
  unsigned char* base = smem_raw + (size_t)gi * L.total;
  cx<T>* S = reinterpret_cast<cx<T>*>(base + L.S);
  cx<T>* rot = reinterpret_cast<cx<T>*>(base + L.rot);
  cx<T>* vbuf = reinterpret_cast<cx<T>*>(base + L.vbuf);   // [0..15] raw column, [16..31] scaled reflector
  cx<T>* ubuf = reinterpret_cast<cx<T>*>(base + L.ubuf);
  cx<T>* As = reinterpret_cast<cx<T>*>(base + L.A);
  cx<T>* Bs = reinterpret_cast<cx<T>*>(base + L.B);
  const T eps = eps_of<T>::v();
  const int maxit = 60;
  const size_t tsz = (size_t)d * F16_N;

  const int64_t stride = (int64_t)gridDim.x * gpc;
  for (int64_t pid0 = (int64_t)blockIdx.x * gpc; pid0 < p.N; pid0 += stride) {
    int64_t pid = pid0 + gi;
    const bool live = pid < p.N;
    if (!live) pid = p.N - 1;
    int64_t ia, ib;
    if (p.pair_mode == 1) { ia = pid / p.NB; ib = pid - ia * p.NB; }
    else if (p.pair_mode == 2) { ib = pid / p.NA; ia = pid - ib * p.NA; }
    else { ia = pid < p.NA ? pid : p.NA - 1; ib = pid < p.NB ? pid : p.NB - 1; }
    const cx<T>* Ag = reinterpret_cast<const cx<T>*>(p.A) + ia * tsz;
    const cx<T>* Bg = reinterpret_cast<const cx<T>*>(p.B) + ib * tsz;
    for (int q = ln; q < d * F16_N; q += 16) { As[q] = Ag[q]; Bs[q] = Bg[q]; }
    __syncwarp();
    // ---- my column of E (or of E^dagger for the left fixed point)
    cx<T> h[F16_N];
#pragma unroll
    for (int r = 0; r < F16_N; ++r) h[r] = mk<T>(0, 0);
    // right: E[(i,k),(j,l)] = sum_s A[s,i,j] conj(B[s,k,l]), my column (j,l) = (jj,ll);
    // left:  E^dagger[(i,k),(j,l)] = conj(A[s,j,i] conj(B[s,l,k])): same products, transposed reads, conjugated
    const int jj = ln >> 2, ll = ln & 3;
    const int sa = p.left ? 1 : 4, oa = p.left ? jj * 4 : jj, ob = p.left ? ll * 4 : ll;
#pragma unroll 1
    for (int s = 0; s < d; ++s) {
      cx<T> a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[s * 16 + i * sa + oa]; b[i] = Bs[s * 16 + i * sa + ob]; }
#pragma unroll
      for (int r = 0; r < F16_N; ++r) cmad_c(h[r], a[r >> 2], b[r & 3]);
    }
    if (p.left) {
#pragma unroll
      for (int r = 0; r < F16_N; ++r) h[r].im = -h[r].im;
    }
    __syncwarp();                                          // A / B staging may share the tile S
    // ---- Householder reduction to Hessenberg form, column-local left updates.
    // A real loop: the reflector v has v[i] = 0 for i <= k, so every step runs the same 16-row body.
#pragma unroll 1
    for (int k = 0; k + 2 < F16_N; ++k) {
      if (ln == k) {
#pragma unroll
        for (int i = 0; i < F16_N; ++i) vbuf[i] = h[i];
      }
      __syncwarp();
      const cx<T> alpha = vbuf[k + 1];
      // |x|^2 of the entries below the sub-diagonal: one entry per lane, butterfly sum over the half-warp
      T xn2 = (ln > k + 1) ? norm2(vbuf[ln]) : T(0);
#pragma unroll
      for (int m = 8; m >= 1; m >>= 1) xn2 += __shfl_xor_sync(0xffffffffu, xn2, m, 16);
      const bool skip = (xn2 == T(0)) && (alpha.im == T(0));
      T beta = sqrt(norm2(alpha) + xn2);
      if (alpha.re > T(0)) beta = -beta;
      cx<T> tau = mk<T>(0, 0), scal = mk<T>(0, 0);
      if (!skip) {
        const T ib = T(1) / beta;
        tau = mk<T>((beta - alpha.re) * ib, -alpha.im * ib);
        scal = cinv(alpha - mk<T>(beta, 0));
      }
      // scaled reflector v (v[k+1] = 1) for everybody; my own component vj
      cx<T> vj = mk<T>(0, 0);
      if (ln == k + 1) vj = mk<T>(1, 0);
      else if (ln > k + 1) vj = vbuf[ln] * scal;
      vbuf[16 + ln] = vj;
      __syncwarp();
      // left:  H <- (1 - conj(tau) v v^H) H   on my column
      cx<T> w = mk<T>(0, 0);
#pragma unroll
      for (int i = 1; i < F16_N; ++i) cmad(w, conj(vbuf[16 + i]), h[i]);
      w = w * conj(tau);
#pragma unroll
      for (int i = 1; i < F16_N; ++i) cmsub(h[i], vbuf[16 + i], w);
      {
        // column k below the diagonal becomes (beta, 0, ..., 0): selects on every row (a store under
        // "i == k + 1" would be turned into a dynamically indexed store and push h[] to local memory)
        const bool fix = !skip && ln == k;
#pragma unroll
        for (int i = 1; i < F16_N; ++i) {
          const bool sub = fix && (i == k + 1), below = fix && (i > k + 1);
          h[i].re = sub ? beta : (below ? T(0) : h[i].re);
          h[i].im = (sub || below) ? T(0) : h[i].im;
        }
      }
      // right: H <- H (1 - tau v v^H):  u = H v (row sums through the shared tile), H -= tau u v^H
#pragma unroll
      for (int i = 0; i < F16_N; ++i) S[i * F16_LD + ln] = h[i] * vj;
      __syncwarp();
      cx<T> u = mk<T>(0, 0);
#pragma unroll
      for (int j = 1; j < F16_N; ++j) u = u + S[ln * F16_LD + j];
      ubuf[ln] = u * tau;
      __syncwarp();
      const cx<T> cvj = conj(vj);
#pragma unroll
      for (int i = 0; i < F16_N; ++i) cmsub(h[i], ubuf[i], cvj);
      __syncwarp();
    }
    // row-major copy for the first deflation test
#pragma unroll
    for (int i = 0; i < F16_N; ++i) S[i * F16_LD + ln] = h[i];
    __syncwarp();

    // ---- shifted QR, all eigenvalues; keep the one of largest modulus
    int en = F16_N - 1, its = 0, fail = 0, sweeps = 0;
    T best2 = T(-1);
    cx<T> best = mk<T>(0, 0);
#pragma unroll 1
    for (;;) {
      // negligible sub-diagonal entries, all at once
      bool neg = false;
      if (ln >= 1) {
        T sc = cabs1(S[(ln - 1) * F16_LD + (ln - 1)]) + cabs1(S[ln * F16_LD + ln]);
        if (sc == T(0)) sc = T(1);
        neg = cabs1(S[ln * F16_LD + (ln - 1)]) <= eps * sc;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, neg);
      const unsigned bits = (bal >> (16 * half)) & 0xffffu;
      int l = 0;
      while (en >= 0) {
        const unsigned m = bits & ((2u << en) - 1u) & ~1u;
        l = m ? (31 - __clz(m)) : 0;
        if (l == en || its >= maxit) {
          if (l != en) fail = 1;
          const cx<T> ev = S[en * F16_LD + en];
          const T a2 = norm2(ev);
          if (a2 > best2) { best2 = a2; best = ev; }
          --en; its = 0;
        } else break;
      }
      if (__all_sync(0xffffffffu, en < 0)) break;
      // shift (Wilkinson; exceptional every 10 stalled sweeps) -- idle problem: l = en = 0, sigma = 0
      cx<T> sigma = mk<T>(0, 0);
      int lw = 0, enw = 0;
      if (en >= 1) {
        lw = l; enw = en;
        const cx<T> a = S[(en - 1) * F16_LD + (en - 1)], b = S[(en - 1) * F16_LD + en];
        const cx<T> c = S[en * F16_LD + (en - 1)], dd = S[en * F16_LD + en];
        if (its == 10 || its == 20 || its == 30 || its == 40) {
          const T t = fabs(c.re) + (en >= 2 ? fabs(S[(en - 1) * F16_LD + (en - 2)].re) : T(0));
          sigma = dd + mk<T>(t, 0);
        } else {
          sigma = dd;
          const cx<T> bc = b * c;
          if (bc.re != T(0) || bc.im != T(0)) {
            const cx<T> y = (a - dd) * T(0.5);
            cx<T> z = csqrt(y * y + bc);
            if (y.re * z.re + y.im * z.im < T(0)) z = -z;
            sigma = dd - cdiv(bc, y + z);
          }
        }
      }
      // union of the two active windows of this warp (an idle problem has the empty window [15, 0])
      const int lo_mine = (en >= 1) ? lw : F16_N - 1, hi_mine = (en >= 1) ? enw : 0;
      // redux.sync results live in uniform registers: the window branches below are provably warp-uniform
      const int lo = __reduce_min_sync(0xffffffffu, lo_mine);
      const int hi = __reduce_max_sync(0xffffffffu, hi_mine);
      __syncwarp();
      fp16_sweep<T>(ln, lw, enw, lo, hi, sigma, S, rot);
      __syncwarp();
      ++its;
      if (en >= 1) ++sweeps;
    }
    if (live && ln == 0) {
      atomicAdd(&g_fp16_dbg[0], 1ull);
      atomicAdd(&g_fp16_dbg[1], (unsigned long long)sweeps);
      if (fail) atomicAdd(&g_fp16_dbg[2], 1ull);
      const T a2 = norm2(best);
      if (p.eta) reinterpret_cast<cx<T>*>(p.eta)[pid] = best;
      if (p.cost) reinterpret_cast<T*>(p.cost)[pid] = -sqrt(sqrt(a2));
      if (p.echo) reinterpret_cast<T*>(p.echo)[pid] = -log(a2);
      if (p.fid) reinterpret_cast<T*>(p.fid)[pid] = a2;
      if (p.status) p.status[pid] = fail ? ST_NO_CONVERGE : ST_OK;
    }
    __syncwarp();
  }
}

}  // namespace qmps
