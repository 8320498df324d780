// qmps_b200 leading eigenvalue of the 16 x 16 mixed transfer matrix (D = 4, any d), shared-resident form.
// Same algorithm and outputs as kernels_fp16.cuh (Householder -> Hessenberg, shifted complex QR for all
// eigenvalues, arg-max |lambda|: what numpy.linalg.eig does for xmps' Map(A,B).right/left_fixed_point;
// cost -sqrt|eta| of qmps/loschmidts/time_evo.py:75-116, fidelity |eta|^2 of qmps/time_evolve_tools.py:84-91).
//
// What the register-resident forms taught (profiles/ncu_fp16_r02b.txt, ncu_fp16_r02c.txt): the QR sweep is a
// serial chain (shuffle -> norm -> rsqrt -> scale -> rotate -> shuffle ...), so one warp issues an
// instruction every ~6 cycles whatever the layout, and with 16 x 16 complex128 = 4 KB of registers per
// problem only 22-28 problems fit on an SM.  Throughput = problems in flight / chain latency, so this form
// spends as little on-chip state per problem as possible:
//   * the matrix lives in shared memory (16 x 16, rows skewed by their index so that both "lane = column" and
//     "lane = row" accesses are conflict-free: 4.5 KB per complex128 problem with the rotation table),
//     48 problems per SM instead of 22;
//   * a lane carries ONE element through a phase (the entry of its column in the row being rotated, or of
//     its row in the column being rotated), ~80 registers per thread, 24 warps per SM;
//   * every loop is a real loop over the union [lo, hi] of the two active windows of the warp -- no
//     unrolled variants, the whole kernel is ~1.5 k instructions and stays in the instruction cache;
//   * half a warp per problem, two problems per warp in one converged instruction stream.
#pragma once
#include <cuda_runtime.h>
#include "kernels_fp16.cuh"

namespace qmps {

// element (i, j) of the swizzled 16 x 16 tile
#define F16S(i, j) (((i) << 4) + (((j) + (i)) & 15))

template <typename T> struct Fp16sLayout { size_t S, rot, total; };
template <typename T> QMPS_HD Fp16sLayout<T> fp16s_layout() {
  Fp16sLayout<T> L;
  L.S = 0;
  L.rot = sizeof(cx<T>) * 256;                  // 32 entries: (c, s) per rotation; the reflector during Hessenberg
  L.total = L.rot + sizeof(cx<T>) * 32;
  return L;
}

template <typename T>
__global__ void __launch_bounds__(128, 6)
fp16s_kernel(FpParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int d = p.d;
  const Fp16sLayout<T> L = fp16s_layout<T>();
  const int half = (threadIdx.x >> 4) & 1;
  const int ln = threadIdx.x & 15;
  const int gi = threadIdx.x >> 4, gpc = blockDim.x >> 4;
  unsigned char* base = smem_raw + (size_t)gi * L.total;
  cx<T>* S = reinterpret_cast<cx<T>*>(base + L.S);
  cx<T>* rot = reinterpret_cast<cx<T>*>(base + L.rot);
  cx<T>* vv = rot;
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
    const cx<T>* __restrict__ Ag = reinterpret_cast<const cx<T>*>(p.A) + ia * tsz;
    const cx<T>* __restrict__ Bg = reinterpret_cast<const cx<T>*>(p.B) + ib * tsz;
    // ---- my column (jj, ll) of E, four rows (i, 0..3) at a time; tensors come through L1 (they are shared by
    // many problems of the batch).  Left fixed point: E^dagger, i.e. transposed reads and a conjugation.
    {
      const int jj = ln >> 2, ll = ln & 3;
      const int sa = p.left ? 1 : 4, oa = p.left ? jj * 4 : jj, ob = p.left ? ll * 4 : ll;
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        cx<T> acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = mk<T>(0, 0);
#pragma unroll 1
        for (int s = 0; s < d; ++s) {
          const cx<T> a = Ag[s * 16 + i * sa + oa];
#pragma unroll
          for (int k = 0; k < 4; ++k) cmad_c(acc[k], a, Bg[s * 16 + k * sa + ob]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (p.left) acc[k].im = -acc[k].im;
          S[F16S(i * 4 + k, ln)] = acc[k];
        }
      }
    }
    __syncwarp();
    // ---- Householder reduction to Hessenberg form, in place in the shared tile
#pragma unroll 1
    for (int k = 0; k + 2 < F16_N; ++k) {
      const cx<T> xk = S[F16S(ln, k)];                       // column k, my row
      const cx<T> alpha = S[F16S(k + 1, k)];
      T xn2 = (ln > k + 1) ? norm2(xk) : T(0);
#pragma unroll
      for (int m = 8; m >= 1; m >>= 1) xn2 += __shfl_xor_sync(0xffffffffu, xn2, m, 16);
      const bool skip = (xn2 == T(0)) && (alpha.im == T(0));
      T beta = sqrt(norm2(alpha) + xn2);
      if (alpha.re > T(0)) beta = -beta;
      cx<T> tau = mk<T>(0, 0), scal = mk<T>(0, 0);
      if (!skip) {
        const T ibeta = T(1) / beta;
        tau = mk<T>((beta - alpha.re) * ibeta, -alpha.im * ibeta);
        scal = cinv(alpha - mk<T>(beta, 0));
      }
      cx<T> vj = mk<T>(0, 0);                                // my component of the scaled reflector (v[k+1] = 1)
      if (ln == k + 1) vj = mk<T>(1, 0);
      else if (ln > k + 1) vj = xk * scal;
      __syncwarp();                                          // everybody has read column k
      vv[ln] = vj;
      if (!skip) {
        if (ln == k + 1) S[F16S(ln, k)] = mk<T>(beta, 0);
        else if (ln > k + 1) S[F16S(ln, k)] = mk<T>(0, 0);
      }
      __syncwarp();
      // left:  H <- (1 - conj(tau) v v^H) H   on my column (columns <= k are untouched: rows > k vanish there
      // except column k, which was just set)
      if (ln > k) {
        cx<T> w = mk<T>(0, 0);
        for (int i = k + 1; i < F16_N; ++i) cmad(w, conj(vv[i]), S[F16S(i, ln)]);
        w = w * conj(tau);
        for (int i = k + 1; i < F16_N; ++i) { cx<T> h = S[F16S(i, ln)]; cmsub(h, vv[i], w); S[F16S(i, ln)] = h; }
      }
      __syncwarp();
      // right: H <- H (1 - tau v v^H)   on my row
      {
        cx<T> u = mk<T>(0, 0);
        for (int j = k + 1; j < F16_N; ++j) cmad(u, S[F16S(ln, j)], vv[j]);
        u = u * tau;
        for (int j = k + 1; j < F16_N; ++j) { cx<T> h = S[F16S(ln, j)]; cmsub(h, u, conj(vv[j])); S[F16S(ln, j)] = h; }
      }
      __syncwarp();
    }

    // ---- shifted QR, all eigenvalues; keep the one of largest modulus
    int en = F16_N - 1, its = 0, fail = 0, sweeps = 0;
    T best2 = T(-1);
    cx<T> best = mk<T>(0, 0);
#pragma unroll 1
    for (;;) {
      // negligible sub-diagonal entries, all at once
      bool neg = false;
      if (ln >= 1) {
        T sc = cabs1(S[F16S(ln - 1, ln - 1)]) + cabs1(S[F16S(ln, ln)]);
        if (sc == T(0)) sc = T(1);
        neg = cabs1(S[F16S(ln, ln - 1)]) <= eps * sc;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, neg);
      const unsigned bits = (bal >> (16 * half)) & 0xffffu;
      int l = 0;
      while (en >= 0) {
        const unsigned m = bits & ((2u << en) - 1u) & ~1u;
        l = m ? (31 - __clz(m)) : 0;
        if (l == en || its >= maxit) {
          if (l != en) fail = 1;
          const cx<T> ev = S[F16S(en, en)];
          const T a2 = norm2(ev);
          if (a2 > best2) { best2 = a2; best = ev; }
          --en; its = 0;
        } else break;
      }
      if (__all_sync(0xffffffffu, en < 0)) break;
      // shift (Wilkinson; exceptional every 10 stalled sweeps) -- idle problem: empty window, sigma = 0
      cx<T> sigma = mk<T>(0, 0);
      const bool busy = en >= 1;
      if (busy) {
        const cx<T> a = S[F16S(en - 1, en - 1)], b = S[F16S(en - 1, en)];
        const cx<T> c = S[F16S(en, en - 1)], dd = S[F16S(en, en)];
        if (its == 10 || its == 20 || its == 30 || its == 40) {
          const T t = fabs(c.re) + (en >= 2 ? fabs(S[F16S(en - 1, en - 2)].re) : T(0));
          sigma = dd + mk<T>(t, 0);
        } else {
          sigma = dd;
          const cx<T> bc = b * c;
          if (bc.re != T(0) || bc.im != T(0)) {
            const cx<T> y = (a - dd) * T(0.5);
            cx<T> z = csqrt_nb(y * y + bc);
            if (y.re * z.re + y.im * z.im < T(0)) z = -z;
            sigma = dd - cdiv_nb(bc, y + z);
          }
        }
      }
      const int lw = busy ? l : F16_N - 1, enw = busy ? en : 0;       // my window (empty when idle)
      const int lo = __reduce_min_sync(0xffffffffu, lw);              // union over the warp: uniform loop bounds
      const int hi = __reduce_max_sync(0xffffffffu, enw);
      const bool in_win = busy && (ln >= lw) && (ln <= enw);
      __syncwarp();
      if (in_win) S[F16S(ln, ln)] = S[F16S(ln, ln)] - sigma;          // H - sigma on the window's diagonal
      if (busy && lw >= 1 && ln == lw - 1) S[F16S(lw, lw - 1)] = mk<T>(0, 0);   // the negligible entry becomes exact
      __syncwarp();
      // left phase: R = G_en ... G_{l+1} (H - sigma); lane = column, carries its entry of the upper row
      {
        cx<T> pu = S[F16S(lo, ln)];
        cx<T> qn = S[F16S(lo + 1, ln)];
#pragma unroll 1
        for (int i = lo + 1; i <= hi; ++i) {
          const cx<T> ql = qn;
          if (i < hi) qn = S[F16S(i + 1, ln)];
          const bool act = busy && (i > lw) && (i <= enw);
          // Givens parameters of rows (i-1, i): every lane runs the arithmetic on its own column, the owner of
          // column i-1 publishes its result through the rotation table (2 stores, 2 broadcast loads -- a
          // shuffle of four doubles costs ~24 instructions here)
          cx<T> c = mk<T>(1, 0), s = mk<T>(0, 0);
          T nr = T(0);
          {
            const T nr2 = norm2(pu) + norm2(ql);
            if (act && nr2 > T(0)) {
              const T inr = rsqrt_t<T>(nr2);
              c = pu * inr; s = ql * inr; nr = nr2 * inr;
            }
          }
          if (ln == i - 1) { rot[2 * i] = c; rot[2 * i + 1] = s; }
          __syncwarp();
          c = rot[2 * i]; s = rot[2 * i + 1];
          cx<T> top = conj(c) * pu; cmad(top, conj(s), ql);
          cx<T> bot = c * ql; cmsub(bot, s, pu);
          if (act && ln == i - 1) { top = mk<T>(nr, 0); bot = mk<T>(0, 0); }
          S[F16S(i - 1, ln)] = top;
          pu = bot;
        }
        S[F16S(hi, ln)] = pu;
      }
      __syncwarp();
      // right phase: H' = R G_{l+1}^H ... G_en^H + sigma; lane = row, carries its entry of the left column
      {
        cx<T> xl = S[F16S(ln, lo)];
#pragma unroll 1
        for (int j = lo + 1; j <= hi; ++j) {
          const cx<T> c = rot[2 * j], s = rot[2 * j + 1];
          const cx<T> yr = S[F16S(ln, j)];
          cx<T> a = xl * c; cmad(a, yr, s);
          cx<T> b = yr * conj(c); cmsub(b, xl, conj(s));
          S[F16S(ln, j - 1)] = a;
          xl = b;
        }
        S[F16S(ln, hi)] = xl;
      }
      if (in_win) S[F16S(ln, ln)] = S[F16S(ln, ln)] + sigma;          // my own row: no barrier needed
      __syncwarp();
      ++its;
      if (busy) ++sweeps;
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

// ---- quarter-warp variant: EIGHT lanes per problem, four problems per warp ------------------------------
// Lane q owns columns (rows) q and q + 8.  The per-step overhead that every lane executes whatever the
// number of problems in the warp (Givens generation, selects, loop and address arithmetic: ~130 of the ~180
// instructions of a step in the half-warp form, profiles/ncu_fp16_r02e.txt) is shared by four problems
// instead of two, and structural zeros are skipped: columns 0..7 have nothing below row 8, rows 8..15
// nothing left of column 7.
// One left-phase step of fp16s8_kernel<T, FASTRSQ, TRIM = true>: rotation i (rows (i-1, i)) on my columns q (LOW: while
// i <= 8) and q + 8.  LOW also says which half holds column i - 1, whose owner lane generates the rotation.
template <typename T, bool LOW>
__device__ __forceinline__ void fp16s8_left_step(cx<T>* S, cx<T>* rot, int q, int q8, int i, int hi, bool busy, int lw, int enw,
                                                 cx<T>& pa, cx<T>& pb, cx<T>& qa, cx<T>& qb) {
  const int in = i + 1 <= hi ? i + 1 : hi;
  cx<T> qa_n = mk<T>(0, 0);
  if (LOW && in <= 8) qa_n = S[F16S(in, q)];
  const cx<T> qb_n = S[F16S(in, q8)];
  const bool act = busy && (i > lw) && (i <= enw);
  const cx<T> f = LOW ? pa : pb, g = LOW ? qa : qb;
  const T nr2 = norm2(f) + norm2(g);
  const bool ok = act && nr2 > rsq_floor<T>::v();
  const T y = rsq_fast<T>(nr2);
  const T inr = ok ? y : T(0);                                 // identity rotation outside my window / for a zero pair
  cx<T> c, s;
  c.re = fma_t(f.re, inr, ok ? T(0) : T(1)); c.im = f.im * inr;
  s.re = g.re * inr; s.im = g.im * inr;
  const int src = (i - 1) & 7;
  const bool owner = q == src;
  if (owner) { rot[2 * i] = c; rot[2 * i + 1] = s; }
  c.re = __shfl_sync(0xffffffffu, c.re, src, 8); c.im = __shfl_sync(0xffffffffu, c.im, src, 8);
  s.re = __shfl_sync(0xffffffffu, s.re, src, 8); s.im = __shfl_sync(0xffffffffu, s.im, src, 8);
  cx<T> top = conj(c) * pb; cmad(top, conj(s), qb);
  cx<T> bot = c * qb; cmsub(bot, s, pb);
  if (LOW) {
    cx<T> ta = conj(c) * pa; cmad(ta, conj(s), qa);
    cx<T> ba = c * qa; cmsub(ba, s, pa);
    if (ok && owner) ba = mk<T>(0, 0);
    S[F16S(i - 1, q)] = ta;
    pa = ba;
  } else if (ok && owner) bot = mk<T>(0, 0);
  S[F16S(i - 1, q8)] = top;
  pb = bot;
  qa = qa_n; qb = qb_n;
}

template <typename T, bool FASTRSQ = false, bool TRIM = false>
__global__ void __launch_bounds__(64, 6)
fp16s8_kernel(FpParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int d = p.d;
  const Fp16sLayout<T> L = fp16s_layout<T>();
  const int grp = (threadIdx.x >> 3) & 3;
  const int q = threadIdx.x & 7, q8 = q + 8;
  const int gi = threadIdx.x >> 3, gpc = blockDim.x >> 3;
  unsigned char* base = smem_raw + (size_t)gi * L.total;
  cx<T>* S = reinterpret_cast<cx<T>*>(base + L.S);
  cx<T>* rot = reinterpret_cast<cx<T>*>(base + L.rot);
  cx<T>* vv = rot;
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
    const cx<T>* __restrict__ Ag = reinterpret_cast<const cx<T>*>(p.A) + ia * tsz;
    const cx<T>* __restrict__ Bg = reinterpret_cast<const cx<T>*>(p.B) + ib * tsz;
    // ---- columns q = (jj, ll) and q + 8 = (jj + 2, ll) of E (E^dagger for the left fixed point)
    {
      const int jj = q >> 2, ll = q & 3;
      const int sa = p.left ? 1 : 4;
      const int oa = p.left ? jj * 4 : jj, oa2 = p.left ? (jj + 2) * 4 : jj + 2, ob = p.left ? ll * 4 : ll;
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        cx<T> acc[4], acc2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc[k] = mk<T>(0, 0); acc2[k] = mk<T>(0, 0); }
#pragma unroll 1
        for (int s = 0; s < d; ++s) {
          const cx<T> a = Ag[s * 16 + i * sa + oa], a2 = Ag[s * 16 + i * sa + oa2];
#pragma unroll
          for (int k = 0; k < 4; ++k) { const cx<T> b = Bg[s * 16 + k * sa + ob]; cmad_c(acc[k], a, b); cmad_c(acc2[k], a2, b); }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (p.left) { acc[k].im = -acc[k].im; acc2[k].im = -acc2[k].im; }
          S[F16S(i * 4 + k, q)] = acc[k];
          S[F16S(i * 4 + k, q8)] = acc2[k];
        }
      }
    }
    __syncwarp();
    // ---- Householder reduction to Hessenberg form, in place in the shared tile
#pragma unroll 1
    for (int k = 0; k + 2 < F16_N; ++k) {
      const cx<T> xa = S[F16S(q, k)], xb = S[F16S(q8, k)];       // column k, my two rows
      const cx<T> alpha = S[F16S(k + 1, k)];
      T xn2 = ((q > k + 1) ? norm2(xa) : T(0)) + ((q8 > k + 1) ? norm2(xb) : T(0));
#pragma unroll
      for (int m = 4; m >= 1; m >>= 1) xn2 += __shfl_xor_sync(0xffffffffu, xn2, m, 8);
      const bool skip = (xn2 == T(0)) && (alpha.im == T(0));
      T beta = sqrt(norm2(alpha) + xn2);
      if (alpha.re > T(0)) beta = -beta;
      cx<T> tau = mk<T>(0, 0), scal = mk<T>(0, 0);
      if (!skip) {
        const T ibeta = T(1) / beta;
        tau = mk<T>((beta - alpha.re) * ibeta, -alpha.im * ibeta);
        scal = cinv(alpha - mk<T>(beta, 0));
      }
      cx<T> va = mk<T>(0, 0), vb = mk<T>(0, 0);
      if (q == k + 1) va = mk<T>(1, 0); else if (q > k + 1) va = xa * scal;
      if (q8 == k + 1) vb = mk<T>(1, 0); else if (q8 > k + 1) vb = xb * scal;
      __syncwarp();
      vv[q] = va; vv[q8] = vb;
      if (!skip) {
        if (q == k + 1) S[F16S(q, k)] = mk<T>(beta, 0); else if (q > k + 1) S[F16S(q, k)] = mk<T>(0, 0);
        if (q8 == k + 1) S[F16S(q8, k)] = mk<T>(beta, 0); else if (q8 > k + 1) S[F16S(q8, k)] = mk<T>(0, 0);
      }
      __syncwarp();
      // left on my columns q (only if q > k) and q + 8 (only if q + 8 > k)
      {
        cx<T> wa = mk<T>(0, 0), wb = mk<T>(0, 0);
        for (int i = k + 1; i < F16_N; ++i) { const cx<T> cv = conj(vv[i]); cmad(wa, cv, S[F16S(i, q)]); cmad(wb, cv, S[F16S(i, q8)]); }
        wa = (q > k) ? wa * conj(tau) : mk<T>(0, 0);
        wb = (q8 > k) ? wb * conj(tau) : mk<T>(0, 0);
        for (int i = k + 1; i < F16_N; ++i) {
          const cx<T> v = vv[i];
          cx<T> ha = S[F16S(i, q)], hb = S[F16S(i, q8)];
          cmsub(ha, v, wa); cmsub(hb, v, wb);
          S[F16S(i, q)] = ha; S[F16S(i, q8)] = hb;
        }
      }
      __syncwarp();
      // right on my rows q and q + 8
      {
        cx<T> ua = mk<T>(0, 0), ub = mk<T>(0, 0);
        for (int j = k + 1; j < F16_N; ++j) { const cx<T> v = vv[j]; cmad(ua, S[F16S(q, j)], v); cmad(ub, S[F16S(q8, j)], v); }
        ua = ua * tau; ub = ub * tau;
        for (int j = k + 1; j < F16_N; ++j) {
          const cx<T> cv = conj(vv[j]);
          cx<T> ha = S[F16S(q, j)], hb = S[F16S(q8, j)];
          cmsub(ha, ua, cv); cmsub(hb, ub, cv);
          S[F16S(q, j)] = ha; S[F16S(q8, j)] = hb;
        }
      }
      __syncwarp();
    }

    // ---- shifted QR, all eigenvalues; keep the one of largest modulus
    int en = F16_N - 1, its = 0, fail = 0, sweeps = 0;
    T best2 = T(-1);
    cx<T> best = mk<T>(0, 0);
#pragma unroll 1
    for (;;) {
      bool neg_a = false, neg_b;
      if (q >= 1) {
        T sc = cabs1(S[F16S(q - 1, q - 1)]) + cabs1(S[F16S(q, q)]);
        if (sc == T(0)) sc = T(1);
        neg_a = cabs1(S[F16S(q, q - 1)]) <= eps * sc;
      }
      {
        T sc = cabs1(S[F16S(q8 - 1, q8 - 1)]) + cabs1(S[F16S(q8, q8)]);
        if (sc == T(0)) sc = T(1);
        neg_b = cabs1(S[F16S(q8, q8 - 1)]) <= eps * sc;
      }
      const unsigned bal_a = __ballot_sync(0xffffffffu, neg_a), bal_b = __ballot_sync(0xffffffffu, neg_b);
      const unsigned bits = ((bal_a >> (8 * grp)) & 0xffu) | (((bal_b >> (8 * grp)) & 0xffu) << 8);
      int l = 0;
      while (en >= 0) {
        const unsigned m = bits & ((2u << en) - 1u) & ~1u;
        l = m ? (31 - __clz(m)) : 0;
        if (l == en || its >= maxit) {
          if (l != en) fail = 1;
          const cx<T> ev = S[F16S(en, en)];
          const T a2 = norm2(ev);
          if (a2 > best2) { best2 = a2; best = ev; }
          --en; its = 0;
        } else break;
      }
      if (__all_sync(0xffffffffu, en < 0)) break;
      cx<T> sigma = mk<T>(0, 0);
      const bool busy = en >= 1;
      if (busy) {
        const cx<T> a = S[F16S(en - 1, en - 1)], b = S[F16S(en - 1, en)];
        const cx<T> c = S[F16S(en, en - 1)], dd = S[F16S(en, en)];
        if (its == 10 || its == 20 || its == 30 || its == 40) {
          const T t = fabs(c.re) + (en >= 2 ? fabs(S[F16S(en - 1, en - 2)].re) : T(0));
          sigma = dd + mk<T>(t, 0);
        } else {
          sigma = dd;
          const cx<T> bc = b * c;
          if (bc.re != T(0) || bc.im != T(0)) {
            const cx<T> y = (a - dd) * T(0.5);
            cx<T> z = csqrt_nb(y * y + bc);
            if (y.re * z.re + y.im * z.im < T(0)) z = -z;
            sigma = dd - cdiv_nb(bc, y + z);
          }
        }
      }
      const int lw = busy ? l : F16_N - 1, enw = busy ? en : 0;
      const int lo = __reduce_min_sync(0xffffffffu, lw);
      const int hi = __reduce_max_sync(0xffffffffu, enw);
      const bool win_a = busy && (q >= lw) && (q <= enw), win_b = busy && (q8 >= lw) && (q8 <= enw);
      __syncwarp();
      if (win_a) S[F16S(q, q)] = S[F16S(q, q)] - sigma;
      if (win_b) S[F16S(q8, q8)] = S[F16S(q8, q8)] - sigma;
      if (busy && lw >= 1 && q == 0) S[F16S(lw, lw - 1)] = mk<T>(0, 0);
      __syncwarp();
      if constexpr (TRIM) {
        // the same two phases with fewer issued instructions (the kernel is issue-bound: ~66 % of the issue slots at cfg 3's
        // size): loops cut at row / column 8 so that the half of the owner lane is a compile-time fact, identity rotations
        // by a zeroed scale instead of eight selects, the exact zero of the annihilated entry kept but not the exact
        // modulus of its partner, bodies unrolled by two
        {
          cx<T> pa = S[F16S(lo < 8 ? lo : 8, q)], pb = S[F16S(lo, q8)];
          cx<T> qa = mk<T>(0, 0), qb = S[F16S(lo + 1 <= hi ? lo + 1 : hi, q8)];
          if (lo + 1 <= 8) qa = S[F16S(lo + 1 <= hi ? lo + 1 : hi, q)];
          int i = lo + 1;
          const int e1 = hi < 8 ? hi : 8;
#pragma unroll 2
          for (; i <= e1; ++i) fp16s8_left_step<T, true>(S, rot, q, q8, i, hi, busy, lw, enw, pa, pb, qa, qb);
#pragma unroll 2
          for (; i <= hi; ++i) fp16s8_left_step<T, false>(S, rot, q, q8, i, hi, busy, lw, enw, pa, pb, qa, qb);
          if (lo < 8) S[F16S(hi < 8 ? hi : 8, q)] = pa;
          S[F16S(hi, q8)] = pb;
        }
        __syncwarp();
        {
          cx<T> xa = S[F16S(q, lo)];
          const int j8 = lo + 1 > 8 ? lo + 1 : 8;             // first rotation that touches rows >= 8
          cx<T> xb = S[F16S(q8, j8 - 1)];
          int j = lo + 1;
          const int e1 = hi < 7 ? hi : 7;
#pragma unroll 2
          for (; j <= e1; ++j) {
            const cx<T> c = rot[2 * j], s = rot[2 * j + 1];
            const cx<T> ya = S[F16S(q, j)];
            cx<T> a = xa * c; cmad(a, ya, s);
            cx<T> b = ya * conj(c); cmsub(b, xa, conj(s));
            S[F16S(q, j - 1)] = a;
            xa = b;
          }
#pragma unroll 2
          for (; j <= hi; ++j) {
            const cx<T> c = rot[2 * j], s = rot[2 * j + 1];
            const cx<T> ya = S[F16S(q, j)], yb = S[F16S(q8, j)];
            cx<T> a = xa * c; cmad(a, ya, s);
            cx<T> b = ya * conj(c); cmsub(b, xa, conj(s));
            cx<T> a2 = xb * c; cmad(a2, yb, s);
            cx<T> b2 = yb * conj(c); cmsub(b2, xb, conj(s));
            S[F16S(q, j - 1)] = a;
            S[F16S(q8, j - 1)] = a2;
            xa = b; xb = b2;
          }
          S[F16S(q, hi)] = xa;
          if (hi >= 8) S[F16S(q8, hi)] = xb;
        }
      } else {
      // left phase.  Column q is carried in pa (rows <= 8 only), column q + 8 in pb.
      {
        cx<T> pa = S[F16S(lo < 8 ? lo : 8, q)], pb = S[F16S(lo, q8)];
        // next row's entries of my two columns, loaded one rotation ahead (row i is not written before rotation i + 1)
        cx<T> qa = mk<T>(0, 0), qb = S[F16S(lo + 1 <= hi ? lo + 1 : hi, q8)];
        if (lo + 1 <= 8) qa = S[F16S(lo + 1 <= hi ? lo + 1 : hi, q)];
#pragma unroll 1
        for (int i = lo + 1; i <= hi; ++i) {
          const bool low = i <= 8;                           // uniform: column q still has entries in rows (i-1, i)
          const int in = i + 1 <= hi ? i + 1 : hi;
          cx<T> qa_n = mk<T>(0, 0);
          if (in <= 8) qa_n = S[F16S(in, q)];
          const cx<T> qb_n = S[F16S(in, q8)];
          const bool act = busy && (i > lw) && (i <= enw);
          const cx<T> f = low ? pa : pb, g = low ? qa : qb;  // column i-1 is an "a" column iff i-1 < 8
          const T nr2 = norm2(f) + norm2(g);
          const bool ok = act && nr2 > (FASTRSQ ? rsq_floor<T>::v() : T(0));
          const T inr = FASTRSQ ? rsq_fast<T>(nr2) : rsqrt_t<T>(nr2);   // FASTRSQ: no slow-path branch inside the sweep body
          cx<T> c = f * inr, s = g * inr;
          const T nr = nr2 * inr;
          const bool owner = q == ((i - 1) & 7);
          if (!ok) { c = mk<T>(1, 0); s = mk<T>(0, 0); }
          if (owner) { rot[2 * i] = c; rot[2 * i + 1] = s; }   // for the right phase (read after the __syncwarp below the loop)
          // the rotation from its owner lane by shuffles: no shared-memory round trip on the critical path
          const int src = (i - 1) & 7;
          c.re = __shfl_sync(0xffffffffu, c.re, src, 8); c.im = __shfl_sync(0xffffffffu, c.im, src, 8);
          s.re = __shfl_sync(0xffffffffu, s.re, src, 8); s.im = __shfl_sync(0xffffffffu, s.im, src, 8);
          cx<T> top = conj(c) * pb; cmad(top, conj(s), qb);
          cx<T> bot = c * qb; cmsub(bot, s, pb);
          if (low) {
            cx<T> ta = conj(c) * pa; cmad(ta, conj(s), qa);
            cx<T> ba = c * qa; cmsub(ba, s, pa);
            if (ok && owner) { ta = mk<T>(nr, 0); ba = mk<T>(0, 0); }
            S[F16S(i - 1, q)] = ta;
            pa = ba;
          } else if (ok && owner) { top = mk<T>(nr, 0); bot = mk<T>(0, 0); }
          S[F16S(i - 1, q8)] = top;
          pb = bot;
          qa = qa_n; qb = qb_n;
        }
        if (lo < 8) S[F16S(hi < 8 ? hi : 8, q)] = pa;
        S[F16S(hi, q8)] = pb;
      }
      __syncwarp();
      // right phase.  Row q is carried in xa, row q + 8 (columns >= 7 only) in xb.
      {
        cx<T> xa = S[F16S(q, lo)];
        const int j8 = lo + 1 > 8 ? lo + 1 : 8;             // first rotation that touches rows >= 8
        cx<T> xb = S[F16S(q8, j8 - 1)];
#pragma unroll 1
        for (int j = lo + 1; j <= hi; ++j) {
          const cx<T> c = rot[2 * j], s = rot[2 * j + 1];
          const cx<T> ya = S[F16S(q, j)];
          cx<T> a = xa * c; cmad(a, ya, s);
          cx<T> b = ya * conj(c); cmsub(b, xa, conj(s));
          S[F16S(q, j - 1)] = a;
          xa = b;
          if (j >= 8) {
            const cx<T> yb = S[F16S(q8, j)];
            cx<T> a2 = xb * c; cmad(a2, yb, s);
            cx<T> b2 = yb * conj(c); cmsub(b2, xb, conj(s));
            S[F16S(q8, j - 1)] = a2;
            xb = b2;
          }
        }
        S[F16S(q, hi)] = xa;
        if (hi >= 8) S[F16S(q8, hi)] = xb;
      }
      }
      if (win_a) S[F16S(q, q)] = S[F16S(q, q)] + sigma;
      if (win_b) S[F16S(q8, q8)] = S[F16S(q8, q8)] + sigma;
      __syncwarp();
      ++its;
      if (busy) ++sweeps;
    }
    if (live && q == 0) {
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
