// qmps_b200 leading eigenvalue of the 16 x 16 mixed transfer matrix (D = 4), PACKED two-kernel form of fp16s8_kernel
// (kernels_fp16s.cuh; same algorithm and outputs: Householder -> Hessenberg, shifted complex QR for all eigenvalues,
// arg-max |lambda|; the cost -sqrt|eta| of qmps/loschmidts/time_evo.py:75-116, BASELINE config 3).
//
// fp16s8_kernel is bound by the latency of the serial Givens chain with 12 warps per SM, limited by the 4.5 KB of
// shared memory per problem (profiles/ncu_fp16s_r02f.txt).  The full 16 x 16 tile is needed by the Householder reduction
// only; the QR phase works on an upper Hessenberg matrix, 151 of 256 entries.  As for D = 8 (kernels_fp64p.cuh):
//   MODE 1  builds E, reduces it in the swizzled tile and writes the Hessenberg matrix packed to a global workspace
//           (2.4 KB per problem, complex128);
//   MODE 2  loads the packed matrix into 2.4 KB + 0.5 KB of shared memory -- 72 problems (18 warps) per SM instead of 48 --
//           and runs the sweeps of fp16s8_kernel on it;
//   MODE 0  is fp16s8_kernel itself (one launch, unpacked), kept for reference.
// Packed layout: 8 lines of 19 entries; line q holds row q (columns q-1..15, from the front) and row 15-q (columns
// 14-q..15, at the back); both halves linear:  row r < 8: 18 r + 1 + j,  row r >= 8: 288 - 19 r + j.  Entries below the
// sub-diagonal do not exist: loads of them are predicated to zero, stores predicated off.
#pragma once
#include <cuda_runtime.h>
#include "kernels_fp16s.cuh"

namespace qmps {

constexpr int F16P_SIZE = 8 * 19;
__device__ __forceinline__ int f16p_row(int r) { return r < 8 ? 18 * r + 1 : 288 - 19 * r; }
template <typename T> QMPS_HD Fp16sLayout<T> fp16p_layout() {
  Fp16sLayout<T> L;
  L.S = 0;
  L.rot = sizeof(cx<T>) * F16P_SIZE;
  L.total = L.rot + sizeof(cx<T>) * 32;
  return L;
}
#define X16(i, j) (PACKED ? f16p_row(i) + (j) : F16S(i, j))
#define PK16(cond) (!PACKED || (cond))
#define LD16(cond, idx) (PK16(cond) ? S[idx] : mk<T>(0, 0))

template <typename T, int MODE>
__global__ void __launch_bounds__(64, MODE == 2 ? 9 : 6)
fp16p8_kernel(FpParams p) {
  constexpr bool PACKED = MODE == 2;
  constexpr bool FASTRSQ = true;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int d = p.d;
  const Fp16sLayout<T> L = PACKED ? fp16p_layout<T>() : fp16s_layout<T>();
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
  for (int64_t k0 = (int64_t)blockIdx.x * gpc; k0 < p.n_chunk; k0 += stride) {
    int64_t slot = k0 + gi;                                   // workspace slot = index inside this launch's slice of the batch
    const bool live = slot < p.n_chunk;
    if (!live) slot = p.n_chunk - 1;
    const int64_t pid = p.pid_offset + slot;
    cx<T>* __restrict__ W = reinterpret_cast<cx<T>*>(p.ws) + (size_t)slot * F16P_SIZE;
    if constexpr (MODE == 2) {
#pragma unroll 1
      for (int e = q; e < F16P_SIZE; e += 8) S[e] = W[e];
      __syncwarp();
    } else {
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

    if constexpr (MODE == 1) {
      // ---- pack: row r, columns r-1..15 (the entries below the sub-diagonal are exact zeros and are not stored)
#pragma unroll 1
      for (int r = 0; r < F16_N; ++r) {
        if (q >= r - 1) W[f16p_row(r) + q] = S[F16S(r, q)];
        if (q8 >= r - 1) W[f16p_row(r) + q8] = S[F16S(r, q8)];
      }
      if (q == 0) W[0] = mk<T>(0, 0);
      __syncwarp();
      continue;
    }
    }  // MODE != 2
    // ---- shifted QR, all eigenvalues; keep the one of largest modulus
    int en = F16_N - 1, its = 0, fail = 0, sweeps = 0;
    T best2 = T(-1);
    cx<T> best = mk<T>(0, 0);
#pragma unroll 1
    for (;;) {
      bool neg_a = false, neg_b;
      if (q >= 1) {
        T sc = cabs1(S[X16(q - 1, q - 1)]) + cabs1(S[X16(q, q)]);
        if (sc == T(0)) sc = T(1);
        neg_a = cabs1(S[X16(q, q - 1)]) <= eps * sc;
      }
      {
        T sc = cabs1(S[X16(q8 - 1, q8 - 1)]) + cabs1(S[X16(q8, q8)]);
        if (sc == T(0)) sc = T(1);
        neg_b = cabs1(S[X16(q8, q8 - 1)]) <= eps * sc;
      }
      const unsigned bal_a = __ballot_sync(0xffffffffu, neg_a), bal_b = __ballot_sync(0xffffffffu, neg_b);
      const unsigned bits = ((bal_a >> (8 * grp)) & 0xffu) | (((bal_b >> (8 * grp)) & 0xffu) << 8);
      int l = 0;
      while (en >= 0) {
        const unsigned m = bits & ((2u << en) - 1u) & ~1u;
        l = m ? (31 - __clz(m)) : 0;
        if (l == en || its >= maxit) {
          if (l != en) fail = 1;
          const cx<T> ev = S[X16(en, en)];
          const T a2 = norm2(ev);
          if (a2 > best2) { best2 = a2; best = ev; }
          --en; its = 0;
        } else break;
      }
      if (__all_sync(0xffffffffu, en < 0)) break;
      cx<T> sigma = mk<T>(0, 0);
      const bool busy = en >= 1;
      if (busy) {
        const cx<T> a = S[X16(en - 1, en - 1)], b = S[X16(en - 1, en)];
        const cx<T> c = S[X16(en, en - 1)], dd = S[X16(en, en)];
        if (its == 10 || its == 20 || its == 30 || its == 40) {
          const T t = fabs(c.re) + (en >= 2 ? fabs(S[X16(en - 1, en - 2)].re) : T(0));
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
      if (win_a) S[X16(q, q)] = S[X16(q, q)] - sigma;
      if (win_b) S[X16(q8, q8)] = S[X16(q8, q8)] - sigma;
      if (busy && lw >= 1 && q == 0) S[X16(lw, lw - 1)] = mk<T>(0, 0);
      __syncwarp();
      // left phase.  Column q is carried in pa (rows <= 8 only), column q + 8 in pb.
      {
        const int r0a = lo < 8 ? lo : 8;
        cx<T> pa = LD16(q >= r0a - 1, X16(r0a, q)), pb = LD16(q8 >= lo - 1, X16(lo, q8));
        // next row's entries of my two columns, loaded one rotation ahead (row i is not written before rotation i + 1)
        const int r1 = lo + 1 <= hi ? lo + 1 : hi;
        cx<T> qa = mk<T>(0, 0), qb = LD16(q8 >= r1 - 1, X16(r1, q8));
        if (lo + 1 <= 8) qa = LD16(q >= r1 - 1, X16(r1, q));
#pragma unroll 1
        for (int i = lo + 1; i <= hi; ++i) {
          const bool low = i <= 8;                           // uniform: column q still has entries in rows (i-1, i)
          const int in = i + 1 <= hi ? i + 1 : hi;
          cx<T> qa_n = mk<T>(0, 0);
          if (in <= 8) qa_n = LD16(q >= in - 1, X16(in, q));
          const cx<T> qb_n = LD16(q8 >= in - 1, X16(in, q8));
          const bool act = busy && (i > lw) && (i <= enw);
          const cx<T> f = low ? pa : pb, g = low ? qa : qb;  // column i-1 is an "a" column iff i-1 < 8
          const T nr2 = norm2(f) + norm2(g);
          const bool ok = act && nr2 > rsq_floor<T>::v();
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
            if (PK16(q >= i - 2)) S[X16(i - 1, q)] = ta;
            pa = ba;
          } else if (ok && owner) { top = mk<T>(nr, 0); bot = mk<T>(0, 0); }
          if (PK16(q8 >= i - 2)) S[X16(i - 1, q8)] = top;
          pb = bot;
          qa = qa_n; qb = qb_n;
        }
        { const int rf = hi < 8 ? hi : 8; if (lo < 8 && PK16(q >= rf - 1)) S[X16(rf, q)] = pa; }
        if (PK16(q8 >= hi - 1)) S[X16(hi, q8)] = pb;
      }
      __syncwarp();
      // right phase.  Row q is carried in xa, row q + 8 (columns >= 7 only) in xb.
      {
        cx<T> xa = LD16(q <= lo + 1, X16(q, lo));
        const int j8 = lo + 1 > 8 ? lo + 1 : 8;             // first rotation that touches rows >= 8
        cx<T> xb = LD16(q8 <= j8, X16(q8, j8 - 1));
#pragma unroll 1
        for (int j = lo + 1; j <= hi; ++j) {
          const cx<T> c = rot[2 * j], s = rot[2 * j + 1];
          const cx<T> ya = LD16(q <= j + 1, X16(q, j));
          cx<T> a = xa * c; cmad(a, ya, s);
          cx<T> b = ya * conj(c); cmsub(b, xa, conj(s));
          if (PK16(q <= j)) S[X16(q, j - 1)] = a;
          xa = b;
          if (j >= 8) {
            const cx<T> yb = LD16(q8 <= j + 1, X16(q8, j));
            cx<T> a2 = xb * c; cmad(a2, yb, s);
            cx<T> b2 = yb * conj(c); cmsub(b2, xb, conj(s));
            if (PK16(q8 <= j)) S[X16(q8, j - 1)] = a2;
            xb = b2;
          }
        }
        if (PK16(q <= hi)) S[X16(q, hi)] = xa;
        if (hi >= 8 && PK16(q8 <= hi)) S[X16(q8, hi)] = xb;
      }
      if (win_a) S[X16(q, q)] = S[X16(q, q)] + sigma;
      if (win_b) S[X16(q8, q8)] = S[X16(q8, q8)] + sigma;
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

#undef X16
#undef PK16
#undef LD16

}  // namespace qmps
