// qmps_b200 leading eigenvalue of the 16 x 16 mixed transfer matrix (D = 4, any d), quarter-warp form.
// Same algorithm and the same outputs as kernels_fp16.cuh (Householder -> Hessenberg, shifted complex QR
// for all eigenvalues, arg-max |lambda|; qmps/loschmidts/time_evo.py:75-116, qmps/time_evolve_tools.py:84-91,
// xmps Map(A,B).right/left_fixed_point eigenvalue), laid out so that the FP64 pipe does less redundant work:
//   * EIGHT lanes per problem, four problems per warp; lane q owns columns q and q + 8 (two register
//     arrays of 16 complex) in the left phase and rows q and q + 8 in the right phase.  The Givens
//     generation (norm, rsqrt, scaling -- a serial chain that every lane executes) is paid once per FOUR
//     problems instead of once per two, and each step carries two independent column (row) updates that
//     fill the latency of that chain;
//   * structural zeros are skipped statically: columns 0..7 have nothing in rows >= 9, rows 8..15 nothing
//     in columns <= 6;
//   * everything addressed by a lane-dependent index (diagonal shift, exact zero at H[l][l-1]) is done on
//     the shared row-major copy; window bounds come from redux.sync so the skip branches are uniform.
#pragma once
#include <cuda_runtime.h>
#include "kernels_fp16.cuh"

namespace qmps {

template <typename T> __device__ __forceinline__ cx<T> shfl8(cx<T> v, int src) {
  cx<T> r;
  r.re = __shfl_sync(0xffffffffu, v.re, src, 8);
  r.im = __shfl_sync(0xffffffffu, v.im, src, 8);
  return r;
}

// rotation of the pair (p, q) by the left Givens (c, s): (conj(c) p + conj(s) q, c q - s p)
template <typename T> __device__ __forceinline__ void rot_left(cx<T>& p, cx<T>& q, cx<T> c, cx<T> s) {
  cx<T> top = conj(c) * p; cmad(top, conj(s), q);
  cx<T> bot = c * q; cmsub(bot, s, p);
  p = top; q = bot;
}
// columns (x, y) times the adjoint rotation: (x c + y s, y conj(c) - x conj(s))
template <typename T> __device__ __forceinline__ void rot_right(cx<T>& x, cx<T>& y, cx<T> c, cx<T> s) {
  cx<T> a = x * c; cmad(a, y, s);
  cx<T> b = y * conj(c); cmsub(b, x, conj(s));
  x = a; y = b;
}

template <typename T>
__device__ __forceinline__ void fp16x8_sweep(int q, int l, int en, int lo, int hi, cx<T> sigma, cx<T>* S,
                                             cx<T>* rot) {
  const int q8 = q + 8;
  const bool win_a = (q >= l) && (q <= en), win_b = (q8 >= l) && (q8 <= en);
  if (win_a) S[q * F16_LD + q] = S[q * F16_LD + q] - sigma;
  if (win_b) S[q8 * F16_LD + q8] = S[q8 * F16_LD + q8] - sigma;
  if (l >= 1 && q == 0) S[l * F16_LD + (l - 1)] = mk<T>(0, 0);
  __syncwarp();
  cx<T> a[F16_N], b[F16_N];                                    // columns q and q + 8
#pragma unroll
  for (int i = 0; i < F16_N; ++i) {
    if (i <= hi) {
      if (i <= 8) a[i] = S[i * F16_LD + q];                    // column q <= 7 is zero below row 8
      b[i] = S[i * F16_LD + q8];
    }
  }
  __syncwarp();
  // left phase
#pragma unroll
  for (int i = 1; i < F16_N; ++i) {
    if (i > lo && i <= hi) {
      const bool act = (i > l) && (i <= en);
      const int owner = (i - 1) & 7;
      const cx<T> f = shfl8((i - 1) < 8 ? a[i - 1] : b[i - 1], owner);
      const cx<T> g = shfl8((i - 1) < 8 ? a[i] : b[i], owner);
      const T nr2 = norm2(f) + norm2(g);
      cx<T> c = mk<T>(1, 0), s = mk<T>(0, 0);
      T nr = T(0);
      if (act && nr2 > T(0)) {
        const T inr = rsqrt_t<T>(nr2);
        c = f * inr; s = g * inr; nr = nr2 * inr;
      }
      if (q == 0) { rot[2 * i] = c; rot[2 * i + 1] = s; }
      if (i <= 8) rot_left(a[i - 1], a[i], c, s);             // rows i-1, i of columns <= 7 vanish for i >= 9
      rot_left(b[i - 1], b[i], c, s);                          // (H[8][7] was annihilated by step 8)
      if (act && q == owner) {
        if ((i - 1) < 8) { a[i - 1] = mk<T>(nr, 0); if (i <= 8) a[i] = mk<T>(0, 0); }
        else { b[i - 1] = mk<T>(nr, 0); b[i] = mk<T>(0, 0); }
      }
    }
  }
  // columns -> rows
#pragma unroll
  for (int i = 0; i < F16_N; ++i) {
    if (i <= hi) {
      S[i * F16_LD + q] = (i <= 8) ? a[i] : mk<T>(0, 0);
      S[i * F16_LD + q8] = b[i];
    }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < F16_N; ++j) {
    if (j <= hi) {
      a[j] = S[q * F16_LD + j];                                // row q
      if (j >= 7) b[j] = S[q8 * F16_LD + j];                   // row q + 8 >= 8 is zero left of column 7
    }
  }
  // right phase
#pragma unroll
  for (int j = 1; j < F16_N; ++j) {
    if (j > lo && j <= hi) {
      const cx<T> c = rot[2 * j], s = rot[2 * j + 1];
      rot_right(a[j - 1], a[j], c, s);
      if (j >= 8) rot_right(b[j - 1], b[j], c, s);
    }
  }
  // rows back to the shared copy, + sigma on the diagonal (same thread owns its rows)
#pragma unroll
  for (int j = 0; j < F16_N; ++j) {
    if (j <= hi) {
      S[q * F16_LD + j] = a[j];
      S[q8 * F16_LD + j] = (j >= 7) ? b[j] : mk<T>(0, 0);
    }
  }
  if (win_a) S[q * F16_LD + q] = S[q * F16_LD + q] + sigma;
  if (win_b) S[q8 * F16_LD + q8] = S[q8 * F16_LD + q8] + sigma;
}

template <typename T, int MINB>
__global__ void __launch_bounds__(64, MINB)
fp16x8_kernel(FpParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int d = p.d;
  const Fp16Layout<T> L = fp16_layout<T>(d);
  const int grp = (threadIdx.x >> 3) & 3;                 // problem slot within the warp
  const int q = threadIdx.x & 7, q8 = q + 8;
  const int gi = threadIdx.x >> 3, gpc = blockDim.x >> 3;
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
    for (int e = q; e < d * F16_N; e += 8) { As[e] = Ag[e]; Bs[e] = Bg[e]; }
    __syncwarp();
    // ---- columns q and q + 8 of E (of E^dagger for the left fixed point: transposed reads, conjugated)
    cx<T> ha[F16_N], hb[F16_N];
#pragma unroll
    for (int r = 0; r < F16_N; ++r) { ha[r] = mk<T>(0, 0); hb[r] = mk<T>(0, 0); }
    {
      const int jj = q >> 2, ll = q & 3;                     // column q = (jj, ll), column q + 8 = (jj + 2, ll)
      const int sa = p.left ? 1 : 4;
      const int oa = p.left ? jj * 4 : jj, oa2 = p.left ? (jj + 2) * 4 : jj + 2, ob = p.left ? ll * 4 : ll;
#pragma unroll 1
      for (int s = 0; s < d; ++s) {
        cx<T> xa[4], xa2[4], xb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          xa[i] = As[s * 16 + i * sa + oa]; xa2[i] = As[s * 16 + i * sa + oa2]; xb[i] = Bs[s * 16 + i * sa + ob];
        }
#pragma unroll
        for (int r = 0; r < F16_N; ++r) { cmad_c(ha[r], xa[r >> 2], xb[r & 3]); cmad_c(hb[r], xa2[r >> 2], xb[r & 3]); }
      }
      if (p.left) {
#pragma unroll
        for (int r = 0; r < F16_N; ++r) { ha[r].im = -ha[r].im; hb[r].im = -hb[r].im; }
      }
    }
    __syncwarp();                                          // A / B staging may share the tile S
    // ---- Householder reduction to Hessenberg form (a real loop: the reflector is zero above column k)
#pragma unroll 1
    for (int k = 0; k + 2 < F16_N; ++k) {
      if (k < 8) {
        if (q == k) {
#pragma unroll
          for (int i = 0; i < F16_N; ++i) vbuf[i] = ha[i];
        }
      } else if (q8 == k) {
#pragma unroll
        for (int i = 0; i < F16_N; ++i) vbuf[i] = hb[i];
      }
      __syncwarp();
      const cx<T> alpha = vbuf[k + 1];
      T xn2 = ((q > k + 1) ? norm2(vbuf[q]) : T(0)) + ((q8 > k + 1) ? norm2(vbuf[q8]) : T(0));
#pragma unroll
      for (int m = 4; m >= 1; m >>= 1) xn2 += __shfl_xor_sync(0xffffffffu, xn2, m, 8);
      const bool skip = (xn2 == T(0)) && (alpha.im == T(0));
      T beta = sqrt(norm2(alpha) + xn2);
      if (alpha.re > T(0)) beta = -beta;
      cx<T> tau = mk<T>(0, 0), scal = mk<T>(0, 0);
      if (!skip) {
        const T ib = T(1) / beta;
        tau = mk<T>((beta - alpha.re) * ib, -alpha.im * ib);
        scal = cinv(alpha - mk<T>(beta, 0));
      }
      cx<T> va = mk<T>(0, 0), vb = mk<T>(0, 0);               // my components of the scaled reflector
      if (q == k + 1) va = mk<T>(1, 0); else if (q > k + 1) va = vbuf[q] * scal;
      if (q8 == k + 1) vb = mk<T>(1, 0); else if (q8 > k + 1) vb = vbuf[q8] * scal;
      vbuf[16 + q] = va; vbuf[16 + q8] = vb;
      __syncwarp();
      // left:  H <- (1 - conj(tau) v v^H) H   on my two columns
      cx<T> wa = mk<T>(0, 0), wb = mk<T>(0, 0);
#pragma unroll
      for (int i = 1; i < F16_N; ++i) { const cx<T> cv = conj(vbuf[16 + i]); cmad(wa, cv, ha[i]); cmad(wb, cv, hb[i]); }
      wa = wa * conj(tau); wb = wb * conj(tau);
#pragma unroll
      for (int i = 1; i < F16_N; ++i) { const cx<T> v = vbuf[16 + i]; cmsub(ha[i], v, wa); cmsub(hb[i], v, wb); }
      {
        const bool fa = !skip && q == k, fb = !skip && q8 == k;
#pragma unroll
        for (int i = 1; i < F16_N; ++i) {
          const bool sub = (i == k + 1), below = (i > k + 1);
          ha[i].re = (fa && sub) ? beta : ((fa && below) ? T(0) : ha[i].re);
          ha[i].im = (fa && (sub || below)) ? T(0) : ha[i].im;
          hb[i].re = (fb && sub) ? beta : ((fb && below) ? T(0) : hb[i].re);
          hb[i].im = (fb && (sub || below)) ? T(0) : hb[i].im;
        }
      }
      // right: u = H v: per-lane partial sums over my two columns, then row sums over the eight lanes
#pragma unroll
      for (int i = 0; i < F16_N; ++i) { cx<T> t = ha[i] * va; cmad(t, hb[i], vb); S[i * F16_LD + q] = t; }
      __syncwarp();
      cx<T> ua = mk<T>(0, 0), ub = mk<T>(0, 0);
#pragma unroll
      for (int j = 0; j < 8; ++j) { ua = ua + S[q * F16_LD + j]; ub = ub + S[q8 * F16_LD + j]; }
      ubuf[q] = ua * tau; ubuf[q8] = ub * tau;
      __syncwarp();
      const cx<T> cva = conj(va), cvb = conj(vb);
#pragma unroll
      for (int i = 0; i < F16_N; ++i) { const cx<T> u = ubuf[i]; cmsub(ha[i], u, cva); cmsub(hb[i], u, cvb); }
      __syncwarp();
    }
    // row-major copy: what the deflation test, the shifts and the sweeps work from
#pragma unroll
    for (int i = 0; i < F16_N; ++i) { S[i * F16_LD + q] = ha[i]; S[i * F16_LD + q8] = hb[i]; }
    __syncwarp();

    // ---- shifted QR, all eigenvalues; keep the one of largest modulus
    int en = F16_N - 1, its = 0, fail = 0, sweeps = 0;
    T best2 = T(-1);
    cx<T> best = mk<T>(0, 0);
#pragma unroll 1
    for (;;) {
      // negligible sub-diagonal entries, all at once: lane q tests rows q and q + 8
      bool neg_a = false, neg_b;
      if (q >= 1) {
        T sc = cabs1(S[(q - 1) * F16_LD + (q - 1)]) + cabs1(S[q * F16_LD + q]);
        if (sc == T(0)) sc = T(1);
        neg_a = cabs1(S[q * F16_LD + (q - 1)]) <= eps * sc;
      }
      {
        T sc = cabs1(S[(q8 - 1) * F16_LD + (q8 - 1)]) + cabs1(S[q8 * F16_LD + q8]);
        if (sc == T(0)) sc = T(1);
        neg_b = cabs1(S[q8 * F16_LD + (q8 - 1)]) <= eps * sc;
      }
      const unsigned bal_a = __ballot_sync(0xffffffffu, neg_a), bal_b = __ballot_sync(0xffffffffu, neg_b);
      const unsigned bits = ((bal_a >> (8 * grp)) & 0xffu) | (((bal_b >> (8 * grp)) & 0xffu) << 8);
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
      const int lo_mine = (en >= 1) ? lw : F16_N - 1, hi_mine = (en >= 1) ? enw : 0;
      const int lo = __reduce_min_sync(0xffffffffu, lo_mine);
      const int hi = __reduce_max_sync(0xffffffffu, hi_mine);
      __syncwarp();
      fp16x8_sweep<T>(q, lw, enw, lo, hi, sigma, S, rot);
      __syncwarp();
      ++its;
      if (en >= 1) ++sweeps;
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
