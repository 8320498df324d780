// qmps_b200 leading eigenvalue of the 64 x 64 mixed transfer matrix (D = 8), PACKED two-kernel form of
// kernels_fp64w.cuh (same algorithm, same outputs: Householder -> Hessenberg, shifted complex QR for all eigenvalues,
// arg-max |lambda|; xmps' Map(A,B).right/left_fixed_point eigenvalue behind qmps/loschmidts/time_evo.py:75-116 and
// qmps/time_evolve_tools.py:84-91).
//
// Why two kernels.  ncu of the one-kernel form (profiles/ncu_fp64w_r02l.txt): one warp per SM sub-partition (65 KB of
// shared memory per complex128 problem: three per SM, a fourth scheduler idle), 3.45 cycles per issued instruction of
// which 1.6 are fixed-latency dependency stalls and 0.7 shared-memory round trips -- the QR sweep is a dependent chain
// and only more warps hide it.  The 65 KB are needed by the Householder reduction alone: the QR phase works on an
// upper Hessenberg matrix, 2143 of 4096 entries.  So:
//   fp64p_hess_kernel   builds E and reduces it (full padded tile, a CTA of 64 threads per problem, three per SM;
//                       throughput work) and writes the Hessenberg matrix PACKED to a global workspace (34 KB each);
//   fp64p_qr_kernel     loads the packed matrix (linear, coalesced) into 34 KB of shared memory -- SIX warps per SM
//                       in complex128, twelve in complex64 -- and runs the sweeps of kernels_fp64w.cuh on it.
// Packed layout: 32 lines of 67 entries; line q holds row q (columns q-1..63, 65-q entries, from the front) and row
// 63-q (columns 62-q..63, q+2 entries, at the back).  Both halves are LINEAR in (row, column):
//   row r < 32:  66 r + 1 + j          row r >= 32:  4224 - 67 r + j
// so the sweep loops still advance pointers by constants; "lane = column" accesses are consecutive, "lane = row"
// accesses have stride 66 (two-way conflict, two accesses per rotation) or 67 (conflict-free).  Entries below the
// sub-diagonal do not exist: loads of them are predicated to zero, stores predicated off.
#pragma once
#include <cuda_runtime.h>
#include "kernels_fp64w.cuh"

namespace qmps {

constexpr int F64P_SIZE = 32 * 67;
__device__ __forceinline__ int f64p_row(int r) { return r < 32 ? 66 * r + 1 : 4224 - 67 * r; }
#define F64P(r, j) (f64p_row(r) + (j))

template <typename T> struct Fp64pLayout { size_t S, rot, total; };
template <typename T> QMPS_HD Fp64pLayout<T> fp64p_layout() {
  Fp64pLayout<T> L;
  L.S = 0;
  L.rot = sizeof(cx<T>) * F64P_SIZE;
  L.total = L.rot + sizeof(cx<T>) * (2 * F64_N + 2);
  return L;
}

// ---- kernel 1: E -> Hessenberg -> packed global workspace --------------------------------------------------------
// problems [p.pid_offset, p.pid_offset + p.n_chunk) of the batch; workspace slot = pid - p.pid_offset.
// The reduction is throughput work (independent columns in the left update, independent rows in the right one), and with
// one warp per problem it ran at a dependent-FMA latency per row (33 of the 59 ms of the first packed version): here a
// CTA of 128 threads owns the problem: thread = (column t (left) / row t (right), half of the summation range), partial
// dot products combined through shared memory, two accumulators each.
template <typename T>
__global__ void __launch_bounds__(128)
fp64p_hess_kernel(FpParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Fp64wLayout<T> L = fp64w_layout<T>();
  const int tid = threadIdx.x, t = tid & 63, hf = tid >> 6;     // thread = (column / row t, half hf of the summation range)
  cx<T>* S = reinterpret_cast<cx<T>*>(smem_raw + L.S);
  cx<T>* vv = reinterpret_cast<cx<T>*>(smem_raw + L.rot);        // 64 reflector entries
  T* red = reinterpret_cast<T*>(vv + F64_N);                     // 4 partial norms (one rotation-table slot = 2 reals x 2)
  cx<T>* part = reinterpret_cast<cx<T>*>(smem_raw + L.total);    // 2 x 64 partial dot products (behind the one-kernel layout)
  const int d = p.d;
  const size_t tsz = (size_t)d * F64_N;
  for (int64_t kk = blockIdx.x; kk < p.n_chunk; kk += gridDim.x) {
    const int64_t pid = p.pid_offset + kk;
    int64_t ia, ib;
    if (p.pair_mode == 1) { ia = pid / p.NB; ib = pid - ia * p.NB; }
    else if (p.pair_mode == 2) { ib = pid / p.NA; ia = pid - ib * p.NA; }
    else { ia = pid < p.NA ? pid : p.NA - 1; ib = pid < p.NB ? pid : p.NB - 1; }
    const cx<T>* __restrict__ Ag = reinterpret_cast<const cx<T>*>(p.A) + ia * tsz;
    const cx<T>* __restrict__ Bg = reinterpret_cast<const cx<T>*>(p.B) + ib * tsz;
    // ---- column t = (jj, ll) of E (E^dagger for the left fixed point); half hf builds rows (i, 0..7), i = 4 hf .. 4 hf + 3
    {
      const int jj = t >> 3, ll = t & 7;
      const int sa = p.left ? 1 : 8;
      const int oa = p.left ? jj * 8 : jj, ob = p.left ? ll * 8 : ll;
#pragma unroll 1
      for (int i = 4 * hf; i < 4 * hf + 4; ++i) {
        cx<T> acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = mk<T>(0, 0);
#pragma unroll 1
        for (int s = 0; s < d; ++s) {
          const cx<T> a = Ag[s * 64 + i * sa + oa];
#pragma unroll
          for (int k = 0; k < 8; ++k) cmad_c(acc[k], a, Bg[s * 64 + k * sa + ob]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (p.left) acc[k].im = -acc[k].im;
          S[F64S(i * 8 + k, t)] = acc[k];
        }
      }
    }
    __syncthreads();
    // ---- Householder reduction to Hessenberg form, in place
#pragma unroll 1
    for (int k = 0; k + 2 < F64_N; ++k) {
      const cx<T> xk = S[F64S(t, k)];                             // column k, my row
      const cx<T> alpha = S[F64S(k + 1, k)];
      T pn = (hf == 0 && t > k + 1) ? norm2(xk) : T(0);
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) pn += __shfl_xor_sync(0xffffffffu, pn, m);
      if ((tid & 31) == 0) red[tid >> 5] = pn;
      __syncthreads();
      const T xn2 = red[0] + red[1];
      if ((xn2 == T(0)) && (alpha.im == T(0))) { __syncthreads(); continue; }   // already reduced (CTA-uniform)
      T beta = sqrt(norm2(alpha) + xn2);
      if (alpha.re > T(0)) beta = -beta;
      const T ibeta = T(1) / beta;
      const cx<T> tau = mk<T>((beta - alpha.re) * ibeta, -alpha.im * ibeta);
      const cx<T> scal = cinv(alpha - mk<T>(beta, 0));
      if (hf == 0) {
        cx<T> v = mk<T>(0, 0);                                    // my component of the scaled reflector (v[k+1] = 1)
        if (t == k + 1) v = mk<T>(1, 0);
        else if (t > k + 1) v = xk * scal;
        vv[t] = v;
        if (t == k + 1) S[F64S(t, k)] = mk<T>(beta, 0);
        else if (t > k + 1) S[F64S(t, k)] = mk<T>(0, 0);
      }
      __syncthreads();
      // the summation range k+1 .. 63 in two halves
      const int mid = (k + 1 + F64_N) >> 1;
      const int lo = hf ? mid : k + 1, hi = hf ? F64_N : mid;
      // left:  H <- (1 - conj(tau) v v^H) H on my column (columns <= k have nothing to update)
      {
        cx<T> w0 = mk<T>(0, 0), w1 = mk<T>(0, 0);
        if (t > k) {
          int i = lo;
#pragma unroll 2
          for (; i + 1 < hi; i += 2) {
            cmad(w0, conj(vv[i]), S[F64S(i, t)]);
            cmad(w1, conj(vv[i + 1]), S[F64S(i + 1, t)]);
          }
          if (i < hi) cmad(w0, conj(vv[i]), S[F64S(i, t)]);
        }
        part[hf * F64_N + t] = w0 + w1;
        __syncthreads();
        if (t > k) {
          const cx<T> w = (part[t] + part[F64_N + t]) * conj(tau);
#pragma unroll 4
          for (int r = lo; r < hi; ++r) { cx<T> h = S[F64S(r, t)]; cmsub(h, vv[r], w); S[F64S(r, t)] = h; }
        }
      }
      __syncthreads();
      // right: H <- H (1 - tau v v^H) on my row
      {
        cx<T> u0 = mk<T>(0, 0), u1 = mk<T>(0, 0);
        int j = lo;
#pragma unroll 2
        for (; j + 1 < hi; j += 2) {
          cmad(u0, S[F64S(t, j)], vv[j]);
          cmad(u1, S[F64S(t, j + 1)], vv[j + 1]);
        }
        if (j < hi) cmad(u0, S[F64S(t, j)], vv[j]);
        part[hf * F64_N + t] = u0 + u1;
        __syncthreads();
        const cx<T> u = (part[t] + part[F64_N + t]) * tau;
#pragma unroll 4
        for (int c = lo; c < hi; ++c) { cx<T> h = S[F64S(t, c)]; cmsub(h, u, conj(vv[c])); S[F64S(t, c)] = h; }
      }
      __syncthreads();
    }
    // ---- pack: row r, columns r-1..63 (half hf takes rows 32 hf .. 32 hf + 31)
    cx<T>* __restrict__ W = reinterpret_cast<cx<T>*>(p.ws) + (size_t)kk * F64P_SIZE;
#pragma unroll 4
    for (int r = 32 * hf; r < 32 * hf + 32; ++r)
      if (t >= r - 1) W[f64p_row(r) + t] = S[F64S(r, t)];
    if (tid == 0) W[0] = mk<T>(0, 0);                            // the one unused slot (row 0 has no column -1)
    __syncthreads();
  }
}

// ---- sweeps on the packed tile ----------------------------------------------------------------------------------------
// Left phase, rotations i0..i1 (rows (i-1, i)), lane = column; as fp64w_left, plus: row bases come from f64p_row,
// loads of entries below the sub-diagonal give zero, stores there are dropped.
template <typename T, bool LO, bool HI, int OWN, int NEXT>
__device__ __forceinline__ void fp64p_left(cx<T>* S, cx<T>* rot, int ln, int i0, int i1, int en, cx<T>& c, cx<T>& s,
                                           cx<T>& pu0, cx<T>& pu1, cx<T>& qn0, cx<T>& qn1) {
  const int l32 = ln + 32;
  cx<T>* pr = rot + 2 * i0 + 2;
#pragma unroll 2
  for (int i = i0; i <= i1; ++i) {
    const cx<T> ql0 = qn0, ql1 = qn1;
    const int inx = i < en ? i + 1 : i;                        // row en + 1 is never used: re-read row en instead
    const cx<T>* pn = S + f64p_row(inx);
    cx<T>* pt = S + f64p_row(i - 1);
    if (LO) qn0 = (ln >= inx - 1) ? pn[ln] : mk<T>(0, 0);
    if (HI) qn1 = (l32 >= inx - 1) ? pn[l32] : mk<T>(0, 0);
    const bool mine = ln == ((i - 1) & 31);
    cx<T> bot0 = pu0, bot1 = pu1, top0, top1;
    if (LO) {
      top0 = conj(c) * pu0; cmad(top0, conj(s), ql0);
      bot0 = c * ql0; cmsub(bot0, s, pu0);
      if (OWN == 0 && mine) bot0 = mk<T>(0, 0);
    }
    if (HI) {
      top1 = conj(c) * pu1; cmad(top1, conj(s), ql1);
      bot1 = c * ql1; cmsub(bot1, s, pu1);
      if (OWN == 1 && mine) bot1 = mk<T>(0, 0);
    }
    cx<T> cn, sn;
    T nrn;
    if (NEXT == 0) givens_gen<T>(bot0, qn0, cn, sn, nrn);
    else givens_gen<T>(bot1, qn1, cn, sn, nrn);
    if (ln == (i & 31) && i < en) { pr[0] = cn; pr[1] = sn; }
    if (LO && ln >= i - 2) pt[ln] = top0;
    if (HI && l32 >= i - 2) pt[l32] = top1;
    pu0 = bot0; pu1 = bot1;
    __syncwarp();
    c = pr[0]; s = pr[1];
    pr += 2;
  }
}

// Right phase, rotations j0..j1 (columns (j-1, j)), lane = row; as fp64w_right with the same predication.
template <typename T, bool LO, bool HI>
__device__ __forceinline__ void fp64p_right(cx<T>* S, const cx<T>* rot, int ln, int j0, int j1, int en, cx<T>& xl0, cx<T>& xl1) {
  const int l32 = ln + 32;
  cx<T>* pc0 = S + f64p_row(ln) + j0;                          // my rows, column j
  cx<T>* pc1 = S + f64p_row(l32) + j0;
  const cx<T>* pr = rot + 2 * j0;
  cx<T> yn0 = mk<T>(0, 0), yn1 = mk<T>(0, 0), cnx = pr[0], snx = pr[1];
  if (LO && ln <= j0 + 1) yn0 = pc0[0];
  if (HI && l32 <= j0 + 1) yn1 = pc1[0];
#pragma unroll 2
  for (int j = j0; j <= j1; ++j) {
    const cx<T> c = cnx, s = snx, yr0 = yn0, yr1 = yn1;
    const int dn = j < en ? 1 : 0;
    cnx = pr[2 * dn]; snx = pr[2 * dn + 1];
    if (LO) yn0 = (ln <= j + dn + 1) ? pc0[dn] : mk<T>(0, 0);
    if (HI) yn1 = (l32 <= j + dn + 1) ? pc1[dn] : mk<T>(0, 0);
    if (LO) {
      cx<T> a = xl0 * c; cmad(a, yr0, s);
      cx<T> b = yr0 * conj(c); cmsub(b, xl0, conj(s));
      if (ln <= j) pc0[-1] = a;
      xl0 = b;
    }
    if (HI) {
      cx<T> a = xl1 * c; cmad(a, yr1, s);
      cx<T> b = yr1 * conj(c); cmsub(b, xl1, conj(s));
      if (l32 <= j) pc1[-1] = a;
      xl1 = b;
    }
    pc0 += 1; pc1 += 1; pr += 2;
  }
}

// ---- kernel 2: shifted QR on the packed Hessenberg matrices, all eigenvalues, keep the one of largest modulus -------
template <typename T>
__global__ void __launch_bounds__(32)
fp64p_qr_kernel(FpParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Fp64pLayout<T> L = fp64p_layout<T>();
  const int ln = threadIdx.x & 31, l32 = ln + 32;
  cx<T>* S = reinterpret_cast<cx<T>*>(smem_raw + L.S);
  cx<T>* rot = reinterpret_cast<cx<T>*>(smem_raw + L.rot);
  const T eps = eps_of<T>::v();
  const int maxit = 60;
  const unsigned FULL = 0xffffffffu;
  const int rb0 = f64p_row(ln), rb1 = f64p_row(l32);           // bases of my two rows
  const int rbm0 = f64p_row(ln > 0 ? ln - 1 : 0), rbm1 = f64p_row(l32 - 1);

  for (int64_t k = blockIdx.x; k < p.n_chunk; k += gridDim.x) {
    const int64_t pid = p.pid_offset + k;
    {
      const cx<T>* __restrict__ W = reinterpret_cast<const cx<T>*>(p.ws) + (size_t)k * F64P_SIZE;
#pragma unroll 4
      for (int e = ln; e < F64P_SIZE; e += 32) S[e] = W[e];
    }
    __syncwarp();
    int en = F64_N - 1, its = 0, fail = 0, sweeps = 0;
    T best2 = T(-1);
    cx<T> best = mk<T>(0, 0);
#pragma unroll 1
    for (;;) {
      // negligible sub-diagonal entries, all at once (bit r: H[r][r-1] is negligible)
      bool neg0 = false, neg1;
      if (ln >= 1) {
        T sc = cabs1(S[rbm0 + ln - 1]) + cabs1(S[rb0 + ln]);
        if (sc == T(0)) sc = T(1);
        neg0 = cabs1(S[rb0 + ln - 1]) <= eps * sc;
      }
      {
        T sc = cabs1(S[rbm1 + l32 - 1]) + cabs1(S[rb1 + l32]);
        if (sc == T(0)) sc = T(1);
        neg1 = cabs1(S[rb1 + l32 - 1]) <= eps * sc;
      }
      const unsigned long long bits = (unsigned long long)__ballot_sync(FULL, neg0) | ((unsigned long long)__ballot_sync(FULL, neg1) << 32);
      int l = 0;
      while (en >= 0) {
        const unsigned long long m = bits & ((2ull << en) - 1ull) & ~1ull;
        l = m ? (63 - __clzll((long long)m)) : 0;
        if (l == en || its >= maxit) {
          if (l != en) fail = 1;
          const cx<T> ev = S[F64P(en, en)];
          const T a2 = norm2(ev);
          if (a2 > best2) { best2 = a2; best = ev; }
          --en; its = 0;
        } else break;
      }
      if (en < 0) break;                                         // warp-uniform
      // shift (Wilkinson; exceptional every 10 stalled sweeps)
      cx<T> sigma;
      {
        const cx<T> a = S[F64P(en - 1, en - 1)], b = S[F64P(en - 1, en)];
        const cx<T> c = S[F64P(en, en - 1)], dd = S[F64P(en, en)];
        if (its == 10 || its == 20 || its == 30 || its == 40) {
          const T t = fabs(c.re) + (en >= 2 ? fabs(S[F64P(en - 1, en - 2)].re) : T(0));
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
      __syncwarp();
      if (ln >= l && ln <= en) S[rb0 + ln] = S[rb0 + ln] - sigma;                // H - sigma on the window's diagonal
      if (l32 >= l && l32 <= en) S[rb1 + l32] = S[rb1 + l32] - sigma;
      if (l >= 1 && ln == 0) S[F64P(l, l - 1)] = mk<T>(0, 0);                    // the negligible entry becomes exact
      __syncwarp();
      // left phase (lane = column): see kernels_fp64w.cuh for the ranges
      {
        const bool hi_half = en >= 32;
        const int bl = f64p_row(l), bl1 = f64p_row(l + 1);
        cx<T> pu0 = (ln >= l - 1) ? S[bl + ln] : mk<T>(0, 0), pu1 = S[bl + l32];
        cx<T> qn0 = (ln >= l) ? S[bl1 + ln] : mk<T>(0, 0), qn1 = (l32 >= l) ? S[bl1 + l32] : mk<T>(0, 0);
        if (l32 < l - 1) pu1 = mk<T>(0, 0);
        cx<T> c, s;
        T nr;
        givens_gen<T>(S[bl + l], S[bl1 + l], c, s, nr);
        if (ln == 0) { rot[2 * (l + 1)] = c; rot[2 * (l + 1) + 1] = s; }
#define FP64P_SEG(LO_, HI_, OWN_, NEXT_, A_, B_)                                                                  \
        { const int a_ = (A_) > l + 1 ? (A_) : l + 1, b_ = (B_) < en ? (B_) : en;                                   \
          if (a_ <= b_) fp64p_left<T, LO_, HI_, OWN_, NEXT_>(S, rot, ln, a_, b_, en, c, s, pu0, pu1, qn0, qn1); }
        if (hi_half) {
          FP64P_SEG(true, true, 0, 0, 0, 31)
          FP64P_SEG(true, true, 0, 1, 32, 32)
          FP64P_SEG(true, true, 1, 1, 33, 33)
          FP64P_SEG(false, true, 1, 1, 34, 63)
        } else {
          FP64P_SEG(true, false, 0, 0, 0, 31)
        }
#undef FP64P_SEG
        const int be = f64p_row(en);
        if (en <= 33 && ln >= en - 1) S[be + ln] = pu0;
        if (hi_half && l32 >= en - 1) S[be + l32] = pu1;
      }
      __syncwarp();
      // right phase (lane = row)
      {
        const bool lo_rows = l < 32;
        cx<T> xl0 = (ln <= l + 1) ? S[rb0 + l] : mk<T>(0, 0), xl1 = (l32 <= l + 1) ? S[rb1 + l] : mk<T>(0, 0);
        if (lo_rows) {
          const int e_lo = en < 31 ? en : 31;
          if (l + 1 <= e_lo) fp64p_right<T, true, false>(S, rot, ln, l + 1, e_lo, en, xl0, xl1);
          const int b_hi = l + 1 > 32 ? l + 1 : 32;
          if (b_hi <= en) fp64p_right<T, true, true>(S, rot, ln, b_hi, en, en, xl0, xl1);
        } else {
          fp64p_right<T, false, true>(S, rot, ln, l + 1, en, en, xl0, xl1);
        }
        if (lo_rows && ln <= en) S[rb0 + en] = xl0;
        if (en >= 32 && l32 <= en) S[rb1 + en] = xl1;
      }
      if (ln >= l && ln <= en) S[rb0 + ln] = S[rb0 + ln] + sigma;                // my own rows: no barrier needed
      if (l32 >= l && l32 <= en) S[rb1 + l32] = S[rb1 + l32] + sigma;
      __syncwarp();
      ++its; ++sweeps;
    }
    if (ln == 0) {
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
