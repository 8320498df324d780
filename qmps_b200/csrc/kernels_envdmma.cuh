// qmps_b200 D = 8 complex128 environment / energy kernel with the elimination on the FP64 tensor pipe
// (a4/a5 TransferMatrix(A).eigs() + cholesky, qmps/tools.py:176-182; a9 energy, qmps/ground_state.py:150-168,
// 251-266; a12 rotosolve shift fan-out -- BASELINE config 4).
//
// Same real linear system as envreal.cuh (r = fixed point of Phi(r) = sum_s A_s r A_s^dagger, Hermitian, so
// D^2 = 64 real unknowns), but
//   * d_0 = r[0][0] is eliminated with the trace condition, d_0 = 1 - sum_{j>0} d_j, and the (redundant) equation 0
//     is dropped: 63 equations x (63 unknowns + right-hand side) = a 64 x 64 real array with one zero padding row --
//     exactly 8 x 8 tiles of the m8n8k4 FP64 MMA;
//   * the array lives in the MMA ACCUMULATOR fragments of two warps (warp w owns tile rows 4w .. 4w+3, all eight tile
//     columns: 64 doubles per lane) for the whole solve;
//   * Gauss-Jordan runs BLOCKED, four columns at a time (16 block steps instead of 64 steps):
//       panel   the four columns go through shared memory to one-thread-per-row form; four pivot steps with partial
//               pivoting (implicit row permutation, arg-max by redux.sync as in kernels_envreal.cuh) act on the
//               64 x 4 panel P and on W = T E_R, the image of the pivot-row unit vectors under the step's row
//               transformation T, which gives T = I - L E_R^T with L = E_R - W;
//       update  M <- M - L . M[R, :] is ONE rank-4 product per tile: 4 x (8 - K/2) DMMAs per warp with the A fragment
//               (-L, 64 x 4) and the B fragment (the four raw pivot rows) read from shared memory.
//     The row-per-thread kernel issued one broadcast LDS.128 per two DFMAs (LSU return path 75 % busy, FP64 pipe 17 %:
//     profiles/ncu_er8_src_r01f.txt); here a lane loads 4 + (8 - K/2) doubles per 4 (8 - K/2) MMAs of 256 FMAs.
//
// Index maps (chosen so that fragments are cheap to build):
//   columns  c = 2 p + part, p = 0..27: x_jl / y_jl = Re / Im r[j][l] of pair p = (j < l) in the order (0,1),(0,2),..;
//            c = 55 + j, j = 1..7: d_j;   c = 63: right-hand side.
//   rows     rho = 8 rb + q (tile row rb, row q inside the tile).  Pair p = 8 m + q (i < k), m = 0..3, has its real
//            part equation in tile row 2 m and its imaginary part equation in tile row 2 m + 1 -- both in the SAME lane,
//            which needs G(j,l) = sum_s A[s,i,j] conj(A[s,k,l]) and G(l,j) only once for the 2 x 2 block
//            (re/im equation) x (x/y unknown).  Slots p = 28..31 (m = 3, q = 4..7) hold the diagonal equations:
//            tile row 6: e = q - 3 (1..4), tile row 7: e = q + 1 (5..7) and the zero padding row (q = 7).
#pragma once
#include <cuda_runtime.h>
#include "envreal.cuh"
#include "kernels_generic.cuh"
#include "kernels_envreal.cuh"

namespace qmps {

struct EdLayout { size_t A, Ap, P, Lb, Ub, rec, cand, x, xraw, r, C, tmp, red, trig, total; };
QMPS_HD EdLayout ed_layout(int d, int nops, int want_tmp) {
  constexpr int D = 8, n = 64;
  EdLayout L;
  Bump b;
  L.A = b.take(sizeof(cx<double>) * (size_t)d * n);
  L.Ap = b.take(sizeof(cx<double>) * (size_t)d * D * (D + 1));
  L.P = b.take(sizeof(double) * n * 4);
  L.Lb = b.take(sizeof(double) * n * 4);
  L.Ub = b.take(sizeof(double) * n * 4);
  L.rec = b.take(sizeof(double) * 2 * 2 * 8);
  L.cand = b.take(sizeof(double) * 2 * 2 * 2);
  L.x = b.take(sizeof(double) * n);
  L.xraw = b.take(sizeof(double) * n);
  L.r = b.take(sizeof(cx<double>) * n);
  L.C = b.take(sizeof(cx<double>) * n);
  L.tmp = b.take(want_tmp ? sizeof(cx<double>) * (size_t)8 * n : 0);
  L.red = b.take(sizeof(double) * n);
  L.trig = b.take(sizeof(double) * 2 * (nops > 0 ? nops : 1));
  L.total = (b.off + 127) & ~size_t(127);
  return L;
}

// pair index p = 0..27 -> (j < l), order (0,1),(0,2),...,(6,7)
QMPS_HD void ed_pair(int p, int* j, int* l) {
  int a = 0;
  while (p >= 7 - a) { p -= 7 - a; ++a; }
  *j = a; *l = a + 1 + p;
}

#if defined(__CUDACC__)

__device__ __forceinline__ void ed_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// G(j,l) = sum_s A[s,i,j] conj(A[s,k,l]); Ai / Ak point at row i / k of the padded tensor (row stride 9, 72 per s)
template <int DP>
__device__ __forceinline__ cx<double> ed_G(const cx<double>* Ai, const cx<double>* Ak, int d, int j, int l) {
  cx<double> g = mk<double>(0, 0);
  if (DP > 0) {
#pragma unroll
    for (int s = 0; s < DP; ++s) cmad_c(g, Ai[s * 72 + j], Ak[s * 72 + l]);
  } else {
    for (int s = 0; s < d; ++s) cmad_c(g, Ai[s * 72 + j], Ak[s * 72 + l]);
  }
  return g;
}

// MODE 0: eta, r, C, status.  MODE 1: energy (+ status).  DP: compile-time physical dimension (0 = run time).
// Only the tile-column index of the elimination is unrolled (it selects accumulator registers); the two halves of a
// tile column and the four pivot steps of a panel are run-time loops, which keeps the kernel inside the instruction
// cache (the fully unrolled first version spent a quarter of its stall samples on instruction fetch).
template <int MODE, int DP>
__global__ void __maxnreg__(168)
env_dmma_kernel(EnvParams p) {
  typedef double T;
  constexpr int D = 8, n = 64;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ unsigned char s_pair[32];
  int gi, gpc;
  const Grp g = make_group<64>(&gi, &gpc);                   // the CTA (two warps) is the group
  const int d = p.d;
  const EdLayout L = ed_layout(d, p.nops, MODE == 1);
  unsigned char* base = smem_raw;
  cx<T>* A = reinterpret_cast<cx<T>*>(base + L.A);
  cx<T>* Ap = reinterpret_cast<cx<T>*>(base + L.Ap);
  T* Pb = reinterpret_cast<T*>(base + L.P);
  T* Lb = reinterpret_cast<T*>(base + L.Lb);
  T* Ub = reinterpret_cast<T*>(base + L.Ub);
  T* recbuf = reinterpret_cast<T*>(base + L.rec);
  T* xs = reinterpret_cast<T*>(base + L.x);
  T* xraw = reinterpret_cast<T*>(base + L.xraw);
  cx<T>* r = reinterpret_cast<cx<T>*>(base + L.r);
  cx<T>* Cc = reinterpret_cast<cx<T>*>(base + L.C);
  cx<T>* tmp = reinterpret_cast<cx<T>*>(base + L.tmp);
  T* red = reinterpret_cast<T*>(base + L.red);
  T* trig = reinterpret_cast<T*>(base + L.trig);
  const int S = (MODE == 1 && p.nshift > 0) ? p.nshift : 1;
  const int64_t total = p.N * S;
  const cx<T>* hmat = reinterpret_cast<const cx<T>*>(p.hmat);
  const int e = threadIdx.x;                                 // panel role: my row of the array
  const int w = e >> 5, lane = e & 31, q = lane >> 2, t = lane & 3;   // fragment role
  // which warp factorises the panels: alternates with the CTA's residency slot (p.ws_stride = SM count), so that the
  // heavier panel warps spread over all four SM sub-partitions instead of piling up on two of them
  const int pw = p.ws_stride ? (int)((blockIdx.x / (unsigned)p.ws_stride) & 1u) : 0;
  if (e < 32) {
    int j = 0, l = 0;
    if (e < 28) ed_pair(e, &j, &l);
    s_pair[e] = (unsigned char)(j | (l << 4));
  }
  __syncthreads();

  for (int64_t pid = blockIdx.x; pid < total; pid += gridDim.x) {
    const int64_t pn = pid / S;
    const int sidx = (int)(pid - pn * S);
    // ---- 1. tensor A[d][D][D] into shared memory (as kernels_envreal.cuh)
    if (p.theta) {
      StateLayout SL; SL.R = 2 * D; SL.ncols = D; SL.a_layout = 1;
      const double sh = p.nshift > 0 ? p.shifts[sidx] : 0.0;
      ansatz_eval<T>(g, p.ops, p.nops, p.theta + pn * p.P, p.nshift > 0 ? p.coord : -1, sh, p.nq, SL, A, trig);
    } else if (p.in_is_U) {
      const cx<T>* U = reinterpret_cast<const cx<T>*>(p.in) + pn * (size_t)(4 * n);
      for (int c = g.lane; c < 2 * n; c += g.size) {
        int s = c / n, ij = c - s * n, i = ij / D, j = ij - i * D;
        A[c] = U[(2 * i + s) * (2 * D) + j];
      }
    } else {
      const cx<T>* src = reinterpret_cast<const cx<T>*>(p.in) + pn * (size_t)(d * n);
      for (int c = g.lane; c < d * n; c += g.size) A[c] = src[c];
    }
    g.sync();
    for (int c = g.lane; c < d * n; c += g.size) {
      const int si = c / D, j = c - si * D;
      Ap[si * (D + 1) + j] = A[c];
    }
    g.sync();

    // ---- 2. my fragments of the 64 x 64 array: acc[rbl][cb][0..1] = M[8 (4 w + rbl) + q][8 cb + 2 t + 0..1]
    double acc[4][8][2];
#pragma unroll
    for (int ml = 0; ml < 2; ++ml) {
      const int pr = 8 * (2 * w + ml) + q;                   // row-pair slot
      const bool regular = pr < 28;
      // rows A (tile row 2 ml) and B (tile row 2 ml + 1): equation (i, k) and which part of G's combination it takes
      int iA, kA, iB, kB;
      bool padB = false;
      if (regular) { const int pk = s_pair[pr]; iA = pk & 15; kA = pk >> 4; iB = iA; kB = kA; }
      else { iA = kA = q - 3; iB = kB = q + 1; if (iB > 7) { iB = kB = 7; padB = true; } }
      const cx<T>* AiA = Ap + iA * 9; const cx<T>* AkA = Ap + kA * 9;
      const cx<T>* AiB = Ap + iB * 9; const cx<T>* AkB = Ap + kB * 9;
#pragma unroll
      for (int cb = 0; cb < 7; ++cb) {
        const int pc = 4 * cb + t;
        const int pk = s_pair[pc];
        const int j = pk & 15, l = pk >> 4;
        const cx<T> g1 = ed_G<DP>(AiA, AkA, d, j, l), g2 = ed_G<DP>(AiA, AkA, d, l, j);
        const T dl = (regular && pr == pc) ? T(1) : T(0);
        acc[2 * ml][cb][0] = (g1.re + g2.re) - dl;           // Re of the x coefficient G(j,l) + G(l,j)
        acc[2 * ml][cb][1] = -(g1.im - g2.im);               // Re of the y coefficient i (G(j,l) - G(l,j))
        if (regular) {
          acc[2 * ml + 1][cb][0] = g1.im + g2.im;
          acc[2 * ml + 1][cb][1] = (g1.re - g2.re) - dl;
        } else {
          const cx<T> h1 = ed_G<DP>(AiB, AkB, d, j, l), h2 = ed_G<DP>(AiB, AkB, d, l, j);
          acc[2 * ml + 1][cb][0] = padB ? T(0) : (h1.re + h2.re);
          acc[2 * ml + 1][cb][1] = padB ? T(0) : -(h1.im - h2.im);
        }
      }
      // tile column 7: d_{2t+1}, d_{2t+2}; for t = 3: d_7 and the right-hand side
      {
        const int j0 = 2 * t + 1, j1 = 2 * t + 2;
        const cx<T> a00 = ed_G<DP>(AiA, AkA, d, 0, 0), a0 = ed_G<DP>(AiA, AkA, d, j0, j0);
        const cx<T> a1 = (t < 3) ? ed_G<DP>(AiA, AkA, d, j1, j1) : mk<T>(0, 0);
        const cx<T> b00 = regular ? a00 : ed_G<DP>(AiB, AkB, d, 0, 0);
        const cx<T> b0 = regular ? a0 : ed_G<DP>(AiB, AkB, d, j0, j0);
        const cx<T> b1 = regular ? a1 : ((t < 3) ? ed_G<DP>(AiB, AkB, d, j1, j1) : mk<T>(0, 0));
        // row A takes the real part; row B the imaginary part (regular) or the real part of its own equation (diagonal)
        const T A00 = a00.re, A0 = a0.re, A1 = a1.re;
        const T B00 = regular ? b00.im : b00.re, B0 = regular ? b0.im : b0.re, B1 = regular ? b1.im : b1.re;
        acc[2 * ml][7][0] = A0 - A00 - ((!regular && iA == j0) ? T(1) : T(0));
        acc[2 * ml][7][1] = (t < 3) ? (A1 - A00 - ((!regular && iA == j1) ? T(1) : T(0))) : -A00;
        acc[2 * ml + 1][7][0] = padB ? T(0) : (B0 - B00 - ((!regular && iB == j0) ? T(1) : T(0)));
        acc[2 * ml + 1][7][1] = padB ? T(0) : ((t < 3) ? (B1 - B00 - ((!regular && iB == j1) ? T(1) : T(0))) : -B00);
      }
    }

    // ---- 3. blocked Gauss-Jordan.  The panel is factorised by warp 0 alone, two rows per lane (rows lane and
    //         lane + 32), pivot rows broadcast by shuffles: no CTA barrier inside the four pivot steps.
    bool doneA = false, doneB = false;
    int bad = 0;
    int colA = 0, colB = 0;
    int* Rb = reinterpret_cast<int*>(recbuf);                // the four pivot rows of the current block
#pragma unroll
    for (int cbK = 0; cbK < 8; ++cbK) {
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int K = 2 * cbK + h;
        const int NJ = (K == 15) ? 3 : 4;
        // panel columns 4K .. 4K+3 from the fragments to row-major form
        if ((t >> 1) == h) {
#pragma unroll
          for (int rbl = 0; rbl < 4; ++rbl)
            *reinterpret_cast<double2*>(Pb + (8 * (4 * w + rbl) + q) * 4 + 2 * (t & 1)) = make_double2(acc[rbl][cbK][0], acc[rbl][cbK][1]);
        }
        g.sync();
        if (w == pw) {
          double pA[4], pB[4], wA[4] = {0.0, 0.0, 0.0, 0.0}, wB[4] = {0.0, 0.0, 0.0, 0.0};
          {
            const double2 a0 = *reinterpret_cast<const double2*>(Pb + lane * 4), a1 = *reinterpret_cast<const double2*>(Pb + lane * 4 + 2);
            const double2 b0 = *reinterpret_cast<const double2*>(Pb + (lane + 32) * 4), b1 = *reinterpret_cast<const double2*>(Pb + (lane + 32) * 4 + 2);
            pA[0] = a0.x; pA[1] = a0.y; pA[2] = a1.x; pA[3] = a1.y;
            pB[0] = b0.x; pB[1] = b0.y; pB[2] = b1.x; pB[3] = b1.y;
          }
          int mineA = -1, mineB = -1;
          int rsel[4] = {0, 0, 0, 0};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < NJ) {
              // arg-max of |column j| over the unused rows: 32-bit key = top bits of the magnitude as a float | (63 - row)
              unsigned kA = doneA ? 0u : ((__float_as_uint((float)fabs(pA[j])) & ~63u) | (unsigned)(63 - lane));
              unsigned kB = doneB ? 0u : ((__float_as_uint((float)fabs(pB[j])) & ~63u) | (unsigned)(31 - lane));
              if (!doneA && kA < 64u) kA = 64u | (unsigned)(63 - lane);
              if (!doneB && kB < 64u) kB = 64u | (unsigned)(31 - lane);
              const unsigned best = __reduce_max_sync(0xffffffffu, kA > kB ? kA : kB);
              const int who = 63 - (int)(best & 63u);
              const int wl = who & 31;
              const bool isB = who >= 32;
              if (!(__uint_as_float(best & ~63u) > 1e-13f)) bad = 1;
              rsel[j] = who;
              // pivot row: panel entries j.., W entries ..j-1, from its lane
              double row[4], wrow[4];
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                if (c >= j) row[c] = __shfl_sync(0xffffffffu, isB ? pB[c] : pA[c], wl);
                if (c < j) wrow[c] = __shfl_sync(0xffffffffu, isB ? wB[c] : wA[c], wl);
              }
              const double inv = 1.0 / row[j];
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                if (c > j) row[c] *= inv;
                if (c < j) wrow[c] *= inv;
              }
              // every other row: x -= x[j] * (scaled pivot row); the pivot row becomes the scaled row
              {
                const bool piv = (who == lane);
                const double f = pA[j];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  if (c > j) pA[c] = piv ? row[c] : fma(-f, row[c], pA[c]);
                  if (c < j) wA[c] = piv ? wrow[c] : fma(-f, wrow[c], wA[c]);
                }
                wA[j] = piv ? inv : -f * inv;
                if (piv) { doneA = true; mineA = j; colA = 4 * K + j; }
              }
              {
                const bool piv = (who == lane + 32);
                const double f = pB[j];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  if (c > j) pB[c] = piv ? row[c] : fma(-f, row[c], pB[c]);
                  if (c < j) wB[c] = piv ? wrow[c] : fma(-f, wrow[c], wB[c]);
                }
                wB[j] = piv ? inv : -f * inv;
                if (piv) { doneB = true; mineB = j; colB = 4 * K + j; }
              }
            }
          }
          // -L = W - E_R
#pragma unroll
          for (int c = 0; c < 4; ++c) { wA[c] -= (mineA == c) ? 1.0 : 0.0; wB[c] -= (mineB == c) ? 1.0 : 0.0; }
          if (K == 15) { wA[3] = 0.0; wB[3] = 0.0; }
          *reinterpret_cast<double2*>(Lb + lane * 4) = make_double2(wA[0], wA[1]);
          *reinterpret_cast<double2*>(Lb + lane * 4 + 2) = make_double2(wA[2], wA[3]);
          *reinterpret_cast<double2*>(Lb + (lane + 32) * 4) = make_double2(wB[0], wB[1]);
          *reinterpret_cast<double2*>(Lb + (lane + 32) * 4 + 2) = make_double2(wB[2], wB[3]);
          if (lane == 0) *reinterpret_cast<int4*>(Rb) = make_int4(rsel[0], rsel[1], rsel[2], rsel[3]);
        }
        g.sync();
        // the raw pivot rows out of their owners' fragments: Ub[col][j]
        {
          const int4 R4 = *reinterpret_cast<const int4*>(Rb);
#pragma unroll 1
          for (int j = 0; j < NJ; ++j) {
            const int gwho = j == 0 ? R4.x : (j == 1 ? R4.y : (j == 2 ? R4.z : R4.w));
            const int prb = gwho >> 3, pq = gwho & 7;
            if ((prb >> 2) == w && q == pq) {
              T* ub = Ub + 8 * t + j;
              switch (prb & 3) {
#define QMPS_ED_PUB(rr)                                                              \
  case rr:                                                                           \
    _Pragma("unroll") for (int cb = 0; cb < 8; ++cb) {                               \
      if (cb >= cbK) { ub[32 * cb] = acc[rr][cb][0]; ub[32 * cb + 4] = acc[rr][cb][1]; } \
    }                                                                                \
    break;
                QMPS_ED_PUB(0) QMPS_ED_PUB(1) QMPS_ED_PUB(2) QMPS_ED_PUB(3)
#undef QMPS_ED_PUB
              }
            }
          }
          if (K == 15) Ub[e * 4 + 3] = 0.0;
        }
        g.sync();
        // rank-4 update of the live tile columns
        {
          double af[4];
#pragma unroll
          for (int rbl = 0; rbl < 4; ++rbl) af[rbl] = Lb[(8 * (4 * w + rbl) + q) * 4 + t];
#pragma unroll
          for (int cb = 0; cb < 8; ++cb) {
            if (cb >= cbK) {
              const double bf = Ub[(8 * cb + q) * 4 + t];
#pragma unroll
              for (int rbl = 0; rbl < 4; ++rbl) ed_dmma(acc[rbl][cb][0], acc[rbl][cb][1], af[rbl], bf);
            }
          }
        }
      }
    }
    // ---- 4. solution: x[c] = (row pivoted at column c)[63]
    if (t == 3) {
#pragma unroll
      for (int rbl = 0; rbl < 4; ++rbl) xraw[8 * (4 * w + rbl) + q] = acc[rbl][7][1];
    }
    g.sync();
    if (w == pw) {
      if (doneA) xs[colA] = xraw[lane];
      if (doneB) xs[colB] = xraw[lane + 32];
    }
    g.sync();
    {
      const int a = e >> 3, b = e & 7;
      cx<T> v;
      if (a == b) {
        if (a == 0) { T s = T(1); for (int j = 1; j < 8; ++j) s -= xs[55 + j]; v = mk<T>(s, 0); }
        else v = mk<T>(xs[55 + a], 0);
      } else {
        const int lo = a < b ? a : b, hi = a < b ? b : a;
        const int pc = lo * 8 - (lo * (lo + 1)) / 2 + (hi - lo - 1);
        v = mk<T>(xs[2 * pc], a < b ? xs[2 * pc + 1] : -xs[2 * pc + 1]);
      }
      r[e] = v;
    }
    g.sync();
    int status = bad ? ST_SINGULAR : ST_OK;
    // ---- 5. outputs (as kernels_envreal.cuh)
    T part = T(0);
    if (MODE == 0 && p.eta) {
      for (int c = g.lane; c < d * D; c += g.size) {
        const cx<T>* row = A + c * D;
        cx<T> s2 = mk<T>(0, 0);
        for (int j = 0; j < D; ++j) {
          cx<T> tt = mk<T>(0, 0);
          for (int l = 0; l < D; ++l) cmad_c(tt, r[j * D + l], row[l]);
          cmad(s2, row[j], tt);
        }
        part += s2.re;
      }
    }
    if (MODE == 0) {
      T eta_r = T(1);
      if (p.eta) eta_r = group_sum<T>(g, part, red);
      if (p.C || p.status) {
        const int cbad = cholesky_lower<T>(g, r, D, Cc, D, D);
        if (cbad && status == ST_OK) status = ST_NOT_PD;
      }
      if (g.lane == 0) {
        if (p.eta) reinterpret_cast<cx<T>*>(p.eta)[pid] = mk<T>(eta_r, 0);
        if (p.status) p.status[pid] = status;
      }
      if (p.r) { cx<T>* o = reinterpret_cast<cx<T>*>(p.r) + pid * (size_t)n; for (int c = g.lane; c < n; c += g.size) o[c] = r[c]; }
      if (p.C) { cx<T>* o = reinterpret_cast<cx<T>*>(p.C) + pid * (size_t)n; for (int c = g.lane; c < n; c += g.size) o[c] = Cc[c]; }
    } else {
      const cx<T>* Mb;
      if (p.two_site) Mb = A;
      else { merge_block<T>(g, A, A, 2, 2, D, tmp); Mb = tmp; }
      const T en = energy_from_block<T>(g, Mb, r, D, hmat, tmp + 4 * n, red);
      if (p.status) {
        const int cbad = cholesky_lower<T>(g, r, D, Cc, D, D);
        if (cbad && status == ST_OK) status = ST_NOT_PD;
      }
      if (g.lane == 0) {
        reinterpret_cast<T*>(p.energy)[pid] = en;
        if (p.status) p.status[pid] = status;
      }
    }
    g.sync();
  }
}

#endif  // __CUDACC__

}  // namespace qmps
