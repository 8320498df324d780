// qmps_b200 small boundary kernels: index shuffles, merge, unitary completion,
// rotosolve closed forms, argmin, and the large-D power method (SIMT version).
#pragma once
#include <cuda_runtime.h>
#include "kernels_generic.cuh"

namespace qmps {

// a1: A[n][s][i][j] = U[n][2i+s][j]   (qmps/tools.py:151-154)
template <typename T>
__global__ void u2t_kernel(int D, int64_t N, const cx<T>* __restrict__ U, cx<T>* __restrict__ A) {
  const int64_t per = 2LL * D * D;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < N * per; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = t / per;
    const int e = (int)(t - n * per);
    const int s = e / (D * D), ij = e - s * D * D, i = ij / D, j = ij - i * D;
    A[t] = U[n * 4 * D * D + (int64_t)(2 * i + s) * (2 * D) + j];
  }
}

// a7: M[n][(s1,s2)][i][j] = sum_k A[s1][i][k] B[s2][k][j], optionally followed by the
// two-site gate  M <- sum_b W[a][b] M[b]  (qmps/loschmidts/time_evo.py:79).
// One CTA per output block; shared staging so the gate mixes a finished block.
template <typename T>
__global__ void __launch_bounds__(128)
merge_kernel(int d1, int d2, int D, int64_t NA, const cx<T>* __restrict__ A, int64_t NB,
             const cx<T>* __restrict__ B, int64_t NW, const cx<T>* __restrict__ W, int64_t N,
             cx<T>* __restrict__ M) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cx<T>* blk = reinterpret_cast<cx<T>*>(smem_raw);
  Grp g; g.lane = threadIdx.x; g.size = blockDim.x; g.mask = 0xffffffffu; g.cta = 1;
  const int dd = d1 * d2, DD = D * D;
  for (int64_t n = blockIdx.x; n < N; n += gridDim.x) {
    const cx<T>* a = A + (n < NA ? n : NA - 1) * (size_t)(d1 * DD);
    const cx<T>* b = B + (n < NB ? n : NB - 1) * (size_t)(d2 * DD);
    cx<T>* out = M + n * (size_t)(dd * DD);
    if (W == nullptr) {
      merge_block<T>(g, a, b, d1, d2, D, out);
    } else {
      merge_block<T>(g, a, b, d1, d2, D, blk);
      __syncthreads();
      const cx<T>* w = W + (n < NW ? n : NW - 1) * (size_t)(dd * dd);
      for (int e = threadIdx.x; e < dd * DD; e += blockDim.x) {
        const int aa = e / DD, ij = e - aa * DD;
        cx<T> acc = mk<T>(0, 0);
        for (int bb = 0; bb < dd; ++bb) cmad(acc, w[aa * dd + bb], blk[bb * DD + ij]);
        out[e] = acc;
      }
      __syncthreads();
    }
  }
}

// a3: environment_to_unitary (qmps/tools.py:97-108).  V = alpha (1 - 2 w w^dagger/|w|^2)
// with w = v/|v| - alpha e0, alpha = -phase(v0): unitary, V e0 = v/|v|.  Column 0 is
// written as v/|v| exactly.  One CTA per problem; V goes straight to global memory.
template <typename T>
__global__ void __launch_bounds__(256)
env2u_kernel(int n, int64_t N, const cx<T>* __restrict__ v_in, cx<T>* __restrict__ V) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cx<T>* w = reinterpret_cast<cx<T>*>(smem_raw);        // n entries: normalised v, then w
  __shared__ T s_red[256];
  for (int64_t pb = blockIdx.x; pb < N; pb += gridDim.x) {
    const cx<T>* v = v_in + pb * (size_t)n;
    T part = T(0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) { cx<T> z = v[i]; w[i] = z; part += norm2(z); }
    s_red[threadIdx.x] = part;
    __syncthreads();
    T nrm2 = T(0);
    for (int i = 0; i < (int)blockDim.x; ++i) nrm2 += s_red[i];
    const T inv = T(1) / sqrt(nrm2);
    const cx<T> v0 = w[0] * inv;
    const T a0 = cabs(v0);
    const cx<T> alpha = (a0 > T(0)) ? (v0 * (T(-1) / a0)) : mk<T>(-1, 0);
    const T wn2 = T(2) * (T(1) + a0);                    // |v/|v| - alpha e0|^2
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      cx<T> z = w[i] * inv;
      if (i == 0) z = z - alpha;
      w[i] = z;
    }
    __syncthreads();
    cx<T>* out = V + pb * (size_t)n * n;
    const T two_over = T(2) / wn2;
    for (int64_t e = threadIdx.x; e < (int64_t)n * n; e += blockDim.x) {
      const int i = (int)(e / n), j = (int)(e - (int64_t)i * n);
      cx<T> val;
      if (j == 0) {
        val = w[i]; if (i == 0) val = val + alpha;       // v/|v|
      } else {
        cx<T> t = w[i] * conj(w[j]) * two_over;
        val = mk<T>((i == j ? T(1) : T(0)) - t.re, -t.im);
        val = alpha * val;
      }
      out[e] = val;
    }
    __syncthreads();
  }
}

// a2: tensor_to_unitary (qmps/tools.py:123-148): iso[(i,s)][j] = A[s][i][j] completed
// to an m x m unitary (m = d*D) by D Householder reflections; U[:, :D] = iso exactly.
// One CTA per problem; shared: Q (m x D), reflectors (D x m), U (m x m).
template <typename T>
__global__ void __launch_bounds__(128)
t2u_kernel(int d, int D, int64_t N, const cx<T>* __restrict__ A, cx<T>* __restrict__ Uo) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int m = d * D;
  cx<T>* Q = reinterpret_cast<cx<T>*>(smem_raw);         // m x D working copy
  cx<T>* R = Q + m * D;                                  // D reflectors of length m
  cx<T>* U = R + D * m;                                  // m x m
  cx<T>* al = U + m * m;                                 // D diagonal phases
  for (int64_t pb = blockIdx.x; pb < N; pb += gridDim.x) {
    const cx<T>* a = A + pb * (size_t)(d * D * D);
    for (int e = threadIdx.x; e < m * D; e += blockDim.x) {
      const int row = e / D, j = e - row * D, i = row / d, s = row - i * d;
      Q[e] = a[(s * D + i) * D + j];
    }
    __syncthreads();
    for (int k = 0; k < D; ++k) {
      // reflector from column k, rows k..m-1 (redundant scalar work on every thread)
      T xn2 = T(0);
      for (int i = k; i < m; ++i) xn2 += norm2(Q[i * D + k]);
      const T xn = sqrt(xn2);
      const cx<T> x0 = Q[k * D + k];
      const T a0 = cabs(x0);
      const cx<T> alpha = (a0 > T(0)) ? (x0 * (-xn / a0)) : mk<T>(-xn, 0);
      const T wn2 = T(2) * xn * (xn + a0);
      const T iw = wn2 > T(0) ? T(1) / sqrt(wn2) : T(0);
      __syncthreads();
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        cx<T> z = mk<T>(0, 0);
        if (i >= k) { z = Q[i * D + k]; if (i == k) z = z - alpha; z = z * iw; }
        R[k * m + i] = z;
      }
      if (threadIdx.x == 0) al[k] = alpha;
      __syncthreads();
      for (int j = k + 1 + threadIdx.x; j < D; j += blockDim.x) {   // apply to remaining columns
        cx<T> sdot = mk<T>(0, 0);
        for (int i = k; i < m; ++i) cmad(sdot, conj(R[k * m + i]), Q[i * D + j]);
        sdot = sdot * T(2);
        for (int i = k; i < m; ++i) cmsub(Q[i * D + j], R[k * m + i], sdot);
      }
      __syncthreads();
    }
    // U = H_0 ... H_{D-1} diag(alpha, 1)
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
      const int i = e / m, j = e - i * m;
      U[e] = (i == j) ? (j < D ? al[j] : mk<T>(1, 0)) : mk<T>(0, 0);
    }
    __syncthreads();
    for (int k = D - 1; k >= 0; --k) {
      for (int j = threadIdx.x; j < m; j += blockDim.x) {
        cx<T> sdot = mk<T>(0, 0);
        for (int i = k; i < m; ++i) cmad(sdot, conj(R[k * m + i]), U[i * m + j]);
        sdot = sdot * T(2);
        for (int i = k; i < m; ++i) cmsub(U[i * m + j], R[k * m + i], sdot);
      }
      __syncthreads();
    }
    cx<T>* out = Uo + pb * (size_t)(m * m);
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
      const int row = e / m, j = e - row * m;
      if (j < D) { const int i = row / d, s = row - i * d; out[e] = a[(s * D + i) * D + j]; }
      else out[e] = U[e];
    }
    __syncthreads();
  }
}

// a12: rotosolve closed forms (double precision, one thread per parameter vector)
__device__ __forceinline__ double wrap_pi(double x) { return atan2(sin(x), cos(x)); }

__global__ void rotosolve_fit_kernel(int64_t N, int nshift, const double* __restrict__ cost,
                                     double* __restrict__ theta_star, double* __restrict__ fit,
                                     double* __restrict__ theta_io, int P, int coord) {
  const double PI = 3.14159265358979323846;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    double th;
    if (nshift == 3) {
      const double e0 = cost[n * 3], ep = cost[n * 3 + 1], em = cost[n * 3 + 2];
      th = -PI / 2 - atan2(2 * e0 - ep - em, ep - em);              // qmps/rotosolve.py:175
    } else {
      const double M0 = cost[n * 6], Mpi = cost[n * 6 + 1], Mp2 = cost[n * 6 + 2], Mm2 = cost[n * 6 + 3],
                   Mp4 = cost[n * 6 + 4], Mm4 = cost[n * 6 + 5];
      const double A = M0 + Mpi, B = M0 - Mpi, C = Mp2 + Mm2, Dd = Mp2 - Mm2, E = Mp4 - Mm4;
      const double a = 0.25 * (2 * E - 1.4142135623730951 * Dd), b = 0.25 * (A - C), c = 0.5 * Dd, d = 0.5 * B;
      const double Pp = sqrt(a * a + b * b), u = atan2(b, a), Q = sqrt(c * c + d * d), v = atan2(d, c);
      if (fit) { double* f = fit + n * 8; f[0] = a; f[1] = b; f[2] = c; f[3] = d; f[4] = Pp; f[5] = u; f[6] = Q; f[7] = v; }
      // global minimiser of f(x) = P sin(2x+u) + Q sin(x+v) on [-pi, pi]: 128 samples, then
      // bisection on f'(x) inside the bracketing cell
      const int NS = 128;
      const double h = 2 * PI / NS;
      double bx = -PI, bf = 1e300;
      for (int k = 0; k < NS; ++k) {
        const double xx = -PI + k * h, fv = Pp * sin(2 * xx + u) + Q * sin(xx + v);
        if (fv < bf) { bf = fv; bx = xx; }
      }
      double lo = bx - h, hi = bx + h;
      double dlo = 2 * Pp * cos(2 * lo + u) + Q * cos(lo + v), dhi = 2 * Pp * cos(2 * hi + u) + Q * cos(hi + v);
      th = bx;
      if (dlo < 0 && dhi > 0) {
        for (int it = 0; it < 60; ++it) {
          const double mid = 0.5 * (lo + hi), dm = 2 * Pp * cos(2 * mid + u) + Q * cos(mid + v);
          if (dm < 0) lo = mid; else hi = mid;
        }
        th = 0.5 * (lo + hi);
      }
    }
    if (theta_star) theta_star[n] = th;
    if (theta_io) {
      double t = theta_io[n * P + coord] + wrap_pi(th);
      if (nshift == 3) t = wrap_pi(t);                               // qmps/rotosolve.py:177
      theta_io[n * P + coord] = t;
    }
  }
}


// a13: exact TFIM Loschmidt rate function (qmps/loschmidts/exact_loschmidt.py:6-20).
// loschmidt(t) = f(it) + f(-it) = -(1/pi) int_0^pi log|cos^2(phi_k) + sin^2(phi_k) e^{-2it eps_k}| dk
// (the two terms are complex conjugates).  One warp per time: composite 8-point
// Gauss-Legendre on 256 panels, lanes stride over panels, shuffle reduction.
__global__ void __launch_bounds__(128)
loschmidt_rate_kernel(int64_t NT, const double* __restrict__ t, double g0, double g1, double* __restrict__ out) {
  const double PI = 3.14159265358979323846;
  const double gx[8] = {-0.96028985649753618, -0.79666647741362673, -0.52553240991632899, -0.18343464249564978,
                        0.18343464249564978, 0.52553240991632899, 0.79666647741362673, 0.96028985649753618};
  const double gw[8] = {0.10122853629037706, 0.22238103445337443, 0.31370664587788688, 0.36268378337836166,
                        0.36268378337836166, 0.31370664587788688, 0.22238103445337443, 0.10122853629037706};
  const int lane = threadIdx.x & 31;
  const int NPANEL = 256;
  for (int64_t it = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < NT; it += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const double tt = t[it];
    double acc = 0.0;
    for (int pnl = lane; pnl < NPANEL; pnl += 32) {
      const double a = PI * pnl / NPANEL, h = PI / NPANEL;
      for (int q = 0; q < 8; ++q) {
        const double k = a + 0.5 * h * (1.0 + gx[q]);
        double sk, ck;
        sincos(k, &sk, &ck);
        const double phi = 0.5 * (atan2(sk, g0 - ck) - atan2(sk, g1 - ck));
        const double eps = -2.0 * sqrt((g1 - ck) * (g1 - ck) + sk * sk);
        double sp, cp, s2, c2;
        sincos(phi, &sp, &cp);
        sincos(-2.0 * tt * eps, &s2, &c2);
        const double re = cp * cp + sp * sp * c2, im = sp * sp * s2;
        acc += gw[q] * 0.5 * h * 0.5 * log(re * re + im * im);     // log|w|
      }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[it] = -acc / PI;
  }
}

// (e): block-level argmin; the last block to finish reduces the per-block results.
__global__ void __launch_bounds__(256)
argmin_kernel(int64_t N, const double* __restrict__ cost, int64_t offset, double* __restrict__ blk_cost,
              int64_t* __restrict__ blk_idx, unsigned int* __restrict__ counter, double* __restrict__ best_cost,
              int64_t* __restrict__ best_idx) {
  __shared__ double s_c[256];
  __shared__ int64_t s_i[256];
  __shared__ bool is_last;
  double bc = INFINITY;
  int64_t bi = INT64_MAX;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < N; t += (int64_t)gridDim.x * blockDim.x) {
    const double c = cost[t];
    if (c < bc) { bc = c; bi = t; }                       // increasing t: first minimum wins
  }
  s_c[threadIdx.x] = bc; s_i[threadIdx.x] = bi;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      const double c2 = s_c[threadIdx.x + s]; const int64_t i2 = s_i[threadIdx.x + s];
      if (c2 < s_c[threadIdx.x] || (c2 == s_c[threadIdx.x] && i2 < s_i[threadIdx.x])) { s_c[threadIdx.x] = c2; s_i[threadIdx.x] = i2; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    blk_cost[blockIdx.x] = s_c[0]; blk_idx[blockIdx.x] = s_i[0];
    __threadfence();
    const unsigned int prev = atomicAdd(counter, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    bc = INFINITY; bi = INT64_MAX;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
      const double c = blk_cost[b]; const int64_t i = blk_idx[b];
      if (c < bc || (c == bc && i < bi)) { bc = c; bi = i; }
    }
    s_c[threadIdx.x] = bc; s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) {
        const double c2 = s_c[threadIdx.x + s]; const int64_t i2 = s_i[threadIdx.x + s];
        if (c2 < s_c[threadIdx.x] || (c2 == s_c[threadIdx.x] && i2 < s_i[threadIdx.x])) { s_c[threadIdx.x] = c2; s_i[threadIdx.x] = i2; }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      best_cost[0] = s_c[0];
      best_idx[0] = (s_i[0] == INT64_MAX) ? -1 : s_i[0] + offset;
      *counter = 0;
    }
  }
}

// ---- large-D power method, SIMT tiles ---------------------------------------------------
// Batched complex GEMM  C[b] = scale[b] * sum_t A[b*nsum+t] . op(B[.])  with 32 x 32
// output tiles, 8 x 32 k-slabs in shared memory, 2 x 2 micro-tiles per thread.
//   trans_b = 0:  C = A (M x K) . B (K x N)         B index = b / b_div
//   trans_b = 1:  C = sum_t A_t (M x K) . B_t^H     B_t is (N x K), index b*nsum + t
template <typename T>
__global__ void __launch_bounds__(256)
zgemm_tile_kernel(int M, int Nn, int K, int nsum, const cx<T>* __restrict__ A, const cx<T>* __restrict__ B,
                  int trans_b, int b_div, cx<T>* __restrict__ C, const T* __restrict__ inv_scale) {
  __shared__ cx<T> sA[32][9];
  __shared__ cx<T> sB[8][33];
  const int b = blockIdx.z, tm = blockIdx.y * 32, tn = blockIdx.x * 32;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // 16 x 16 threads, 2 x 2 each
  cx<T> acc[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j] = mk<T>(0, 0);
  for (int t = 0; t < nsum; ++t) {
    const cx<T>* a = A + ((size_t)b * nsum + t) * M * K;
    const cx<T>* bb = trans_b ? B + ((size_t)b * nsum + t) * Nn * K : B + (size_t)(b / b_div) * K * Nn;
    for (int k0 = 0; k0 < K; k0 += 8) {
      {  // 32 x 8 slab of A: 256 threads, one element each
        const int r = threadIdx.x >> 3, c = threadIdx.x & 7;
        sA[r][c] = (tm + r < M && k0 + c < K) ? a[(size_t)(tm + r) * K + k0 + c] : mk<T>(0, 0);
      }
      {  // 8 x 32 slab of op(B)
        if (trans_b) {
          const int nn = threadIdx.x >> 3, c = threadIdx.x & 7;   // B_t[n][k] contiguous in k
          sB[c][nn] = (tn + nn < Nn && k0 + c < K) ? conj(bb[(size_t)(tn + nn) * K + k0 + c]) : mk<T>(0, 0);
        } else {
          const int c = threadIdx.x >> 5, nn = threadIdx.x & 31;
          sB[c][nn] = (tn + nn < Nn && k0 + c < K) ? bb[(size_t)(k0 + c) * Nn + tn + nn] : mk<T>(0, 0);
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const cx<T> a0 = sA[ty * 2][kk], a1 = sA[ty * 2 + 1][kk];
        const cx<T> b0 = sB[kk][tx * 2], b1 = sB[kk][tx * 2 + 1];
        cmad(acc[0][0], a0, b0); cmad(acc[0][1], a0, b1);
        cmad(acc[1][0], a1, b0); cmad(acc[1][1], a1, b1);
      }
      __syncthreads();
    }
  }
  const T sc = inv_scale ? inv_scale[trans_b ? b : b / b_div] : T(1);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = tm + ty * 2 + i, c = tn + tx * 2 + j;
      if (r < M && c < Nn) C[(size_t)b * M * Nn + (size_t)r * Nn + c] = acc[i][j] * sc;
    }
}

// per-problem 1/|r|_F  (one CTA per problem)
template <typename T>
__global__ void __launch_bounds__(256)
inv_norm_kernel(int64_t len, const cx<T>* __restrict__ r, T* __restrict__ inv_norm) {
  __shared__ T s[256];
  const cx<T>* p = r + (size_t)blockIdx.x * len;
  T part = T(0);
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) part += norm2(p[i]);
  s[threadIdx.x] = part;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) { if ((int)threadIdx.x < k) s[threadIdx.x] += s[threadIdx.x + k]; __syncthreads(); }
  if (threadIdx.x == 0) inv_norm[blockIdx.x] = T(1) / sqrt(s[0]);
}

// r <- r * inv_norm ; optional rayleigh[b] = <r_old_normalised, Er> where Er is passed in `er`
template <typename T>
__global__ void __launch_bounds__(256)
scale_kernel(int64_t len, cx<T>* __restrict__ r, const T* __restrict__ inv_norm) {
  cx<T>* p = r + (size_t)blockIdx.x * len;
  const T sc = inv_norm[blockIdx.x];
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) p[i] = p[i] * sc;
}
template <typename T>
__global__ void __launch_bounds__(256)
vdot_kernel(int64_t len, const cx<T>* __restrict__ a, const cx<T>* __restrict__ b, cx<T>* __restrict__ out) {
  __shared__ T sr[256], si[256];
  const cx<T>* pa = a + (size_t)blockIdx.x * len;
  const cx<T>* pb = b + (size_t)blockIdx.x * len;
  cx<T> acc = mk<T>(0, 0);
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) cmad(acc, conj(pa[i]), pb[i]);
  sr[threadIdx.x] = acc.re; si[threadIdx.x] = acc.im;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if ((int)threadIdx.x < k) { sr[threadIdx.x] += sr[threadIdx.x + k]; si[threadIdx.x] += si[threadIdx.x + k]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = mk<T>(sr[0], si[0]);
}

}  // namespace qmps
