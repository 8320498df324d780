// qmps_b200 large-D transfer-matrix application (cfg 5): batched complex128 GEMM
// on the FP64 tensor pipe (mma.sync.m8n8k4.f64 -> DMMA), cp.async double-buffered
// K-slabs, fused scale / norm epilogues.  tcgen05.mma has no FP64 kind, so for
// complex128 at 1e-10 this is the widest FP64 datapath the SM has.
//
// One application of E_AB to r (qmps.ipynb cells 29-32; SURVEY 8(d) cfg 5):
//     stage 1   T[b,s] = alpha_b * A[b,s] . r[b]                 (BFORM = 0)
//     stage 2   r'[b]  = sum_s T[b,s] . B[b,s]^dagger            (BFORM = 1)
// alpha_b = 1/|r[b]|_F is rebuilt by every CTA of stage 1 from the per-tile partial
// sums of |r'|^2 that stage 2 of the previous application wrote (fixed summation
// order: deterministic, no atomics).
//
// CTA tile 64 x 64 complex, 8 warps as 4 (m) x 2 (n), warp tile 16 x 32 = 2 x 4 DMMA
// blocks, accumulators (re, im) in registers: 32 doubles per lane.
// Shared-memory slabs (KS = 16 complex per stage):
//     sA [64][KS+4]   row-major in k: a quarter-warp LDS.128 touches 2 rows x 64 B,
//                     row stride = 320 B = 64 mod 128  -> conflict-free
//     sB BFORM 0: [KS][64+2]  (k-major): quarter-warp touches 4 rows x 32 B,
//                     row stride = 1056 B = 32 mod 128 -> conflict-free
//        BFORM 1: [64][KS+4]  same shape as sA (B_s stored [n][k], conjugated on read)
#pragma once
#include <cuda_runtime.h>
#include "core.cuh"

namespace qmps {

constexpr int ZG_TM = 64, ZG_TN = 64, ZG_KS = 16;
constexpr int ZG_LDA = ZG_KS + 4;          // complex elements
constexpr int ZG_LDB0 = ZG_TN + 2;
constexpr int ZG_SA_ELEMS = ZG_TM * ZG_LDA;                                  // 1280
constexpr int ZG_SB_ELEMS = (ZG_KS * ZG_LDB0 > ZG_TN * ZG_LDA) ? ZG_KS * ZG_LDB0 : ZG_TN * ZG_LDA;  // 1280
constexpr int ZG_STAGE_ELEMS = ZG_SA_ELEMS + ZG_SB_ELEMS;
constexpr int ZG_STAGES = 2;
constexpr int ZG_SMEM_BYTES = ZG_STAGES * ZG_STAGE_ELEMS * 16;               // 81920

__device__ __forceinline__ void zg_cp16(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;            // src-size 0: nothing is read, 16 zero bytes are written
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void zg_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void zg_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct ZgParams {
  int M, N, K;              // C is M x N, inner dimension K (per summand)
  int nsum;                 // summands t = 0 .. nsum-1 (stage 2: the physical index s)
  const cx<double>* A;      // [batch*nsum][M][K] row-major
  const cx<double>* B;      // BFORM 0: [batch / b_div][K][N];  BFORM 1: [batch*nsum][N][K]
  int b_div;
  cx<double>* C;            // [batch][M][N]
  // alpha: C *= 1/sqrt(sum_j norm_in[(b / b_div) * n_in + j]), norm_in == nullptr -> 1
  const double* norm_in; int n_in;
  // norm_out[b * (tiles) + tile] = sum |C tile|^2 (after alpha), nullptr -> skip
  double* norm_out;
};

template <int BFORM>
__global__ void __launch_bounds__(256, 2)
zgemm_dmma_kernel(ZgParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cx<double>* smem = reinterpret_cast<cx<double>*>(smem_raw);
  __shared__ double s_red[8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 1, wn = warp & 1;             // 4 x 2 warps
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z;
  const int tm0 = blockIdx.y * ZG_TM, tn0 = blockIdx.x * ZG_TN;
  const int M = p.M, N = p.N, K = p.K;
  const int kslabs = (K + ZG_KS - 1) / ZG_KS;
  const int total = kslabs * p.nsum;

  double cre[2][4][2], cim[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { cre[i][j][0] = cre[i][j][1] = 0.0; cim[i][j][0] = cim[i][j][1] = 0.0; }

  auto issue = [&](int it, int stage) {
    const int ts = it / kslabs, k0 = (it - ts * kslabs) * ZG_KS;
    cx<double>* sA = smem + stage * ZG_STAGE_ELEMS;
    cx<double>* sB = sA + ZG_SA_ELEMS;
    const cx<double>* Ag = p.A + ((size_t)b * p.nsum + ts) * (size_t)M * K;
    // A slab: 64 rows x 16 k = 1024 elements, 4 per thread; consecutive threads -> consecutive k
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = q * 256 + tid, r = e >> 4, c = e & 15;
      const bool ok = (tm0 + r < M) && (k0 + c < K);
      zg_cp16(sA + r * ZG_LDA + c, ok ? Ag + (size_t)(tm0 + r) * K + k0 + c : Ag, ok);
    }
    if (BFORM == 0) {
      const cx<double>* Bg = p.B + (size_t)(b / p.b_div) * (size_t)K * N;
#pragma unroll
      for (int q = 0; q < 4; ++q) {                      // [k][n]: consecutive threads -> consecutive n
        const int e = q * 256 + tid, kk = e >> 6, c = e & 63;
        const bool ok = (k0 + kk < K) && (tn0 + c < N);
        zg_cp16(sB + kk * ZG_LDB0 + c, ok ? Bg + (size_t)(k0 + kk) * N + tn0 + c : Bg, ok);
      }
    } else {
      const cx<double>* Bg = p.B + ((size_t)b * p.nsum + ts) * (size_t)N * K;
#pragma unroll
      for (int q = 0; q < 4; ++q) {                      // [n][k]
        const int e = q * 256 + tid, r = e >> 4, c = e & 15;
        const bool ok = (tn0 + r < N) && (k0 + c < K);
        zg_cp16(sB + r * ZG_LDA + c, ok ? Bg + (size_t)(tn0 + r) * K + k0 + c : Bg, ok);
      }
    }
  };

  issue(0, 0);
  zg_commit();
  for (int it = 0; it < total; ++it) {
    const int stage = it & 1;
    if (it + 1 < total) issue(it + 1, stage ^ 1);
    zg_commit();
    zg_wait<1>();
    __syncthreads();
    const cx<double>* sA = smem + stage * ZG_STAGE_ELEMS;
    const cx<double>* sB = sA + ZG_SA_ELEMS;
#pragma unroll
    for (int kk = 0; kk < ZG_KS; kk += 4) {
      cx<double> af[2], bf[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) af[i] = sA[(wm * 16 + i * 8 + g) * ZG_LDA + kk + t];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (BFORM == 0) bf[j] = sB[(kk + t) * ZG_LDB0 + wn * 32 + j * 8 + g];
        else { bf[j] = sB[(wn * 32 + j * 8 + g) * ZG_LDA + kk + t]; bf[j].im = -bf[j].im; }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double nai = -af[i].im;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dmma884(cre[i][j][0], cre[i][j][1], af[i].re, bf[j].re);
          dmma884(cim[i][j][0], cim[i][j][1], af[i].re, bf[j].im);
          dmma884(cre[i][j][0], cre[i][j][1], nai, bf[j].im);
          dmma884(cim[i][j][0], cim[i][j][1], af[i].im, bf[j].re);
        }
      }
    }
    __syncthreads();
  }
  zg_wait<0>();

  // ---- epilogue: alpha, store, optional |C|^2 partial
  double alpha = 1.0;
  if (p.norm_in) {
    double s = 0.0;
    const double* ni = p.norm_in + (size_t)(b / p.b_div) * p.n_in;
    for (int j = 0; j < p.n_in; ++j) s += ni[j];
    alpha = 1.0 / sqrt(s);
  }
  double part = 0.0;
  cx<double>* Cg = p.C + (size_t)b * (size_t)M * N;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = tm0 + wm * 16 + i * 8 + g;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = tn0 + wn * 32 + j * 8 + 2 * t;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const double re = cre[i][j][h] * alpha, im = cim[i][j][h] * alpha;
        if (row < M && col + h < N) {
          Cg[(size_t)row * N + col + h] = mk<double>(re, im);
          part += re * re + im * im;
        }
      }
    }
  }
  if (p.norm_out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_red[warp] = part;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += s_red[w];
      p.norm_out[(size_t)b * (gridDim.x * gridDim.y) + blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
  }
}

// r[b] *= 1/sqrt(sum_j norm[b*n + j])   (final normalisation of the power method)
__global__ void __launch_bounds__(256)
zg_scale_kernel(int64_t len, cx<double>* __restrict__ r, const double* __restrict__ norm, int n) {
  double s = 0.0;
  for (int j = 0; j < n; ++j) s += norm[(size_t)blockIdx.x * n + j];
  const double a = 1.0 / sqrt(s);
  cx<double>* p = r + (size_t)blockIdx.x * len;
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) p[i] = p[i] * a;
}

}  // namespace qmps
