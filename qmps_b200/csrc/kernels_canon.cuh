// qmps_b200 canonical-form / expectation-value kernels: one warp per problem,
// tensors staged in shared memory; D <= 16, d <= 4.  The eigen-solve that produces the fixed point
// is the expensive part (fixed_point_kernel / env kernels); these kernels are O(d D^3) per problem
// and bound by reading A and writing A'.  Algorithms: canon.cuh.
#pragma once
#include <cuda_runtime.h>
#include "kernels_generic.cuh"
#include "canon.cuh"

namespace qmps {

struct GaugeParams {
  int d, D;
  int64_t N;
  const void* A;        // [N][d][D][D]
  const void* X;        // [N][D][D]
  const void* eta;      // [N] complex, optional: output scaled by 1/sqrt|eta|
  int x_kind;           // 0: X = l Hermitian PD -> A' = L A L^-1;  1: X = C lower triangular -> A' = C^-1 A C
  void* A_out;          // [N][d][D][D]
  void* G_out;          // [N][D][D] optional (x_kind 0: L with tr(L^dagger L) = D)
  int32_t* status;      // optional
  int keep_status;      // 1: only overwrite status[n] when this kernel has something to report
};

constexpr int CANON_WARPS = 4;

template <typename T> QMPS_HD size_t gauge_smem_per_warp(int d, int D) {
  return sizeof(cx<T>) * (size_t)(d * D * D + 3 * D * D);
}
template <typename T> QMPS_HD size_t expect_smem_per_warp(int d, int D) {
  return sizeof(cx<T>) * (size_t)(2 * d * D * D + 2 * D * D + d * d * D + d * d);
}

template <typename T>
__global__ void __launch_bounds__(CANON_WARPS * 32)
gauge_kernel(GaugeParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  Grp g; g.lane = threadIdx.x & 31; g.size = 32; g.mask = 0xffffffffu; g.cta = 0;
  const int d = p.d, D = p.D, DD = D * D;
  cx<T>* sA = reinterpret_cast<cx<T>*>(smem_raw + (size_t)warp * gauge_smem_per_warp<T>(d, D));
  cx<T>* sG = sA + d * DD;
  cx<T>* sGi = sG + DD;
  cx<T>* sT = sGi + DD;
  const int64_t wstride = (int64_t)gridDim.x * CANON_WARPS;
  for (int64_t n = (int64_t)blockIdx.x * CANON_WARPS + warp; n < p.N; n += wstride) {
    T scale = T(1);
    if (p.eta) {
      const T m = cabs(((const cx<T>*)p.eta)[n]);
      scale = m > T(0) ? rsqrt_hd(m) : T(1);
    }
    const int st = gauge_problem<T>(g, (const cx<T>*)p.A + n * (size_t)(d * DD), (const cx<T>*)p.X + n * (size_t)DD,
                                    p.x_kind, scale, d, D, sA, sG, sGi, sT, (cx<T>*)p.A_out + n * (size_t)(d * DD),
                                    p.G_out ? (cx<T>*)p.G_out + n * (size_t)DD : nullptr);
    if (p.status && g.lane == 0 && (st != ST_OK || !p.keep_status)) p.status[n] = st;
  }
}

struct ExpectParams {
  int d, D, nops;
  int64_t N;
  const void* A;        // [N][d][D][D]
  const void* r;        // [N][D][D]
  const void* lvec;     // [N][D][D] optional
  const void* eta;      // [N] optional
  const void* ops;      // [nops][d][d]
  void* out;            // [N][nops] complex
};

template <typename T>
__global__ void __launch_bounds__(CANON_WARPS * 32)
expect_kernel(ExpectParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  Grp g; g.lane = threadIdx.x & 31; g.size = 32; g.mask = 0xffffffffu; g.cta = 0;
  const int d = p.d, D = p.D, DD = D * D;
  cx<T>* sA = reinterpret_cast<cx<T>*>(smem_raw + (size_t)warp * expect_smem_per_warp<T>(d, D));
  cx<T>* sR = sA + d * DD;
  cx<T>* sL = sR + DD;
  cx<T>* sP = sL + DD;
  cx<T>* sQ = sP + d * DD;
  cx<T>* sM = sQ + d * d * D;
  const int64_t wstride = (int64_t)gridDim.x * CANON_WARPS;
  for (int64_t n = (int64_t)blockIdx.x * CANON_WARPS + warp; n < p.N; n += wstride) {
    expect_problem<T>(g, (const cx<T>*)p.A + n * (size_t)(d * DD), (const cx<T>*)p.r + n * (size_t)DD,
                      p.lvec ? (const cx<T>*)p.lvec + n * (size_t)DD : nullptr,
                      p.eta ? (const cx<T>*)p.eta + n : nullptr, (const cx<T>*)p.ops, p.nops, d, D, sA, sR, sL, sP, sQ,
                      sM, (cx<T>*)p.out + n * (size_t)p.nops);
  }
}

}  // namespace qmps
