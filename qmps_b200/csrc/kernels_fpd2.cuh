// qmps_b200 D = 2 mixed fixed points (eigenvalue, optionally the eigenvector): one THREAD per (A, B) pair, the 4 x 4 map and
// its Hessenberg + QR iteration in registers (fp_d2.cuh).  This is the Loschmidt / TDVP-step cost
// kernel of the D = 2 scripts (qmps/loschmidts/time_evo.py:75-116 = scripts/loschmidt.py:209-239,
// qmps/time_evolve_tools.py:84-91): in the outer-product mode consecutive threads take consecutive
// B tensors against one A (a time step), so A is a warp-wide broadcast from L1 and every 16-byte
// load of B is used; the output is one coalesced number per thread.
#pragma once
#include <cuda_runtime.h>
#include "fp_d2.cuh"
#include "kernels_generic.cuh"

namespace qmps {

template <typename T> struct vec2_of;
template <> struct vec2_of<double> { typedef double2 type; };
template <> struct vec2_of<float> { typedef float2 type; };

template <typename T> __device__ __forceinline__ cx<T> ld_cx(const cx<T>* p) {
  const typename vec2_of<T>::type v = __ldg(reinterpret_cast<const typename vec2_of<T>::type*>(p));
  return mk<T>(v.x, v.y);
}

// VEC: also the eigenvector (a separate instantiation: the eigenvalue-only Loschmidt path keeps its
// 112 registers)
template <typename T, bool VEC>
__global__ void __launch_bounds__(128)
fp_d2_kernel(FpParams p) {
  const int d = p.d;
  const size_t tsz = (size_t)d * 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t pid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pid < p.N; pid += stride) {
    int64_t ia, ib;
    if (p.pair_mode == 1) { ia = pid / p.NB; ib = pid - ia * p.NB; }
    else if (p.pair_mode == 2) { ib = pid / p.NA; ia = pid - ib * p.NA; }
    else { ia = pid < p.NA ? pid : p.NA - 1; ib = pid < p.NB ? pid : p.NB - 1; }
    const cx<T>* A = reinterpret_cast<const cx<T>*>(p.A) + ia * tsz;
    const cx<T>* B = reinterpret_cast<const cx<T>*>(p.B) + ib * tsz;
    cx<T> E[4][4];
    auto build = [&]() {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) E[r][c] = mk<T>(0, 0);
      for (int s = 0; s < d; ++s) {
        cx<T> a[4], b[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { a[q] = ld_cx<T>(A + s * 4 + q); b[q] = ld_cx<T>(B + s * 4 + q); }
        fpd2_accumulate<T>(E, a, b, p.left);
      }
    };
    build();
    cx<T> lam;
    const int status = fpd2_leading_of<T>(E, &lam);
    if (VEC && p.vec) {                            // one inverse iteration on the rebuilt map (operands are in L1)
      build();
      cx<T> x[4];
      fpd2_inverse_iteration<T>(E, lam, x);
      fpd2_fix_gauge<T>(x, p.vec_gauge);
      cx<T>* o = reinterpret_cast<cx<T>*>(p.vec) + pid * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = x[i];
    }
    const T a2 = norm2(lam);
    if (p.eta) reinterpret_cast<cx<T>*>(p.eta)[pid] = lam;
    if (p.cost) reinterpret_cast<T*>(p.cost)[pid] = -sqrt(sqrt(a2));
    if (p.echo) reinterpret_cast<T*>(p.echo)[pid] = -log(a2);
    if (p.fid) reinterpret_cast<T*>(p.fid)[pid] = a2;
    if (p.status) p.status[pid] = status;
  }
}

}  // namespace qmps
