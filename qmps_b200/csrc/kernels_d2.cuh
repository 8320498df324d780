// qmps_b200 D = 2 kernels (sm_100a).  HBM-bound streaming kernels:
//   env_d2_stream_kernel   complex128: 128 B in, 16 + 64 (+64) B out per problem.
//   env_d2_simple_kernel   any dtype, direct per-thread I/O (complex64 mode).
//   energy_d2_theta_kernel theta -> U -> A -> r -> energy, all in registers, with the
//                          rotosolve shift fan-out.
//
// Data movement of the streaming kernel (per warp, 32 problems per tile):
//   global -> shared : cp.async 16 B per lane, 8 per lane per tile, each warp
//                      instruction a fully coalesced 512 B; the destination is
//                      XOR-swizzled on 16 B units so that the per-thread read-back
//                      of "my 128 B" is bank-conflict free; two stages in flight
//                      (next tile's loads are issued before this tile's math).
//   shared -> global : results are staged per warp, then written as coalesced
//                      16 B-per-lane streaming stores.
#pragma once
#include <cuda_runtime.h>
#include "d2.cuh"

namespace qmps {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void st_stream16(void* gdst, const cx<double>& v) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};\n" ::"l"(gdst), "d"(v.re), "d"(v.im));
}

constexpr int D2_WARPS = 8;                    // warps per CTA
constexpr int D2_STAGE_BYTES = 32 * 128;       // one input tile of a warp
constexpr int D2_OUT_BYTES = 32 * 64;          // one staged output (r or C) of a warp
constexpr int D2_WARP_BYTES = 2 * D2_STAGE_BYTES + D2_OUT_BYTES;
constexpr int D2_SMEM_BYTES = D2_WARPS * D2_WARP_BYTES;

template <bool IN_U>
__device__ __forceinline__ void d2_issue_tile(const cx<double>* __restrict__ in, int64_t N, int64_t tile,
                                              cx<double>* stage, int lane) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int unit = k * 32 + lane;
    const int pl = unit >> 3, e = unit & 7;     // local problem, entry within the 8 needed
    const int64_t gp = tile * 32 + pl;
    if (gp < N) {
      const cx<double>* src;
      int aidx;
      if (IN_U) {                               // U[row][col], row = 2i+s, needed cols 0,1
        const int row = e >> 1, j = e & 1;
        src = in + gp * 16 + row * 4 + j;
        aidx = (row & 1) * 4 + (row >> 1) * 2 + j;
      } else {
        src = in + gp * 8 + e;
        aidx = e;
      }
      cp_async16(stage + pl * 8 + (aidx ^ (pl & 7)), src);
    }
  }
}

// PACKED: one 64-byte record per problem through r_out instead of eta / r / C / status (see qmps_env_exact_packed):
//   [ r00, Re r01, Im r01, c00, Re c10, Im c10, c11, (double) status ]   (r11 = 1 - r00, r10 = conj r01, c01 = 0, eta = 1)
template <bool IN_U, bool WANT_C, bool PACKED = false>
__global__ void __launch_bounds__(D2_WARPS * 32, 2)
env_d2_stream_kernel(const cx<double>* __restrict__ in, int64_t N, cx<double>* __restrict__ eta_out,
                     cx<double>* __restrict__ r_out, cx<double>* __restrict__ C_out,
                     int32_t* __restrict__ status_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // programmatic dependent launch: let the next launch on the stream become resident while this
  // grid drains, and do not touch global memory before the previous grid has completed and
  // flushed (both are no-ops when the launch does not carry the PDL attribute)
  asm volatile("griddepcontrol.launch_dependents;\n" ::);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wb = smem_raw + warp * D2_WARP_BYTES;
  cx<double>* stage0 = reinterpret_cast<cx<double>*>(wb);
  cx<double>* stage1 = reinterpret_cast<cx<double>*>(wb + D2_STAGE_BYTES);
  cx<double>* ost = reinterpret_cast<cx<double>*>(wb + 2 * D2_STAGE_BYTES);

  const int64_t ntiles = (N + 31) >> 5;
  const int64_t stride = (int64_t)gridDim.x * D2_WARPS;
  int64_t tile = (int64_t)blockIdx.x * D2_WARPS + warp;
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
  if (tile < ntiles) d2_issue_tile<IN_U>(in, N, tile, stage0, lane);
  cp_async_commit();
  int buf = 0;
  for (; tile < ntiles; tile += stride, buf ^= 1) {
    cx<double>* cur = buf ? stage1 : stage0;
    cx<double>* nxt = buf ? stage0 : stage1;
    if (tile + stride < ntiles) d2_issue_tile<IN_U>(in, N, tile + stride, nxt, lane);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    cx<double> a[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = cur[lane * 8 + (e ^ (lane & 7))];
    cx<double> r[4], C[4];
    double eta;
    const int st = env_d2_solve<double, WANT_C>(a, r, &eta, C);
    const int64_t gp = tile * 32 + lane;
    if (PACKED) {
      const cx<double> P0 = mk<double>(r[0].re, r[1].re), P1 = mk<double>(r[1].im, C[0].re);
      const cx<double> P3 = mk<double>(C[3].re, (double)st);
      r[0] = P0; r[1] = P1; r[2] = C[2]; r[3] = P3;
    }
    if (!PACKED && gp < N) {
      if (eta_out) st_stream16(eta_out + gp, mk<double>(eta, 0.0));
      if (status_out) status_out[gp] = st;
    }
    if (r_out) {
#pragma unroll
      for (int e = 0; e < 4; ++e) ost[lane * 4 + (e ^ ((lane >> 1) & 3))] = r[e];
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int unit = k * 32 + lane, row = unit >> 2, e = unit & 3;
        const int64_t g2 = tile * 32 + row;
        if (g2 < N) st_stream16(r_out + g2 * 4 + e, ost[row * 4 + (e ^ ((row >> 1) & 3))]);
      }
      __syncwarp();
    }
    if (WANT_C && !PACKED) {
#pragma unroll
      for (int e = 0; e < 4; ++e) ost[lane * 4 + (e ^ ((lane >> 1) & 3))] = C[e];
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int unit = k * 32 + lane, row = unit >> 2, e = unit & 3;
        const int64_t g2 = tile * 32 + row;
        if (g2 < N) st_stream16(C_out + g2 * 4 + e, ost[row * 4 + (e ^ ((row >> 1) & 3))]);
      }
    }
    __syncwarp();   // everyone is done with `cur` before the next iteration refills it
  }
  cp_async_wait<0>();
}

// 16-byte vector I/O of one problem's tensors (8 entries in, 4 out): every sector a thread touches
// is used completely within one or two instructions instead of 8-byte accesses strided across the warp
__device__ __forceinline__ void d2_load8(const cx<float>* p, cx<float>* a) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(p) + q);
    a[2 * q] = mk<float>(v.x, v.y); a[2 * q + 1] = mk<float>(v.z, v.w);
  }
}
__device__ __forceinline__ void d2_load8(const cx<double>* p, cx<double>* a) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const double2 v = __ldcs(reinterpret_cast<const double2*>(p) + q);
    a[q] = mk<double>(v.x, v.y);
  }
}
__device__ __forceinline__ void d2_store4(cx<float>* p, const cx<float>* r) {
#pragma unroll
  for (int q = 0; q < 2; ++q)
    __stcs(reinterpret_cast<float4*>(p) + q, make_float4(r[2 * q].re, r[2 * q].im, r[2 * q + 1].re, r[2 * q + 1].im));
}
__device__ __forceinline__ void d2_store4(cx<double>* p, const cx<double>* r) {
#pragma unroll
  for (int q = 0; q < 4; ++q) __stcs(reinterpret_cast<double2*>(p) + q, make_double2(r[q].re, r[q].im));
}

// direct I/O variant (complex64 mode; also the reference point the streaming kernel
// is measured against)
template <typename T>
__global__ void __launch_bounds__(256)
env_d2_simple_kernel(const cx<T>* __restrict__ in, int64_t N, int in_is_U, cx<T>* __restrict__ eta_out,
                     cx<T>* __restrict__ r_out, cx<T>* __restrict__ C_out, int32_t* __restrict__ status_out) {
  for (int64_t gp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gp < N; gp += (int64_t)gridDim.x * blockDim.x) {
    cx<T> a[8];
    if (in_is_U) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int row = e >> 1, j = e & 1;
        a[(row & 1) * 4 + (row >> 1) * 2 + j] = in[gp * 16 + row * 4 + j];
      }
    } else {
      d2_load8(in + gp * 8, a);
    }
    cx<T> r[4], C[4];
    T eta;
    int st;
    if (C_out) st = env_d2_solve<T, true>(a, r, &eta, C); else st = env_d2_solve<T, false>(a, r, &eta, C);
    if (eta_out) eta_out[gp] = mk<T>(eta, 0);
    if (status_out) status_out[gp] = st;
    if (r_out) d2_store4(r_out + gp * 4, r);
    if (C_out) d2_store4(C_out + gp * 4, C);
  }
}

// theta -> energy for two-qubit ansaetze (D = 2).  Problem (n, s) evaluates
// theta[n] + shifts[s] e_coord.  ops/hmat/shifts live in constant-like global memory
// (uniform across the grid).
constexpr int D2_MAX_OPS = 256;
template <typename T>
__global__ void __launch_bounds__(128)
energy_d2_theta_kernel(const GateOp* __restrict__ ops, int nops, int64_t N, int P,
                       const double* __restrict__ theta, const cx<T>* __restrict__ hmat, int coord,
                       const double* __restrict__ shifts, int nshift, T* __restrict__ energy,
                       int32_t* __restrict__ status, cx<T>* __restrict__ A_out) {
  __shared__ GateOp s_ops[D2_MAX_OPS];
  __shared__ cx<T> s_h[16];
  for (int k = threadIdx.x; k < nops; k += blockDim.x) s_ops[k] = ops[k];
  if (hmat && threadIdx.x < 16) s_h[threadIdx.x] = hmat[threadIdx.x];
  __syncthreads();
  const int S = nshift > 0 ? nshift : 1;
  const int64_t total = N * S;
  for (int64_t pid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pid < total; pid += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = pid / S;
    const int s = (int)(pid - n * S);
    const double sh = nshift > 0 ? shifts[s] : 0.0;
    cx<T> x[8];                                      // rows 0..3 x cols 0..1 of U
    ansatz_reg2<T, 2>(s_ops, nops, theta + n * P, nshift > 0 ? coord : -1, sh, x);
    cx<T> a[8];                                      // A[s][i][j] = U[2i+s][j]
#pragma unroll
    for (int row = 0; row < 4; ++row)
#pragma unroll
      for (int j = 0; j < 2; ++j) a[(row & 1) * 4 + (row >> 1) * 2 + j] = x[row * 2 + j];
    if (A_out) {
#pragma unroll
      for (int e = 0; e < 8; ++e) A_out[pid * 8 + e] = a[e];
    }
    if (energy) {
      cx<T> r[4], C[4];
      T eta;
      const int st = env_d2_solve<T, true>(a, r, &eta, C);
      energy[pid] = energy_d2<T>(a, r, s_h);
      if (status) status[pid] = st;
    }
  }
}

}  // namespace qmps
