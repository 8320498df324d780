// qmps_b200 brick-wall kernel: one group of G lanes per (ket, bra) problem, per-problem scratch in
// shared memory, persistent grid-stride over the batch.  Modes (uniform over the launch):
//   BW_ENV      RightEnvironment / LeftEnvironment .exact_environment_circuit + .exact_environment
//   BW_APPLY    RightEnvironment.circuit  (one application of the right map to M)
//   BW_EXPECT   OverlapCalculator.expectation_value (2- or 4-qubit operator)
//   BW_OVERLAP  ManifoldOverlap.circuit with given Mr, Ml
//   BW_COST     Evolve.exact_cost_function body: environment of the mixed map -> overlap -> -|.|^2
// (new_tdvp/ClassicalTDVPStripped.py:228-533, 777-790).  HBM traffic per problem is the candidate
// unitaries in (2 x 16 complex) and one number out; W / O / the ket state are shared and stay in L1/L2.
#pragma once
#include <cuda_runtime.h>
#include "brickwall.cuh"
#include "kernels_generic.cuh"

namespace qmps {

enum { BW_ENV = 0, BW_APPLY = 1, BW_EXPECT = 2, BW_OVERLAP = 3, BW_COST = 4 };

struct BwParams {
  int mode, side, bra_undaggered, mbits;
  int64_t N;                 // problems
  int64_t NK, NB, NM, NW;    // ket pairs, bra pairs, (Mr, Ml) pairs, operators: each N or 1 (broadcast)
  const void* U1; const void* U2;     // [NK][4][4]
  const void* B1; const void* B2;     // [NB][4][4]
  const void* Mr; const void* Ml;     // [NM][2][2]
  const void* W;                      // [NW][16][16] (BW_EXPECT with mbits = 2: [NW][4][4])
  void* mat;                 // [N][4][4]   optional (BW_ENV)
  void* eta;                 // [N] complex optional (BW_ENV, BW_COST)
  void* vec;                 // [N][2][2]   optional (BW_ENV, BW_COST); BW_APPLY: the result
  void* overlap;             // [N] complex optional (BW_OVERLAP, BW_COST)
  void* real_out;            // [N] real    BW_EXPECT: <O>;  BW_COST: -|overlap|^2
  int32_t* status;           // [N] optional
};

template <typename T, int G>
__global__ void __launch_bounds__(128)
bw_kernel(BwParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  int gi, gpc;
  const Grp g = make_group<G>(&gi, &gpc);
  const size_t wb = bw_work_bytes<T>(G);
  const BwWork<T> W = bw_carve<T>(smem_raw + (size_t)gi * wb, G);
  typedef cx<T> Z;
  const int64_t stride = (int64_t)gridDim.x * gpc;
  // every group of a warp runs the same trip count (sub-warp masks keep the groups independent,
  // but the loop bound must not depend on the group when G < 32 shares a warp's control flow)
  for (int64_t pid0 = (int64_t)blockIdx.x * gpc; pid0 < p.N; pid0 += stride) {
    int64_t pid = pid0 + gi;
    const bool live = pid < p.N;
    if (!live) pid = p.N - 1;
    const Z* U1 = reinterpret_cast<const Z*>(p.U1) + (p.NK == 1 ? 0 : pid) * 16;
    const Z* U2 = reinterpret_cast<const Z*>(p.U2) + (p.NK == 1 ? 0 : pid) * 16;
    const Z* B1 = p.B1 ? reinterpret_cast<const Z*>(p.B1) + (p.NB == 1 ? 0 : pid) * 16 : nullptr;
    const Z* B2 = p.B2 ? reinterpret_cast<const Z*>(p.B2) + (p.NB == 1 ? 0 : pid) * 16 : nullptr;
    const int wsz = (p.mode == BW_EXPECT && p.mbits == 2) ? 16 : 256;
    const Z* Wop = p.W ? reinterpret_cast<const Z*>(p.W) + (p.NW == 1 ? 0 : pid) * wsz : nullptr;
    bw_load<T>(g, U1, U2, B1, B2, p.bra_undaggered, W);
    int status = ST_OK;
    if (p.mode == BW_ENV) {
      if (p.mat) {
        bw_env_matrix<T>(g, W, p.side, W.E, 5);
        g.sync();
        if (live) { Z* o = reinterpret_cast<Z*>(p.mat) + pid * 16; for (int e = g.lane; e < 16; e += g.size) o[e] = W.E[(e >> 2) * 5 + (e & 3)]; }
        g.sync();
      }
      Z lam;
      status = bw_exact_environment<T>(g, W, p.side, &lam, p.vec != nullptr);
      if (live) {
        if (g.lane == 0 && p.eta) reinterpret_cast<Z*>(p.eta)[pid] = lam;
        if (p.vec) { Z* o = reinterpret_cast<Z*>(p.vec) + pid * 4; for (int e = g.lane; e < 4; e += g.size) o[e] = W.x[e]; }
      }
    } else if (p.mode == BW_APPLY) {
      const Z* M = reinterpret_cast<const Z*>(p.Mr) + (p.NM == 1 ? 0 : pid) * 4;
      for (int e = g.lane; e < 4; e += g.size) W.mr[e] = M[e];
      g.sync();
      bw_env_apply<T>(g, W, W.mr, W.x);
      g.sync();
      if (live) { Z* o = reinterpret_cast<Z*>(p.vec) + pid * 4; for (int e = g.lane; e < 4; e += g.size) o[e] = W.x[e]; }
    } else if (p.mode == BW_EXPECT) {
      const T v = bw_expectation<T>(g, W, Wop, p.mbits);
      if (live && g.lane == 0) reinterpret_cast<T*>(p.real_out)[pid] = v;
    } else {
      if (p.mode == BW_COST) {
        Z lam;
        status = bw_exact_environment<T>(g, W, 0, &lam, 1);
        for (int e = g.lane; e < 4; e += g.size) {
          W.mr[e] = W.x[e];
          W.ml[e] = conj(W.x[(e & 1) * 2 + (e >> 1)]);            // Ml = Mr^dagger (:783-786)
        }
        if (live) {
          if (g.lane == 0 && p.eta) reinterpret_cast<Z*>(p.eta)[pid] = lam;
          if (p.vec) { Z* o = reinterpret_cast<Z*>(p.vec) + pid * 4; for (int e = g.lane; e < 4; e += g.size) o[e] = W.x[e]; }
        }
      } else {
        const Z* Mr = reinterpret_cast<const Z*>(p.Mr) + (p.NM == 1 ? 0 : pid) * 4;
        const Z* Ml = reinterpret_cast<const Z*>(p.Ml) + (p.NM == 1 ? 0 : pid) * 4;
        for (int e = g.lane; e < 4; e += g.size) { W.mr[e] = Mr[e]; W.ml[e] = Ml[e]; }
      }
      g.sync();
      const Z ov = bw_overlap<T>(g, W, Wop);
      if (live && g.lane == 0) {
        if (p.overlap) reinterpret_cast<Z*>(p.overlap)[pid] = ov;
        if (p.real_out) reinterpret_cast<T*>(p.real_out)[pid] = -(ov.re * ov.re + ov.im * ov.im);
      }
    }
    if (live && g.lane == 0 && p.status) p.status[pid] = status;
    g.sync();
  }
}

// chi = (1 (x) W (x) 1)(1 (x) U1 (x) U1 (x) 1)(U2 (x) U2 (x) U2)|0..0>: once per launch of the thread-per-candidate
// cost kernel (one warp, 64 amplitudes)
template <typename T>
__global__ void __launch_bounds__(32)
bw_chi_kernel(const cx<T>* __restrict__ U1, const cx<T>* __restrict__ U2, const cx<T>* __restrict__ W,
              cx<T>* __restrict__ chi) {
  __shared__ cx<T> k1[16], k2[4], psi[64], tmp[64];
  Grp g; g.lane = threadIdx.x; g.size = 32; g.mask = 0xffffffffu; g.cta = 0;
  for (int e = g.lane; e < 16; e += 32) k1[e] = U1[e];
  for (int q = g.lane; q < 4; q += 32) k2[q] = U2[q * 4];
  g.sync();
  bw_build_state<T>(g, k2, k1, 3, 0, psi, tmp);
  bw_apply_mid<T>(g, W, 4, 6, psi, tmp);
  g.sync();
  for (int e = g.lane; e < 64; e += 32) chi[e] = tmp[e];
}

// Evolve.exact_cost_function for a population of candidates against ONE ket state and ONE W: a thread per
// candidate, registers only (brickwall.cuh::bw_cost_thread); the shared ket data sit in shared memory and
// are read as warp-wide broadcasts.  HBM traffic: the candidate's two unitaries in, one number out.
template <typename T>
__global__ void __launch_bounds__(128)
bw_cost_thread_kernel(BwParams p, const cx<T>* __restrict__ chi_g) {
  __shared__ cx<T> k1[16], k2[4], chi[64];
  for (int e = threadIdx.x; e < 16; e += blockDim.x) k1[e] = reinterpret_cast<const cx<T>*>(p.U1)[e];
  for (int q = threadIdx.x; q < 4; q += blockDim.x) k2[q] = reinterpret_cast<const cx<T>*>(p.U2)[q * 4];
  for (int e = threadIdx.x; e < 64; e += blockDim.x) chi[e] = chi_g[e];
  __syncthreads();
  typedef cx<T> Z;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t pid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pid < p.N; pid += stride) {
    const Z* V1 = reinterpret_cast<const Z*>(p.B1) + (p.NB == 1 ? 0 : pid) * 16;
    const Z* V2 = reinterpret_cast<const Z*>(p.B2) + (p.NB == 1 ? 0 : pid) * 16;
    Z ov, lam, mr[4];
    const int status = bw_cost_thread<T>(k1, k2, chi, V1, V2, &ov, &lam, mr);
    if (p.real_out) reinterpret_cast<T*>(p.real_out)[pid] = -(ov.re * ov.re + ov.im * ov.im);
    if (p.overlap) reinterpret_cast<Z*>(p.overlap)[pid] = ov;
    if (p.eta) reinterpret_cast<Z*>(p.eta)[pid] = lam;
    if (p.vec) {
      Z* o = reinterpret_cast<Z*>(p.vec) + pid * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = mr[i];
    }
    if (p.status) p.status[pid] = status;
  }
}

// exact_environment(_circuit) with a thread per problem (brickwall.cuh::bw_env_thread), any broadcast shape
template <typename T>
__global__ void __launch_bounds__(128)
bw_env_thread_kernel(BwParams p) {
  typedef cx<T> Z;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t pid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pid < p.N; pid += stride) {
    const Z* U1 = reinterpret_cast<const Z*>(p.U1) + (p.NK == 1 ? 0 : pid) * 16;
    const Z* U2 = reinterpret_cast<const Z*>(p.U2) + (p.NK == 1 ? 0 : pid) * 16;
    const Z* B1 = reinterpret_cast<const Z*>(p.B1) + (p.NB == 1 ? 0 : pid) * 16;
    const Z* B2 = reinterpret_cast<const Z*>(p.B2) + (p.NB == 1 ? 0 : pid) * 16;
    Z lam, x[4];
    const int status = bw_env_thread<T>(U1, U2, B1, B2, p.bra_undaggered, p.side,
                                        p.mat ? reinterpret_cast<Z*>(p.mat) + pid * 16 : nullptr, &lam,
                                        p.vec ? x : nullptr);
    if (p.eta) reinterpret_cast<Z*>(p.eta)[pid] = lam;
    if (p.vec) {
      Z* o = reinterpret_cast<Z*>(p.vec) + pid * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = x[i];
    }
    if (p.status) p.status[pid] = status;
  }
}

}  // namespace qmps
