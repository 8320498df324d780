// qmps_b200 generic-D kernels: one group of G lanes per problem
//   G <= 32 : sub-warp groups, several problems per warp, __syncwarp(mask)
//   G  > 32 : the CTA is the group (blockDim.x == G), __syncthreads()
// Grid-stride (persistent) over problems; grid sized by the host as a multiple of
// the SM count.
#pragma once
#include <cuda_runtime.h>
#include "generic.cuh"

namespace qmps {

template <int G> __device__ __forceinline__ Grp make_group(int* group_in_cta, int* groups_per_cta) {
  Grp g;
  if (G > 32) {
    g.lane = threadIdx.x; g.size = blockDim.x; g.mask = 0xffffffffu; g.cta = 1;
    *group_in_cta = 0; *groups_per_cta = 1;
  } else {
    const int l = threadIdx.x & 31, sub = l / G;
    g.lane = l % G; g.size = G; g.cta = 0;
    g.mask = (G >= 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (sub * G));
    *group_in_cta = threadIdx.x / G; *groups_per_cta = blockDim.x / G;
  }
  return g;
}

// bump allocator over a byte range; identical call sequence on host (sizing) and device
struct Bump {
  size_t off;
  QMPS_HD Bump() : off(0) {}
  QMPS_HD size_t take(size_t bytes) { size_t o = off; off = (off + bytes + 15) & ~size_t(15); return o; }
};

template <typename T> struct EnvLayout {
  size_t A, E, x, C, tmp, step, done, red, trig, total;
};
// e_in_smem = 0: the n x (n+1) matrix lives in a global workspace
template <typename T>
QMPS_HD EnvLayout<T> env_layout(int d, int D, int nops, int G, int e_in_smem, int want_tmp) {
  EnvLayout<T> L;
  Bump b;
  const int n = D * D;
  L.A = b.take(sizeof(cx<T>) * (size_t)d * D * D);                 // also holds the 2D x D ansatz state
  L.E = b.take(e_in_smem ? sizeof(cx<T>) * (size_t)n * (n + 1) : 0);
  L.x = b.take(sizeof(cx<T>) * n);
  L.C = b.take(sizeof(cx<T>) * n);
  L.tmp = b.take(want_tmp ? sizeof(cx<T>) * (size_t)8 * n : 0);   // energy blocks / QR scratch
  L.step = b.take(sizeof(int) * n);
  L.done = b.take(sizeof(int) * n);
  L.red = b.take(sizeof(T) * G);
  L.trig = b.take(sizeof(T) * 2 * (nops > 0 ? nops : 1));
  L.total = b.off;
  return L;
}

struct EnvParams {
  int d, D;
  int64_t N;
  const void* in;        // tensors/unitaries, or nullptr when theta-driven
  int in_is_U;
  int assume_lc;
  void* eta; void* r; void* C; int32_t* status;
  void* ws; size_t ws_stride;   // global workspace per CTA (elements) when E is not in smem
  // theta-driven front-end (energy kernels)
  const GateOp* ops; int nops; int nq; int P; const double* theta;
  int coord; const double* shifts; int nshift;
  // energy
  const void* hmat; void* energy; int two_site;
};

// mode 0: environment outputs (eta, r, C, status); mode 1: energy
template <typename T, int G, int MODE>
__global__ void __launch_bounds__(G > 32 ? G : 128)
env_generic_kernel(EnvParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  int gi, gpc;
  const Grp g = make_group<G>(&gi, &gpc);
  const int D = p.D, d = p.d, n = D * D;
  const int e_in_smem = (p.ws == nullptr);
  const EnvLayout<T> L = env_layout<T>(d, D, p.nops, G > 32 ? (int)blockDim.x : G, e_in_smem,
                                       (MODE == 1) || !p.assume_lc);
  unsigned char* base = smem_raw + (size_t)gi * L.total;
  cx<T>* A = reinterpret_cast<cx<T>*>(base + L.A);
  cx<T>* E = e_in_smem ? reinterpret_cast<cx<T>*>(base + L.E)
                       : reinterpret_cast<cx<T>*>(p.ws) + (size_t)blockIdx.x * p.ws_stride;
  cx<T>* x = reinterpret_cast<cx<T>*>(base + L.x);
  cx<T>* Cc = reinterpret_cast<cx<T>*>(base + L.C);
  cx<T>* tmp = reinterpret_cast<cx<T>*>(base + L.tmp);
  int* step_row = reinterpret_cast<int*>(base + L.step);
  int* done = reinterpret_cast<int*>(base + L.done);
  T* red = reinterpret_cast<T*>(base + L.red);
  T* trig = reinterpret_cast<T*>(base + L.trig);
  const int ld = n + 1;
  const int S = (MODE == 1 && p.nshift > 0) ? p.nshift : 1;
  const int64_t total = p.N * S;
  const cx<T>* hmat = reinterpret_cast<const cx<T>*>(p.hmat);

  for (int64_t pid = (int64_t)blockIdx.x * gpc + gi; pid < total; pid += (int64_t)gridDim.x * gpc) {
    const int64_t pn = pid / S;
    const int sidx = (int)(pid - pn * S);
    // ---- 1. the tensor A[d][D][D] into shared memory
    if (p.theta) {
      StateLayout SL; SL.R = 2 * D; SL.ncols = D; SL.a_layout = 1;
      const double sh = p.nshift > 0 ? p.shifts[sidx] : 0.0;
      ansatz_eval<T>(g, p.ops, p.nops, p.theta + pn * p.P, p.nshift > 0 ? p.coord : -1, sh, p.nq, SL, A, trig);
    } else if (p.in_is_U) {
      const cx<T>* U = reinterpret_cast<const cx<T>*>(p.in) + pn * (size_t)(4 * n);
      for (int e = g.lane; e < 2 * n; e += g.size) {       // A[s][i][j] = U[2i+s][j]
        int s = e / n, ij = e - s * n, i = ij / D, j = ij - i * D;
        A[e] = U[(2 * i + s) * (2 * D) + j];
      }
    } else {
      const cx<T>* src = reinterpret_cast<const cx<T>*>(p.in) + pn * (size_t)(d * n);
      for (int e = g.lane; e < d * n; e += g.size) A[e] = src[e];
    }
    g.sync();
    // ---- 2. fixed point
    int status = ST_OK;
    cx<T> eta = mk<T>(1, 0);
    if (p.assume_lc) {
      T eta_r;
      status = env_solve_direct<T>(g, A, d, D, E, ld, x, step_row, done, red, &eta_r);
      eta = mk<T>(eta_r, 0);
    } else {
      // general tensor: leading eigenpair of E_AA, rotate to Hermitian, trace 1
      // (scratch vectors w, vv, rc, rs, rn are carved out of the tmp / C areas)
      cx<T>* w = Cc;                 // n entries
      cx<T>* vv = tmp;               // MODE 1 always has tmp; MODE 0 general path sizes it too
      cx<T>* rc = tmp + n;
      cx<T>* rs = tmp + 2 * n;
      T* rn = reinterpret_cast<T*>(tmp + 3 * n);
      status = leading_eigenpair<T>(g, A, A, d, D, 0, 1, E, ld, w, vv, rc, rs, rn, x, step_row, done, &eta);
      hermitise<T>(g, x, D);
      T tr = T(0);
      for (int i = 0; i < D; ++i) tr += x[i * D + i].re;
      T itr = T(1) / tr;
      g.sync();
      for (int e = g.lane; e < n; e += g.size) x[e] = x[e] * itr;
      g.sync();
    }
    if (MODE == 0) {
      if (p.C || p.status) {
        int bad = cholesky_lower<T>(g, x, D, Cc, D, D);
        if (bad && status == ST_OK) status = ST_NOT_PD;
      }
      if (g.lane == 0) {
        if (p.eta) reinterpret_cast<cx<T>*>(p.eta)[pid] = eta;
        if (p.status) p.status[pid] = status;
      }
      if (p.r) { cx<T>* o = reinterpret_cast<cx<T>*>(p.r) + pid * (size_t)n; for (int e = g.lane; e < n; e += g.size) o[e] = x[e]; }
      if (p.C) { cx<T>* o = reinterpret_cast<cx<T>*>(p.C) + pid * (size_t)n; for (int e = g.lane; e < n; e += g.size) o[e] = Cc[e]; }
    } else {
      // ---- 3. energy
      const cx<T>* M;
      if (p.two_site) M = A;                      // input already is the two-site block (d = 4)
      else { merge_block<T>(g, A, A, 2, 2, D, tmp); M = tmp; }
      T e = energy_from_block<T>(g, M, x, D, hmat, tmp + 4 * n, red);
      if (p.status) {
        int bad = cholesky_lower<T>(g, x, D, Cc, D, D);
        if (bad && status == ST_OK) status = ST_NOT_PD;
      }
      if (g.lane == 0) {
        reinterpret_cast<T*>(p.energy)[pid] = e;
        if (p.status) p.status[pid] = status;
      }
    }
    g.sync();
  }
}

// ---- mixed fixed points ---------------------------------------------------------------
template <typename T> struct FpLayout { size_t H, w, vv, rc, rs, rn, x, step, done, total; };
template <typename T> QMPS_HD FpLayout<T> fp_layout(int D, int h_in_smem) {
  FpLayout<T> L;
  Bump b;
  const int n = D * D;
  // Scratch with disjoint lifetimes shares storage (5.2 KB instead of 5.9 KB per 16 x 16 complex128
  // problem: 5 instead of 4 CTAs per SM): vv is live in hessenberg only, rc / rs / rn in the QR sweeps,
  // x / step / done in the inverse iteration that follows, and the leading eigenvalue is copied out
  // of w before that.
  L.H = b.take(h_in_smem ? sizeof(cx<T>) * (size_t)n * (n + 1) : 0);
  L.w = b.take(sizeof(cx<T>) * n);
  L.rc = b.take(sizeof(cx<T>) * n);
  L.rs = b.take(sizeof(cx<T>) * n);
  L.rn = b.take(sizeof(T) * n);
  L.vv = L.rs;
  L.x = L.rc;
  L.step = L.rn;      // n ints <= n reals
  L.done = L.w;       // n ints <= n complex
  L.total = b.off;
  return L;
}

struct FpParams {
  int d, D;
  int64_t NA, NB, N;
  const void* A; const void* B;
  int pair_mode, left;
  void* eta; void* vec; void* cost; void* echo; void* fid; int32_t* status;
  void* ws; size_t ws_stride;
  int vec_gauge;                 // QMPS_GAUGE_TRACE (0) / QMPS_GAUGE_ZGEEV (1)
  int64_t pid_offset, n_chunk;   // kernels_fp64p.cuh: the slice of the batch one launch pair works on (set by its launcher)
};

template <typename T, int G>
__global__ void __launch_bounds__(G > 32 ? G : 128)
fixed_point_kernel(FpParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  int gi, gpc;
  const Grp g = make_group<G>(&gi, &gpc);
  const int D = p.D, d = p.d, n = D * D;
  const int h_in_smem = (p.ws == nullptr);
  const FpLayout<T> L = fp_layout<T>(D, h_in_smem);
  unsigned char* base = smem_raw + (size_t)gi * L.total;
  cx<T>* H = h_in_smem ? reinterpret_cast<cx<T>*>(base + L.H)
                       : reinterpret_cast<cx<T>*>(p.ws) + (size_t)blockIdx.x * p.ws_stride;
  cx<T>* w = reinterpret_cast<cx<T>*>(base + L.w);
  cx<T>* vv = reinterpret_cast<cx<T>*>(base + L.vv);
  cx<T>* rc = reinterpret_cast<cx<T>*>(base + L.rc);
  cx<T>* rs = reinterpret_cast<cx<T>*>(base + L.rs);
  T* rn = reinterpret_cast<T*>(base + L.rn);
  cx<T>* x = reinterpret_cast<cx<T>*>(base + L.x);
  int* step_row = reinterpret_cast<int*>(base + L.step);
  int* done = reinterpret_cast<int*>(base + L.done);
  const size_t tsz = (size_t)d * n;
  for (int64_t pid = (int64_t)blockIdx.x * gpc + gi; pid < p.N; pid += (int64_t)gridDim.x * gpc) {
    int64_t ia, ib;
    if (p.pair_mode == 1) { ia = pid / p.NB; ib = pid - ia * p.NB; }
    else if (p.pair_mode == 2) { ib = pid / p.NA; ia = pid - ib * p.NA; }
    else { ia = pid < p.NA ? pid : p.NA - 1; ib = pid < p.NB ? pid : p.NB - 1; }
    const cx<T>* A = reinterpret_cast<const cx<T>*>(p.A) + ia * tsz;
    const cx<T>* B = reinterpret_cast<const cx<T>*>(p.B) + ib * tsz;
    cx<T> lam;
    const int status = leading_eigenpair<T>(g, A, B, d, D, p.left, p.vec != nullptr, H, n + 1, w, vv, rc, rs,
                                            rn, x, step_row, done, &lam, p.vec_gauge);
    if (g.lane == 0) {
      const T a2 = norm2(lam);
      if (p.eta) reinterpret_cast<cx<T>*>(p.eta)[pid] = lam;
      if (p.cost) reinterpret_cast<T*>(p.cost)[pid] = -sqrt(sqrt(a2));
      if (p.echo) reinterpret_cast<T*>(p.echo)[pid] = -log(a2);
      if (p.fid) reinterpret_cast<T*>(p.fid)[pid] = a2;
      if (p.status) p.status[pid] = status;
    }
    if (p.vec) {
      cx<T>* o = reinterpret_cast<cx<T>*>(p.vec) + pid * (size_t)n;
      for (int e = g.lane; e < n; e += g.size) o[e] = x[e];
    }
    g.sync();
  }
}

// ---- ansatz only: theta -> A or U ---------------------------------------------------------
template <typename T, int G>
__global__ void __launch_bounds__(G > 32 ? G : 128)
ansatz_kernel(const GateOp* __restrict__ ops, int nops, int nq, int64_t N, int P,
              const double* __restrict__ theta, int full_unitary, cx<T>* __restrict__ out,
              int coord, const double* __restrict__ shifts, int nshift) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  int gi, gpc;
  const Grp g = make_group<G>(&gi, &gpc);
  const int R = 1 << nq, D = R >> 1;
  const int nc = full_unitary ? R : D;
  const size_t per = (((size_t)R * nc * sizeof(cx<T>) + 15) & ~size_t(15)) + (((size_t)2 * nops * sizeof(T) + 15) & ~size_t(15));
  unsigned char* base = smem_raw + (size_t)gi * per;
  cx<T>* S = reinterpret_cast<cx<T>*>(base);
  T* trig = reinterpret_cast<T*>(base + (((size_t)R * nc * sizeof(cx<T>) + 15) & ~size_t(15)));
  StateLayout SL; SL.R = R; SL.ncols = nc; SL.a_layout = full_unitary ? 0 : 1;
  // shift fan-out (rotosolve): output (n, s) is the tensor of theta[n] + shifts[s] e_coord, index n * nshift + s
  const int SF = nshift > 0 ? nshift : 1;
  for (int64_t pid = (int64_t)blockIdx.x * gpc + gi; pid < N * SF; pid += (int64_t)gridDim.x * gpc) {
    const int64_t pn = pid / SF;
    const double sh = nshift > 0 ? shifts[(int)(pid - pn * SF)] : 0.0;
    ansatz_eval<T>(g, ops, nops, theta + pn * P, nshift > 0 ? coord : -1, sh, nq, SL, S, trig);
    cx<T>* o = out + pid * (size_t)(R * nc);
    for (int e = g.lane; e < R * nc; e += g.size) o[e] = S[e];
    g.sync();
  }
}

}  // namespace qmps
