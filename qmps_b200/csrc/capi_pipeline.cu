// qmps_b200 C ABI, single-call pipelines (SURVEY 8(b) proposal): the whole Loschmidt / TDVP-step grid, whole
// rotosolve sweeps, their host-buffer forms, and the cross-rank argmin over NCCL.  Everything here is a
// composition of the entries of capi.cu on ONE stream -- a C caller gets in one call what
// qmps_b200/batched.py composes in Python.
#include <dlfcn.h>
#include <math.h>
#include <mutex>

#include "api_common.cuh"

using namespace qmps;
using namespace qmps_host;

namespace {

size_t csize(int dtype) { return dtype == QMPS_C128 ? 16 : 8; }
size_t rsize(int dtype) { return dtype == QMPS_C128 ? 8 : 4; }

// per-device stream of the host-buffer entries
int host_stream(int device, cudaStream_t* out) {
  static std::mutex mu;
  static cudaStream_t st[64] = {nullptr};
  if (device < 0 || device >= 64) return fail(QMPS_ERR_ARG, "bad device");
  std::lock_guard<std::mutex> lock(mu);
  CK(cudaSetDevice(device));
  if (!st[device]) CK(cudaStreamCreateWithFlags(&st[device], cudaStreamNonBlocking));
  *out = st[device];
  return 0;
}

// ---- NCCL, bound at run time to the library already loaded in the process (torch's), else the system one:
// no link-time dependency, and a communicator handed in by the caller belongs to the same library.
struct Id128 { char bytes[128]; };     // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed by value
struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*CommCount)(void*, int*) = nullptr;
  int (*CommUserRank)(void*, int*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
int nccl_api(Nccl** out) {
  static Nccl api;
  static std::mutex mu;
  static int state = 0;           // 0 untried, 1 ok, -1 unavailable
  std::lock_guard<std::mutex> lock(mu);
  if (state == 0) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    state = -1;
    if (h) {
      api.lib = h;
      api.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(h, "ncclCommInitRank");
      api.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
      api.CommCount = (int (*)(void*, int*))dlsym(h, "ncclCommCount");
      api.CommUserRank = (int (*)(void*, int*))dlsym(h, "ncclCommUserRank");
      api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
      api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
      if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.CommCount && api.AllGather) state = 1;
    }
  }
  if (state != 1) return fail(QMPS_ERR_UNSUPPORTED, "NCCL (libnccl.so.2) is not available in this process");
  *out = &api;
  return 0;
}
int nccl_fail(Nccl* n, int rc, const char* what) {
  return fail(QMPS_ERR_CUDA, std::string(what) + ": " + (n->GetErrorString ? n->GetErrorString(rc) : "NCCL error"));
}

// final pass of the cross-rank argmin: world (cost, index) pairs -> the best one (ties: smallest index;
// index < 0 marks an empty shard)
__global__ void argmin_pairs_kernel(const long long* __restrict__ pairs, int world, double* best_cost, long long* best_index) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double bc = INFINITY;
  long long bi = -1;
  for (int r = 0; r < world; ++r) {
    const double c = __longlong_as_double(pairs[2 * r]);
    const long long i = pairs[2 * r + 1];
    if (i < 0) continue;
    if (bi < 0 || c < bc || (c == bc && i < bi)) { bc = c; bi = i; }
  }
  *best_cost = bc;
  *best_index = bi;
}
__global__ void widen_kernel(int64_t n, const float* __restrict__ in, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (double)in[i];
}
__global__ void pack_pair_kernel(const double* c, const long long* i, long long* pair) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { pair[0] = __double_as_longlong(*c); pair[1] = *i; }
}

// ---- device-resident evolution strategy for the TDVP-step optimisation -----------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {        // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// candidates of one generation: row 0 is the centre itself, row c >= 1 is centre + sigma * 10^(-(c mod 4)/2) * normal
__global__ void es_propose_kernel(int npop, int P, const double* __restrict__ centre, const double* __restrict__ sigma,
                                  unsigned long long seed, unsigned long long gen, double* __restrict__ cand) {
  const double sg = *sigma;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < npop * P; e += gridDim.x * blockDim.x) {
    const int c = e / P, j = e - c * P;
    double v = centre[j];
    if (c > 0) {
      const unsigned long long h1 = mix64(seed ^ mix64(gen * 0x100000001B3ull + (unsigned long long)e));
      const unsigned long long h2 = mix64(h1 ^ 0xD6E8FEB86659FD93ull);
      const double u1 = ((double)(h1 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
      const double u2 = ((double)(h2 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
      const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
      const int grp = c & 3;
      const double sc = grp == 0 ? 1.0 : grp == 1 ? 0.316227766016838 : grp == 2 ? 0.1 : 0.0316227766016838;
      v += sg * sc * z;
    }
    cand[e] = v;
  }
}
// the winner becomes the centre; the step size follows the scale that won (halved when nothing beat the centre)
__global__ void es_select_kernel(int P, const double* __restrict__ cand, const long long* __restrict__ best_index,
                                 const double* __restrict__ best_cost, double* __restrict__ centre, double* __restrict__ sigma,
                                 double* __restrict__ cost_out) {
  const long long bi = *best_index;
  for (int j = threadIdx.x; j < P; j += blockDim.x) centre[j] = cand[bi * P + j];
  if (threadIdx.x == 0) {
    const int grp = (int)(bi & 3);
    const double sc = grp == 0 ? 1.0 : grp == 1 ? 0.316227766016838 : grp == 2 ? 0.1 : 0.0316227766016838;
    double sg = *sigma;
    sg = bi == 0 ? sg * 0.5 : sg * sc * 2.0;
    if (sg < 1e-9) sg = 1e-9;
    if (sg > 1.0) sg = 1.0;
    *sigma = sg;
    if (cost_out) *cost_out = *best_cost;
  }
}
// ---- device-resident BFGS with batched finite-difference gradients and a batched line search ----------------
// (what scipy.optimize.minimize's default does for the reference, scripts/loschmidt.py:371, with every cost
// evaluation of an iteration issued as ONE launch and no value ever leaving the device)
constexpr int BFGS_LS = 24;                 // line-search step sizes 2, 1, 1/2, ..., 2^-22
__global__ void bfgs_probe_kernel(int P, const double* __restrict__ theta, double h, double* __restrict__ cand) {
  for (int e = threadIdx.x; e < (2 * P + 1) * P; e += blockDim.x) {
    const int c = e / P, j = e - c * P;
    double v = theta[j];
    if (c >= 1 && c <= P && c - 1 == j) v += h;
    if (c > P && c - P - 1 == j) v -= h;
    cand[e] = v;
  }
}
// state: g[P], gprev[P], s[P], d[P], H[P*P], f0, have_prev.  One CTA of 64 threads, P <= 64.
__global__ void bfgs_direction_kernel(int P, const double* __restrict__ costs, double h, double* g, double* gprev, double* s,
                                      double* d, double* H, double* f0, int* have_prev, double* __restrict__ theta,
                                      double* __restrict__ cand) {
  __shared__ double sh_y[64], sh_Hy[64], sh_red[64];
  const int j = threadIdx.x;
  if (j < P) g[j] = (costs[1 + j] - costs[1 + P + j]) / (2.0 * h);
  if (j == 0) *f0 = costs[0];
  __syncthreads();
  if (*have_prev) {
    // BFGS update of the inverse Hessian with s = last step, y = g - gprev (skipped unless y.s > 0)
    if (j < P) sh_y[j] = g[j] - gprev[j];
    __syncthreads();
    double ys = 0.0, ss = 0.0;
    for (int k = 0; k < P; ++k) { ys += sh_y[k] * s[k]; ss += s[k] * s[k]; }
    if (ys > 1e-300 && ss > 0.0 && ys > 1e-12 * sqrt(ss)) {
      if (j < P) { double a = 0.0; for (int k = 0; k < P; ++k) a += H[j * P + k] * sh_y[k]; sh_Hy[j] = a; }
      __syncthreads();
      double yHy = 0.0;
      for (int k = 0; k < P; ++k) yHy += sh_y[k] * sh_Hy[k];
      const double rho = 1.0 / ys;
      if (j < P)
        for (int k = 0; k < P; ++k)
          H[j * P + k] += -rho * (sh_Hy[j] * s[k] + s[j] * sh_Hy[k]) + rho * rho * yHy * s[j] * s[k] + rho * s[j] * s[k];
    }
    __syncthreads();
  } else {                                   // first iteration, or the last line search failed: steepest descent restart
    if (j < P) for (int k = 0; k < P; ++k) H[j * P + k] = (j == k) ? 1.0 : 0.0;
    __syncthreads();
  }
  if (j < P) { double a = 0.0; for (int k = 0; k < P; ++k) a += H[j * P + k] * g[k]; d[j] = -a; }
  __syncthreads();
  if (j < P) sh_red[j] = d[j] * g[j];
  __syncthreads();
  double dg = 0.0;
  for (int k = 0; k < P; ++k) dg += sh_red[k];
  if (!(dg < 0.0)) {                         // not a descent direction: restart from steepest descent
    __syncthreads();
    if (j < P) { for (int k = 0; k < P; ++k) H[j * P + k] = (j == k) ? 1.0 : 0.0; d[j] = -g[j]; }
  }
  __syncthreads();
  if (j < P) gprev[j] = g[j];
  // line-search candidates theta + alpha_l d
  for (int e = threadIdx.x; e < BFGS_LS * P; e += blockDim.x) {
    const int l = e / P, k = e - l * P;
    cand[e] = theta[k] + ldexp(2.0, -l) * d[k];
  }
}
__global__ void bfgs_step_kernel(int P, const double* __restrict__ costs, const double* __restrict__ d, double* s, const double* f0,
                                 int* have_prev, double* __restrict__ theta, double* __restrict__ fbest) {
  __shared__ int best_l;
  if (threadIdx.x == 0) {
    int bl = -1; double bc = *f0;
    for (int l = 0; l < BFGS_LS; ++l) if (costs[l] < bc) { bc = costs[l]; bl = l; }
    best_l = bl;
    *fbest = bc;
    *have_prev = bl >= 0;                    // no progress: the next direction restarts without an update
  }
  __syncthreads();
  const int j = threadIdx.x;
  if (j < P) {
    const double step = best_l >= 0 ? ldexp(2.0, -best_l) * d[j] : 0.0;
    s[j] = step;
    theta[j] += step;
  }
}
__global__ void bfgs_init_kernel(int P, double* H, int* have_prev) {
  for (int e = threadIdx.x; e < P * P; e += blockDim.x) H[e] = (e / P == e % P) ? 1.0 : 0.0;
  if (threadIdx.x == 0) *have_prev = 0;
}
__global__ void narrow_kernel(const double* in, float* out) { if (threadIdx.x == 0 && blockIdx.x == 0) *out = (float)*in; }
__global__ void set_scalar_kernel(double* p, double v) { if (threadIdx.x == 0 && blockIdx.x == 0) *p = v; }

}  // namespace

extern "C" {

// a11 in one call: cost / echo / eta [NP][NT] for NP parameter vectors and NT two-site gates
int qmps_loschmidt_batched(const qmps_gate_op* ops, int nops, int nq, int64_t NP, int P, const double* theta,
                           const void* A0, int64_t NT, const void* W, void* cost, void* echo, void* eta,
                           int32_t* status, int dtype, void* stream) {
  if (!ops || nq < 2 || nq > 5 || NP < 0 || NT < 0 || (NP && !theta) || (NT && !W) || !A0)
    return fail(QMPS_ERR_ARG, "loschmidt_batched: bad arguments");
  if (dtype != QMPS_C128 && dtype != QMPS_C64) return fail(QMPS_ERR_ARG, "loschmidt_batched: bad dtype");
  if (NP == 0 || NT == 0) return 0;
  const int D = 1 << (nq - 1);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t cs = csize(dtype), tsz = (size_t)D * D;
  Scratch scratch(st);
  char *B = nullptr, *MB = nullptr, *WMA = nullptr;
  CK(scratch.get(&B, cs * NP * 2 * tsz));
  CK(scratch.get(&MB, cs * NP * 4 * tsz));
  CK(scratch.get(&WMA, cs * NT * 4 * tsz));
  if (int rc = qmps_ansatz(ops, nops, nq, NP, P, theta, 0, B, dtype, stream)) return rc;            // B_p
  if (int rc = qmps_merge(2, 2, D, NP, B, NP, B, 0, nullptr, MB, dtype, stream)) return rc;        // merge(B_p, B_p)
  if (int rc = qmps_merge(2, 2, D, 1, A0, 1, A0, NT, W, WMA, dtype, stream)) return rc;            // W_k . merge(A0, A0)
  // Map(W_k . merge(A0,A0), merge(B_p,B_p)): A side indexed by k, B side by p, output [p][k]
  return qmps_fixed_point_ex(4, D, NT, WMA, NP, MB, 2, 0, QMPS_GAUGE_ZGEEV, eta, nullptr, cost, echo, nullptr, status,
                             dtype, stream);
}

int qmps_loschmidt_batched_host(const qmps_gate_op* ops, int nops, int nq, int64_t NP, int P, const double* theta,
                                const void* A0, int64_t NT, const void* W, void* cost, void* echo, void* eta,
                                int32_t* status, int dtype, int device) {
  if (NP < 0 || NT < 0 || (NP && !theta) || (NT && !W) || !A0 || nq < 2 || nq > 5)
    return fail(QMPS_ERR_ARG, "loschmidt_batched_host: bad arguments");
  if (NP == 0 || NT == 0) return 0;
  cudaStream_t st;
  if (int rc = host_stream(device, &st)) return rc;
  const int D = 1 << (nq - 1);
  const size_t cs = csize(dtype), rs = rsize(dtype), n = (size_t)NP * NT;
  Scratch scratch(st);
  double* dth = nullptr; char *dA0 = nullptr, *dW = nullptr, *dcost = nullptr, *decho = nullptr, *deta = nullptr; int32_t* dst = nullptr;
  CK(scratch.get(&dth, sizeof(double) * NP * P));
  CK(scratch.get(&dA0, cs * 2 * D * D));
  CK(scratch.get(&dW, cs * NT * 16));
  if (cost) CK(scratch.get(&dcost, rs * n));
  if (echo) CK(scratch.get(&decho, rs * n));
  if (eta) CK(scratch.get(&deta, cs * n));
  if (status) CK(scratch.get(&dst, sizeof(int32_t) * n));
  CK(cudaMemcpyAsync(dth, theta, sizeof(double) * NP * P, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dA0, A0, cs * 2 * D * D, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dW, W, cs * NT * 16, cudaMemcpyHostToDevice, st));
  if (int rc = qmps_loschmidt_batched(ops, nops, nq, NP, P, dth, dA0, NT, dW, dcost, decho, deta, dst, dtype, st)) return rc;
  if (cost) CK(cudaMemcpyAsync(cost, dcost, rs * n, cudaMemcpyDeviceToHost, st));
  if (echo) CK(cudaMemcpyAsync(echo, decho, rs * n, cudaMemcpyDeviceToHost, st));
  if (eta) CK(cudaMemcpyAsync(eta, deta, cs * n, cudaMemcpyDeviceToHost, st));
  if (status) CK(cudaMemcpyAsync(status, dst, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

// a9 / a12 host-buffer form: theta [N][P] in (8 P bytes per vector), energy [N][max(nshift,1)] out
int qmps_energy_theta_host(const qmps_gate_op* ops, int nops, int nq, int64_t N, int P, const double* theta,
                           const void* hmat, int coord, const double* shifts, int nshift, void* energy,
                           int32_t* status, int dtype, int device) {
  if (N < 0 || (N && (!theta || !energy)) || !hmat) return fail(QMPS_ERR_ARG, "energy_theta_host: bad arguments");
  if (N == 0) return 0;
  cudaStream_t st;
  if (int rc = host_stream(device, &st)) return rc;
  const size_t ne = (size_t)N * (nshift > 0 ? nshift : 1);
  Scratch scratch(st);
  double* dth = nullptr; char *dH = nullptr, *dE = nullptr; int32_t* dst = nullptr;
  CK(scratch.get(&dth, sizeof(double) * N * P));
  CK(scratch.get(&dH, csize(dtype) * 16));
  CK(scratch.get(&dE, rsize(dtype) * ne));
  if (status) CK(scratch.get(&dst, sizeof(int32_t) * ne));
  CK(cudaMemcpyAsync(dth, theta, sizeof(double) * N * P, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dH, hmat, csize(dtype) * 16, cudaMemcpyHostToDevice, st));
  if (int rc = qmps_energy_theta(ops, nops, nq, N, P, dth, dH, coord, shifts, nshift, dE, dst, dtype, st)) return rc;
  CK(cudaMemcpyAsync(energy, dE, rsize(dtype) * ne, cudaMemcpyDeviceToHost, st));
  if (status) CK(cudaMemcpyAsync(status, dst, sizeof(int32_t) * ne, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

// a12: n_sweeps whole coordinate sweeps on the device (qmps/rotosolve.py:154-181 with three shifts,
// qmps/tools.py:422-457 with six); theta [N][P] is updated in place, energy [N] (optional) is the cost after
// the last sweep.  2 P launches per sweep, no host round trip.
int qmps_rotosolve_sweep(const qmps_gate_op* ops, int nops, int nq, int64_t N, int P, double* theta, const void* hmat,
                         int n_sweeps, int two_frequency, void* energy, int dtype, void* stream) {
  if (!ops || N < 0 || P < 1 || (N && !theta) || !hmat || n_sweeps < 0)
    return fail(QMPS_ERR_ARG, "rotosolve_sweep: bad arguments");
  if (N == 0) return 0;
  static const double sh3[3] = {0.0, M_PI / 2, -M_PI / 2};
  static const double sh6[6] = {0.0, M_PI, M_PI / 2, -M_PI / 2, M_PI / 4, -M_PI / 4};
  const int ns = two_frequency ? 6 : 3;
  const double* sh = two_frequency ? sh6 : sh3;
  cudaStream_t st = (cudaStream_t)stream;
  Scratch scratch(st);
  char* e = nullptr; double *e64 = nullptr, *tstar = nullptr, *fit = nullptr;
  CK(scratch.get(&e, rsize(dtype) * N * ns));
  CK(scratch.get(&tstar, sizeof(double) * N));
  if (two_frequency) CK(scratch.get(&fit, sizeof(double) * N * 8));
  if (dtype != QMPS_C128) CK(scratch.get(&e64, sizeof(double) * N * ns));
  for (int w = 0; w < n_sweeps; ++w) {
    for (int i = 0; i < P; ++i) {
      if (int rc = qmps_energy_theta(ops, nops, nq, N, P, theta, hmat, i, sh, ns, e, nullptr, dtype, stream)) return rc;
      const double* costs = (const double*)e;
      if (dtype != QMPS_C128) {
        widen_kernel<<<(unsigned)((N * ns + 255) / 256 < 4096 ? (N * ns + 255) / 256 : 4096), 256, 0, st>>>(N * ns, (const float*)e, e64);
        costs = e64;
      }
      if (int rc = qmps_rotosolve_fit(N, ns, costs, tstar, fit, theta, P, i, stream)) return rc;
    }
  }
  if (energy) return qmps_energy_theta(ops, nops, nq, N, P, theta, hmat, -1, nullptr, 0, energy, nullptr, dtype, stream);
  return 0;
}

// (f)-2: the reference's time-evolution loop (scripts/loschmidt.py:367-375 = qmps/loschmidts/time_evo.py:140-150)
// on the device, no host round trip: for each step  theta <- argmin_p obj(p, A(theta), W)  with
// obj = -sqrt|eta(Map(W . merge(A,A), merge(B_p,B_p)))|, minimised by a population search (npop candidates per
// generation in ONE qmps_loschmidt_batched call, the argmin fed back on the device), then the echo series
// |eta(E_{A_t A_0})|^2 (``A_.overlap(A)``) for the whole trajectory.
int qmps_loschmidt_trajectory(const qmps_gate_op* ops, int nops, int nq, int P, const double* theta0, const void* W,
                              int n_steps, int n_gen, int npop, double sigma0, uint64_t seed, int n_bfgs, double* theta_traj,
                              void* step_cost, void* echo, int dtype, void* stream) {
  if (!ops || nq < 2 || nq > 5 || P < 1 || P > 64 || !theta0 || !W || n_steps < 0 || n_gen < 0 || n_bfgs < 0 || (n_gen && (npop < 8 || !(sigma0 > 0))) || !theta_traj)
    return fail(QMPS_ERR_ARG, "loschmidt_trajectory: bad arguments");
  if (dtype != QMPS_C128 && dtype != QMPS_C64) return fail(QMPS_ERR_ARG, "loschmidt_trajectory: bad dtype");
  cudaStream_t st = (cudaStream_t)stream;
  const int D = 1 << (nq - 1);
  const size_t cs = csize(dtype), rs = rsize(dtype);
  Scratch scratch(st);
  double *cand = nullptr, *sigma = nullptr, *bc = nullptr; long long* bi = nullptr; char *At = nullptr, *cost = nullptr, *Aall = nullptr;
  double* cost64 = nullptr;
  const int ncand = npop > 2 * P + 1 ? (npop > BFGS_LS ? npop : BFGS_LS) : (2 * P + 1 > BFGS_LS ? 2 * P + 1 : BFGS_LS);
  double *bg = nullptr, *bgp = nullptr, *bs = nullptr, *bd = nullptr, *bH = nullptr, *bf0 = nullptr; int* bhave = nullptr;
  CK(scratch.get(&bg, sizeof(double) * P)); CK(scratch.get(&bgp, sizeof(double) * P)); CK(scratch.get(&bs, sizeof(double) * P));
  CK(scratch.get(&bd, sizeof(double) * P)); CK(scratch.get(&bH, sizeof(double) * P * P)); CK(scratch.get(&bf0, sizeof(double)));
  CK(scratch.get(&bhave, sizeof(int)));
  CK(scratch.get(&cand, sizeof(double) * (size_t)ncand * P));
  CK(scratch.get(&sigma, sizeof(double)));
  CK(scratch.get(&bc, sizeof(double)));
  CK(scratch.get(&bi, sizeof(long long)));
  CK(scratch.get(&At, cs * 2 * D * D));
  CK(scratch.get(&cost, rs * ncand));
  if (dtype != QMPS_C128) CK(scratch.get(&cost64, sizeof(double) * ncand));
  CK(cudaMemcpyAsync(theta_traj, theta0, sizeof(double) * P, cudaMemcpyDeviceToDevice, st));
  const int pg = (npop * P + 255) / 256;
  for (int step = 0; step < n_steps; ++step) {
    double* centre = theta_traj + (size_t)(step + 1) * P;
    CK(cudaMemcpyAsync(centre, theta_traj + (size_t)step * P, sizeof(double) * P, cudaMemcpyDeviceToDevice, st));
    if (int rc = qmps_ansatz(ops, nops, nq, 1, P, theta_traj + (size_t)step * P, 0, At, dtype, stream)) return rc;      // A_t
    set_scalar_kernel<<<1, 32, 0, st>>>(sigma, sigma0);
    for (int gen = 0; gen < n_gen; ++gen) {
      es_propose_kernel<<<pg, 256, 0, st>>>(npop, P, centre, sigma, (unsigned long long)seed,
                                            (unsigned long long)step * (unsigned long long)n_gen + gen, cand);
      if (int rc = qmps_loschmidt_batched(ops, nops, nq, npop, P, cand, At, 1, W, cost, nullptr, nullptr, nullptr, dtype, stream)) return rc;
      const double* c64 = (const double*)cost;
      if (dtype != QMPS_C128) { widen_kernel<<<(npop + 255) / 256, 256, 0, st>>>(npop, (const float*)cost, cost64); c64 = cost64; }
      if (int rc = qmps_argmin(npop, c64, 0, bc, (int64_t*)bi, stream)) return rc;
      es_select_kernel<<<1, 64, 0, st>>>(P, cand, bi, bc, centre, sigma, nullptr);
    }
    // BFGS refinement from the population's best point (or from theta_t when n_gen = 0)
    auto evaluate = [&](int n) -> int {
      if (int rc = qmps_loschmidt_batched(ops, nops, nq, n, P, cand, At, 1, W, cost, nullptr, nullptr, nullptr, dtype, stream)) return rc;
      if (dtype != QMPS_C128) widen_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, (const float*)cost, cost64);
      return 0;
    };
    const double* c64 = dtype == QMPS_C128 ? (const double*)cost : cost64;
    const double hfd = dtype == QMPS_C128 ? 1e-5 : 3e-3;
    if (n_bfgs > 0) bfgs_init_kernel<<<1, 64, 0, st>>>(P, bH, bhave);
    for (int it = 0; it < n_bfgs; ++it) {
      bfgs_probe_kernel<<<1, 256, 0, st>>>(P, centre, hfd, cand);
      if (int rc = evaluate(2 * P + 1)) return rc;
      bfgs_direction_kernel<<<1, 64, 0, st>>>(P, c64, hfd, bg, bgp, bs, bd, bH, bf0, bhave, centre, cand);
      if (int rc = evaluate(BFGS_LS)) return rc;
      bfgs_step_kernel<<<1, 64, 0, st>>>(P, c64, bd, bs, bf0, bhave, centre, bc);
    }
    if (step_cost) {
      if (dtype == QMPS_C128) CK(cudaMemcpyAsync((double*)step_cost + step, bc, sizeof(double), cudaMemcpyDeviceToDevice, st));
      else narrow_kernel<<<1, 32, 0, st>>>(bc, (float*)step_cost + step);
    }
  }
  if (echo) {
    const int64_t NT = (int64_t)n_steps + 1;
    CK(scratch.get(&Aall, cs * NT * 2 * D * D));
    if (int rc = qmps_ansatz(ops, nops, nq, NT, P, theta_traj, 0, Aall, dtype, stream)) return rc;
    if (int rc = qmps_fixed_point(2, D, NT, Aall, 1, Aall, 0, 0, nullptr, nullptr, nullptr, nullptr, echo, nullptr, dtype, stream)) return rc;
  }
  CK(cudaGetLastError());
  return 0;
}


// ---- (f)-4 PXP scar dynamics (the reference's scars.py): batched step cost and the whole trajectory on the device -----
int qmps_scars_cost(int64_t N, const double* params, int64_t NC, const double* current, const void* W, void* cost, void* eta,
                    int32_t* status, int dtype, void* stream) {
  if (N < 0 || (N && (!params || !current || !W || !cost)) || (NC != 1 && NC != N))
    return fail(QMPS_ERR_ARG, "scars_cost: bad arguments (NC must be 1 or N)");
  if (dtype != QMPS_C128 && dtype != QMPS_C64) return fail(QMPS_ERR_ARG, "scars_cost: bad dtype");
  return scars_cost_any(N, params, NC, current, W, cost, eta, status, dtype, (cudaStream_t)stream);
}

// simulate_scars (scars.py:157-170): per time step minimise the step cost over the four angles, starting from the
// current ones -- here a Gaussian population (one launch per generation) followed by BFGS with batched
// finite-difference gradients, the argmin fed back on the device.  traj [n_steps + 1][4] (DEVICE, row 0 = params0),
// step_cost [n_steps] (real of the precision, optional).
int qmps_scars_trajectory(const double* params0, const void* W, int n_steps, int n_gen, int npop, double sigma0, uint64_t seed,
                          int n_bfgs, double* traj, void* step_cost, int dtype, void* stream) {
  const int P = 4;
  if (!params0 || !W || n_steps < 0 || n_gen < 0 || n_bfgs < 0 || (n_gen && (npop < 8 || !(sigma0 > 0))) || !traj)
    return fail(QMPS_ERR_ARG, "scars_trajectory: bad arguments");
  if (dtype != QMPS_C128 && dtype != QMPS_C64) return fail(QMPS_ERR_ARG, "scars_trajectory: bad dtype");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t rs = rsize(dtype);
  Scratch scratch(st);
  const int ncand = npop > 2 * P + 1 ? (npop > BFGS_LS ? npop : BFGS_LS) : (2 * P + 1 > BFGS_LS ? 2 * P + 1 : BFGS_LS);
  double *cand = nullptr, *sigma = nullptr, *bc = nullptr, *cost64 = nullptr; long long* bi = nullptr; char* cost = nullptr;
  double *bg = nullptr, *bgp = nullptr, *bs = nullptr, *bd = nullptr, *bH = nullptr, *bf0 = nullptr; int* bhave = nullptr;
  CK(scratch.get(&bg, sizeof(double) * P)); CK(scratch.get(&bgp, sizeof(double) * P)); CK(scratch.get(&bs, sizeof(double) * P));
  CK(scratch.get(&bd, sizeof(double) * P)); CK(scratch.get(&bH, sizeof(double) * P * P)); CK(scratch.get(&bf0, sizeof(double)));
  CK(scratch.get(&bhave, sizeof(int)));
  CK(scratch.get(&cand, sizeof(double) * (size_t)ncand * P));
  CK(scratch.get(&sigma, sizeof(double))); CK(scratch.get(&bc, sizeof(double))); CK(scratch.get(&bi, sizeof(long long)));
  CK(scratch.get(&cost, rs * ncand));
  if (dtype != QMPS_C128) CK(scratch.get(&cost64, sizeof(double) * ncand));
  CK(cudaMemcpyAsync(traj, params0, sizeof(double) * P, cudaMemcpyDeviceToDevice, st));
  const double* c64 = dtype == QMPS_C128 ? (const double*)cost : cost64;
  for (int step = 0; step < n_steps; ++step) {
    const double* cur = traj + (size_t)step * P;
    double* centre = traj + (size_t)(step + 1) * P;
    CK(cudaMemcpyAsync(centre, cur, sizeof(double) * P, cudaMemcpyDeviceToDevice, st));
    auto evaluate = [&](int n) -> int {
      if (int rc = scars_cost_any(n, cand, 1, cur, W, cost, nullptr, nullptr, dtype, st)) return rc;
      if (dtype != QMPS_C128) widen_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, (const float*)cost, cost64);
      return 0;
    };
    set_scalar_kernel<<<1, 32, 0, st>>>(sigma, sigma0);
    for (int gen = 0; gen < n_gen; ++gen) {
      es_propose_kernel<<<(npop * P + 255) / 256, 256, 0, st>>>(npop, P, centre, sigma, (unsigned long long)seed,
                                                                (unsigned long long)step * (unsigned long long)n_gen + gen, cand);
      if (int rc = evaluate(npop)) return rc;
      if (int rc = qmps_argmin(npop, c64, 0, bc, (int64_t*)bi, stream)) return rc;
      es_select_kernel<<<1, 64, 0, st>>>(P, cand, bi, bc, centre, sigma, nullptr);
    }
    const double hfd = dtype == QMPS_C128 ? 1e-5 : 3e-3;
    if (n_bfgs > 0) bfgs_init_kernel<<<1, 64, 0, st>>>(P, bH, bhave);
    for (int it = 0; it < n_bfgs; ++it) {
      bfgs_probe_kernel<<<1, 256, 0, st>>>(P, centre, hfd, cand);
      if (int rc = evaluate(2 * P + 1)) return rc;
      bfgs_direction_kernel<<<1, 64, 0, st>>>(P, c64, hfd, bg, bgp, bs, bd, bH, bf0, bhave, centre, cand);
      if (int rc = evaluate(BFGS_LS)) return rc;
      bfgs_step_kernel<<<1, 64, 0, st>>>(P, c64, bd, bs, bf0, bhave, centre, bc);
    }
    if (step_cost) {
      if (dtype == QMPS_C128) CK(cudaMemcpyAsync((double*)step_cost + step, bc, sizeof(double), cudaMemcpyDeviceToDevice, st));
      else narrow_kernel<<<1, 32, 0, st>>>(bc, (float*)step_cost + step);
    }
  }
  CK(cudaGetLastError());
  return 0;
}

// ---- (e) cross-rank argmin over NCCL ------------------------------------------------------------------
int qmps_nccl_unique_id(void* id128) {
  Nccl* n;
  if (int rc = nccl_api(&n)) return rc;
  if (!id128) return fail(QMPS_ERR_ARG, "nccl_unique_id: null buffer");
  if (int rc = n->GetUniqueId(id128)) return nccl_fail(n, rc, "ncclGetUniqueId");
  return 0;
}

int qmps_nccl_comm_create(const void* id128, int world, int rank, void** comm) {
  Nccl* n;
  if (int rc = nccl_api(&n)) return rc;
  if (!id128 || !comm || world < 1 || rank < 0 || rank >= world) return fail(QMPS_ERR_ARG, "nccl_comm_create: bad arguments");
  Id128 id;
  memcpy(id.bytes, id128, 128);
  if (int rc = n->CommInitRank(comm, world, id, rank)) return nccl_fail(n, rc, "ncclCommInitRank");
  return 0;
}

int qmps_nccl_comm_destroy(void* comm) {
  Nccl* n;
  if (int rc = nccl_api(&n)) return rc;
  if (!comm) return 0;
  if (int rc = n->CommDestroy(comm)) return nccl_fail(n, rc, "ncclCommDestroy");
  return 0;
}

// (min cost, global argmin) over the shards of all ranks of `comm` (an ncclComm_t: the caller's own, or
// torch's ProcessGroupNCCL communicator): local block reduction, one 16-byte-per-rank ncclAllGather and a
// one-thread final pass, all on `stream`; best_cost / best_index are DEVICE scalars, identical on every rank.
// comm == NULL: single rank.
int qmps_argmin_allreduce(void* comm, int64_t N, const double* cost, int64_t index_offset, double* best_cost,
                          int64_t* best_index, void* stream) {
  if (N < 0 || (N && !cost) || !best_cost || !best_index) return fail(QMPS_ERR_ARG, "argmin_allreduce: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = qmps_argmin(N, cost, index_offset, best_cost, best_index, stream)) return rc;
  if (!comm) return 0;
  Nccl* n;
  if (int rc = nccl_api(&n)) return rc;
  int world = 1;
  if (int rc = n->CommCount(comm, &world)) return nccl_fail(n, rc, "ncclCommCount");
  if (world <= 1) return 0;
  Scratch scratch(st);
  long long *pair = nullptr, *all = nullptr;
  CK(scratch.get(&pair, sizeof(long long) * 2));
  CK(scratch.get(&all, sizeof(long long) * 2 * world));
  pack_pair_kernel<<<1, 32, 0, st>>>(best_cost, (const long long*)best_index, pair);
  if (int rc = n->AllGather(pair, all, 2, /* ncclInt64 */ 4, comm, st)) return nccl_fail(n, rc, "ncclAllGather");
  argmin_pairs_kernel<<<1, 32, 0, st>>>(all, world, best_cost, (long long*)best_index);
  CK(cudaGetLastError());
  return 0;
}

}  // extern "C"
