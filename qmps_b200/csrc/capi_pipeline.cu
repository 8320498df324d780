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
