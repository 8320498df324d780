// qmps_b200 host-side helpers shared by the translation units of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <string>
#include <mutex>
#include <unordered_map>
#include <functional>

#include "../../include/qmps_b200.h"
#include "kernels_generic.cuh"

namespace qmps_host {

std::string& last_error();                        // thread-local, defined in capi.cu
inline int fail(int code, const std::string& msg) { last_error() = msg; return code; }
#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return qmps_host::fail(QMPS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// tuning knobs (qmps_set_option); defaults are the measured best
enum { OPT_D2_PDL = 0, OPT_D2_CTAS_PER_SM = 1, OPT_FP16_FAST = 2, OPT_ENV_REAL = 3, OPT_TC_POWER = 4, OPT_TC_PERSISTENT = 5, OPT_FP_GROUP = 6, OPT_FP_BLOCK = 7, OPT_ER_WIDE = 8, OPT_FP_D2 = 9, OPT_BW_THREAD = 10, OPT_I8_POWER = 11, OPT_TC_PRESPLIT = 12, OPT_FP64_FAST = 13, OPT_COUNT = 14 };
int option_get(int key);                          // defined in capi.cu

inline int sm_count() {
  static int cached[64] = {0};
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev];
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
  if (dev >= 0 && dev < 64) cached[dev] = v;
  return v;
}

inline bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

// per-(kernel, block, smem) launch set-up is cached: the occupancy query and the shared-memory
// opt-in cost microseconds of host time, which matters for 40 us kernels launched back to back
struct LaunchKey {
  const void* fn; int block; size_t smem;
  bool operator==(const LaunchKey& o) const { return fn == o.fn && block == o.block && smem == o.smem; }
};
struct LaunchKeyHash {
  size_t operator()(const LaunchKey& k) const { return std::hash<const void*>()(k.fn) ^ (k.smem * 1315423911u) ^ (size_t)k.block; }
};
std::unordered_map<LaunchKey, int, LaunchKeyHash>& occupancy_cache();   // defined in capi.cu
std::mutex& occupancy_mutex();

// persistent grid: one wave of resident CTAs (SM count x occupancy), never more than needed
template <typename K>
int persistent_grid(K kernel, int block, size_t smem, int64_t blocks_needed, int* out_grid) {
  int per_sm = 0;
  {
    std::lock_guard<std::mutex> lock(occupancy_mutex());
    auto& cache = occupancy_cache();
    LaunchKey key{(const void*)kernel, block, smem};
    auto it = cache.find(key);
    if (it != cache.end()) per_sm = it->second;
    else {
      cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem);
      if (e != cudaSuccess) return fail(QMPS_ERR_CUDA, std::string("occupancy: ") + cudaGetErrorString(e));
      cache[key] = per_sm;
    }
  }
  if (per_sm < 1) return fail(QMPS_ERR_UNSUPPORTED, "kernel does not fit on an SM with this configuration");
  int64_t cap = (int64_t)sm_count() * per_sm;
  int64_t g = blocks_needed < cap ? blocks_needed : cap;
  *out_grid = (int)(g < 1 ? 1 : g);
  return 0;
}

template <typename K> int allow_smem(K kernel, size_t smem) {
  if (smem <= 48 * 1024) return 0;
  {
    std::lock_guard<std::mutex> lock(occupancy_mutex());
    auto& cache = occupancy_cache();
    int dev = 0;
    cudaGetDevice(&dev);
    LaunchKey key{(const void*)kernel, -1 - dev, smem};       // block < 0: "opt-in done on device -1-block"
    if (cache.find(key) != cache.end()) return 0;
    cache[key] = 1;
  }
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return 0;
}

// The stream-ordered allocator returns freed memory to the driver at the next synchronisation
// unless a release threshold is set; scratch buffers of tens of MiB would then be re-mapped on
// every call (milliseconds).  Keep the pool: once per device.
inline void keep_mempool() {
  static bool done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  done[dev] = true;
}

inline cudaError_t malloc_async(void** ptr, size_t bytes, cudaStream_t st) {
  keep_mempool();
  return cudaMallocAsync(ptr, bytes, st);
}

// stream-ordered scratch that is released on every exit path of a C-ABI entry (early CK returns included)
struct Scratch {
  cudaStream_t st;
  void* ptrs[16];
  int n;
  explicit Scratch(cudaStream_t s) : st(s), n(0) {}
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
  template <typename U> cudaError_t get(U** p, size_t bytes) {
    *p = nullptr;
    if (n >= 16) return cudaErrorMemoryAllocation;
    void* q = nullptr;
    cudaError_t e = malloc_async(&q, bytes ? bytes : 1, st);
    if (e == cudaSuccess) { ptrs[n++] = q; *p = (U*)q; }
    return e;
  }
  ~Scratch() { for (int k = n - 1; k >= 0; --k) cudaFreeAsync(ptrs[k], st); }
};

// stream-ordered device copy of a small host array
template <typename U> int to_device_async(const U* host, size_t count, U** dev, cudaStream_t st) {
  *dev = nullptr;
  if (count == 0) return 0;
  CK(malloc_async((void**)dev, count * sizeof(U), st));
  CK(cudaMemcpyAsync(*dev, host, count * sizeof(U), cudaMemcpyHostToDevice, st));
  return 0;
}

inline int group_for_n(int n) { return n <= 4 ? 4 : n <= 16 ? 16 : n <= 64 ? 128 : 256; }

// ---- entry points implemented in the per-dtype / per-path translation units ----
// capi_generic_f64.cu / capi_generic_f32.cu
int env_generic_f64(const qmps::EnvParams& p, int mode, cudaStream_t st);
int env_generic_f32(const qmps::EnvParams& p, int mode, cudaStream_t st);
int fixed_point_f64(const qmps::FpParams& p, cudaStream_t st);
int fixed_point_f32(const qmps::FpParams& p, cudaStream_t st);
int fp16_debug_f64(unsigned long long* out, int reset);
// coord / dshifts (DEVICE) / nshift: rotosolve shift fan-out, output index n * nshift + s (nshift = 0: plain)
int scars_cost_any(int64_t N, const double* params, int64_t NC, const double* current, const void* W, void* cost, void* eta, int32_t* status, int dtype, cudaStream_t st);
int env_d2_packed(int64_t N, const void* in, int in_is_U, void* packed, cudaStream_t st);
int ansatz_f64(const qmps::GateOp* dops, int nops, int nq, int64_t N, int P, const double* theta, int full, void* out, int coord, const double* dshifts, int nshift, cudaStream_t st);
int ansatz_f32(const qmps::GateOp* dops, int nops, int nq, int64_t N, int P, const double* theta, int full, void* out, int coord, const double* dshifts, int nshift, cudaStream_t st);
// capi_tc_i8.cu
bool i8_shape_ok(int M, int N, int K);
int zgemm_c128_i8(int64_t batch, int M, int N, int K, const void* X, const void* Y, int conj_y, void* C, cudaStream_t st);
bool tm_power_i8_applies(int d, int D, int64_t N);
int tm_power_i8(int d, int D, int64_t N, const void* A, const void* B, void* r_io, int K, void* rayleigh, cudaStream_t st);
// capi_d2.cu
int env_d2(int64_t N, const void* in, int in_is_U, void* eta, void* r, void* C, int32_t* status, int dtype, cudaStream_t st);
int energy_d2_theta(const qmps::GateOp* dops, int nops, int64_t N, int P, const double* theta, const void* hmat, int coord,
                    const double* dshifts, int nshift, void* energy, int32_t* status, int dtype, cudaStream_t st);
int d2_max_ops();
// capi_tc.cu (tcgen05, complex64)
bool tm_power_tc_applies(int d, int D, int64_t N);
int tm_power_tc(int d, int D, int64_t N, const void* A, const void* B, void* r_io, int K, void* rayleigh, cudaStream_t st);
int cgemm_c64_tc(int64_t batch, int nsum, int M, int N, int K, const void* X, const void* Y, int conj_y, void* C, cudaStream_t st);

}  // namespace qmps_host
