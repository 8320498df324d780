// qmps_b200 host-side helpers shared by the translation units of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <string>

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

inline int sm_count() {
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
  return v;
}

inline bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

// persistent grid: one wave of resident CTAs (SM count x occupancy), never more than needed
template <typename K>
int persistent_grid(K kernel, int block, size_t smem, int64_t blocks_needed, int* out_grid) {
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem);
  if (e != cudaSuccess) return fail(QMPS_ERR_CUDA, std::string("occupancy: ") + cudaGetErrorString(e));
  if (per_sm < 1) return fail(QMPS_ERR_UNSUPPORTED, "kernel does not fit on an SM with this configuration");
  int64_t cap = (int64_t)sm_count() * per_sm;
  int64_t g = blocks_needed < cap ? blocks_needed : cap;
  *out_grid = (int)(g < 1 ? 1 : g);
  return 0;
}

template <typename K> int allow_smem(K kernel, size_t smem) {
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return 0;
}

// stream-ordered device copy of a small host array
template <typename U> int to_device_async(const U* host, size_t count, U** dev, cudaStream_t st) {
  *dev = nullptr;
  if (count == 0) return 0;
  CK(cudaMallocAsync((void**)dev, count * sizeof(U), st));
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
int ansatz_f64(const qmps::GateOp* dops, int nops, int nq, int64_t N, int P, const double* theta, int full, void* out, cudaStream_t st);
int ansatz_f32(const qmps::GateOp* dops, int nops, int nq, int64_t N, int P, const double* theta, int full, void* out, cudaStream_t st);
// capi_d2.cu
int env_d2(int64_t N, const void* in, int in_is_U, void* eta, void* r, void* C, int32_t* status, int dtype, cudaStream_t st);
int energy_d2_theta(const qmps::GateOp* dops, int nops, int64_t N, int P, const double* theta, const void* hmat, int coord,
                    const double* dshifts, int nshift, void* energy, int32_t* status, int dtype, cudaStream_t st);
int d2_max_ops();

}  // namespace qmps_host
