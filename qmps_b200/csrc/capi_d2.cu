// qmps_b200: D = 2 fast-path launchers (own translation unit: this is the kernel that
// gets tuned, keep its rebuild short).
#include "api_common.cuh"
#include "kernels_d2.cuh"

using namespace qmps;
namespace qmps_host {
namespace {
int32_t* const kPackedTag = reinterpret_cast<int32_t*>(~uintptr_t(0));   // internal marker: write packed records
int env_exact_d2_c128(int64_t N, const void* in, int in_is_U, void* eta, void* r, void* C, int32_t* status,
                      cudaStream_t st) {
  const int64_t ntiles = (N + 31) / 32;
  const int64_t blocks = (ntiles + D2_WARPS - 1) / D2_WARPS;
  int grid = 1;
#define QMPS_D2_LAUNCH(INU, WC, ...)                                                                      \
  do {                                                                                                    \
    auto kern = env_d2_stream_kernel<INU, WC, ##__VA_ARGS__>;                                             \
    if (int rc = allow_smem(kern, D2_SMEM_BYTES)) return rc;                                              \
    if (int rc = persistent_grid(kern, D2_WARPS * 32, D2_SMEM_BYTES, blocks, &grid)) return rc;          \
    if (option_get(OPT_D2_CTAS_PER_SM) > 0) {                                                             \
      const int64_t cap = (int64_t)sm_count() * option_get(OPT_D2_CTAS_PER_SM);                           \
      if (grid > cap) grid = (int)cap;                                                                    \
    }                                                                                                     \
    cudaLaunchConfig_t cfg;                                                                               \
    memset(&cfg, 0, sizeof(cfg));                                                                         \
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(D2_WARPS * 32);                                         \
    cfg.dynamicSmemBytes = D2_SMEM_BYTES; cfg.stream = st;                                                \
    cudaLaunchAttribute at[1];                                                                            \
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                        \
    at[0].val.programmaticStreamSerializationAllowed = 1;                                                 \
    cfg.attrs = at; cfg.numAttrs = option_get(OPT_D2_PDL) ? 1 : 0;                                        \
    CK(cudaLaunchKernelEx(&cfg, kern, (const cx<double>*)in, N, (cx<double>*)eta, (cx<double>*)r,         \
                          (cx<double>*)C, status));                                                       \
  } while (0)
  if (status == kPackedTag) {                  // packed 64-byte records through r (env_d2_packed)
    status = nullptr;
    if (in_is_U) QMPS_D2_LAUNCH(true, true, true); else QMPS_D2_LAUNCH(false, true, true);
  } else if (in_is_U) { if (C) QMPS_D2_LAUNCH(true, true); else QMPS_D2_LAUNCH(true, false); }
  else { if (C) QMPS_D2_LAUNCH(false, true); else QMPS_D2_LAUNCH(false, false); }
#undef QMPS_D2_LAUNCH
  CK(cudaGetLastError());
  return 0;
}

template <typename T>
int env_exact_d2_simple(int64_t N, const void* in, int in_is_U, void* eta, void* r, void* C, int32_t* status,
                        cudaStream_t st) {
  int grid = 1;
  auto kern = env_d2_simple_kernel<T>;
  if (int rc = persistent_grid(kern, 256, 0, (N + 255) / 256, &grid)) return rc;
  kern<<<grid, 256, 0, st>>>((const cx<T>*)in, N, in_is_U, (cx<T>*)eta, (cx<T>*)r, (cx<T>*)C, status);
  CK(cudaGetLastError());
  return 0;
}
}  // namespace

int d2_max_ops() { return D2_MAX_OPS; }

int env_d2(int64_t N, const void* in, int in_is_U, void* eta, void* r, void* C, int32_t* status, int dtype,
           cudaStream_t st) {
  if (N == 0) return 0;
  if (dtype == QMPS_C128) return env_exact_d2_c128(N, in, in_is_U, eta, r, C, status, st);
  return env_exact_d2_simple<float>(N, in, in_is_U, eta, r, C, status, st);
}

// D = 2 complex128, left-canonical: one 64-byte record per problem (layout: kernels_d2.cuh)
int env_d2_packed(int64_t N, const void* in, int in_is_U, void* packed, cudaStream_t st) {
  if (N == 0) return 0;
  return env_exact_d2_c128(N, in, in_is_U, nullptr, packed, nullptr, kPackedTag, st);
}

int energy_d2_theta(const GateOp* dops, int nops, int64_t N, int P, const double* theta, const void* hmat, int coord,
                    const double* dsh, int nshift, void* energy, int32_t* status, int dtype, cudaStream_t st) {
  const int64_t total = N * (nshift > 0 ? nshift : 1);
  int grid = 1;
  if (dtype == QMPS_C128) {
    if (int rc = persistent_grid(energy_d2_theta_kernel<double>, 128, 0, (total + 127) / 128, &grid)) return rc;
    energy_d2_theta_kernel<double><<<grid, 128, 0, st>>>(dops, nops, N, P, theta, (const cx<double>*)hmat, coord, dsh,
                                                         nshift, (double*)energy, status, nullptr);
  } else {
    if (int rc = persistent_grid(energy_d2_theta_kernel<float>, 128, 0, (total + 127) / 128, &grid)) return rc;
    energy_d2_theta_kernel<float><<<grid, 128, 0, st>>>(dops, nops, N, P, theta, (const cx<float>*)hmat, coord, dsh,
                                                        nshift, (float*)energy, status, nullptr);
  }
  CK(cudaGetLastError());
  return 0;
}
}  // namespace qmps_host
