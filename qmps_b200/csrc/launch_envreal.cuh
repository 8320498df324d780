// qmps_b200: launcher of the register-resident real-form environment solver (D = 4, 8),
// included by the per-precision translation units.
#pragma once
#include "api_common.cuh"
#include "kernels_envreal.cuh"

namespace qmps_host {

template <typename REAL, int D, int MODE, int WIDE>
int launch_env_real(qmps::EnvParams p, cudaStream_t st) {
  using namespace qmps;
  constexpr int n = D * D;
  const int block = D == 8 ? 64 : 128;
  const int gpc = block / n;
  const ErLayout<REAL, D> L = er_layout<REAL, D>(p.d, p.nops, MODE == 1);
  const size_t smem = L.total * gpc;
  auto kern = env_real_kernel<REAL, D, MODE, WIDE>;
  if (int rc = allow_smem(kern, smem)) return rc;
  const int S = (MODE == 1 && p.nshift > 0) ? p.nshift : 1;
  const int64_t total = p.N * S;
  int grid = 1;
  if (int rc = persistent_grid(kern, block, smem, (total + gpc - 1) / gpc, &grid)) return rc;
  p.ws = nullptr; p.ws_stride = 0;
  kern<<<grid, block, smem, st>>>(p);
  CK(cudaGetLastError());
  return 0;
}

// D = 8 complex128 with the elimination on the FP64 tensor pipe (kernels_envdmma.cuh, compiled in capi_envdmma.cu)
int launch_env_dmma(int mode, qmps::EnvParams p, cudaStream_t st);

// true if the fast path applies
inline bool env_real_applies(const qmps::EnvParams& p) {
  return option_get(OPT_ENV_REAL) && p.assume_lc && (p.D == 4 || p.D == 8) && p.d >= 1 && p.d <= 4;
}

template <typename REAL, int MODE>
int dispatch_env_real(const qmps::EnvParams& p, cudaStream_t st) {
  if (p.D == 4) return launch_env_real<REAL, 4, MODE, 0>(p, st);
  // er_wide: -1 = measured best per precision (profiles/sweep_er_r01*.jsonl, sweep_er_r02*.jsonl), 0 / 1 / 2 force a
  // row-per-thread variant, 3 = blocked elimination on the FP64 tensor pipe (complex128 only)
  int v = option_get(OPT_ER_WIDE);
  if (v < 0) v = sizeof(REAL) == 4 ? 2 : 3;
  if (v == 3) {
    if constexpr (sizeof(REAL) == 8) return launch_env_dmma(MODE, p, st);
    v = 2;
  }
  if (v == 1) return launch_env_real<REAL, 8, MODE, 1>(p, st);
  if (v == 2) return launch_env_real<REAL, 8, MODE, 2>(p, st);
  return launch_env_real<REAL, 8, MODE, 0>(p, st);
}

}  // namespace qmps_host
