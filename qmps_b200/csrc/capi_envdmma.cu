// qmps_b200: launcher of the D = 8 complex128 environment / energy kernel whose elimination runs on the FP64 tensor
// pipe (kernels_envdmma.cuh).  Its own translation unit: the fully unrolled kernel takes a while to compile.
#include <stdlib.h>
#include "api_common.cuh"
#include "kernels_envdmma.cuh"

namespace qmps_host {

template <int MODE, int DP>
static int launch_env_dmma_t(qmps::EnvParams p, cudaStream_t st) {
  using namespace qmps;
  const EdLayout L = ed_layout(p.d, p.nops, MODE == 1);
  auto kern = env_dmma_kernel<MODE, DP>;
  if (int rc = allow_smem(kern, L.total)) return rc;
  const int S = (MODE == 1 && p.nshift > 0) ? p.nshift : 1;
  int grid = 1;
  if (int rc = persistent_grid(kern, 64, L.total, p.N * S, &grid)) return rc;
  if (const char* cap = getenv("QMPS_ED_CTAS")) {          // experiment knob: resident CTAs per SM
    const int64_t g2 = (int64_t)sm_count() * atoi(cap);
    if (g2 >= 1 && g2 < grid) grid = (int)g2;
  }
  p.ws = nullptr; p.ws_stride = (size_t)sm_count();
  kern<<<grid, 64, L.total, st>>>(p);
  CK(cudaGetLastError());
  return 0;
}

int launch_env_dmma(int mode, qmps::EnvParams p, cudaStream_t st) {
  if (p.d == 2) return mode == 1 ? launch_env_dmma_t<1, 2>(p, st) : launch_env_dmma_t<0, 2>(p, st);
  return mode == 1 ? launch_env_dmma_t<1, 0>(p, st) : launch_env_dmma_t<0, 0>(p, st);
}

}  // namespace qmps_host
