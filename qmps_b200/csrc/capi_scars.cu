// qmps_b200: launcher of the PXP scar-dynamics step cost (kernels_scars.cuh).
#include "api_common.cuh"
#include "kernels_scars.cuh"

namespace qmps_host {

int scars_cost_any(int64_t N, const double* params, int64_t NC, const double* current, const void* W, void* cost, void* eta,
                   int32_t* status, int dtype, cudaStream_t st) {
  using namespace qmps;
  if (N == 0) return 0;
  int64_t blocks = (N + 3) / 4;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (dtype == QMPS_C128)
    scars_cost_kernel<double><<<(unsigned)blocks, 128, 0, st>>>(N, params, NC, current, (const cx<double>*)W, (double*)cost, (cx<double>*)eta, status);
  else
    scars_cost_kernel<float><<<(unsigned)blocks, 128, 0, st>>>(N, params, NC, current, (const cx<float>*)W, (float*)cost, (cx<float>*)eta, status);
  CK(cudaGetLastError());
  return 0;
}

}  // namespace qmps_host
