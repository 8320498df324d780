// qmps_b200: generic-D launchers, REAL = double instantiation (own translation unit so
// the two precisions compile in parallel).
#include "api_common.cuh"
#include "launch_envreal.cuh"
#include "kernels_fp16x8.cuh"
#include "kernels_fp16s.cuh"
#include "kernels_fp16p.cuh"
#include "kernels_fp64w.cuh"
#include "kernels_fp64p.cuh"
#include "kernels_fpd2.cuh"

using namespace qmps;
namespace qmps_host {
namespace {
typedef double REAL;

template <int G, int MODE>
int launch_env_generic(EnvParams p, cudaStream_t st) {
  const int n = p.D * p.D;
  const int e_in_smem = n <= 64;
  const int block = G > 32 ? G : 128;
  const int gpc = G > 32 ? 1 : block / G;
  const EnvLayout<REAL> L = env_layout<REAL>(p.d, p.D, p.nops, G > 32 ? block : G, e_in_smem, (MODE == 1) || !p.assume_lc);
  const size_t smem = L.total * gpc;
  auto kern = env_generic_kernel<REAL, G, MODE>;
  if (int rc = allow_smem(kern, smem)) return rc;
  const int S = (MODE == 1 && p.nshift > 0) ? p.nshift : 1;
  const int64_t total = p.N * S;
  int grid = 1;
  if (int rc = persistent_grid(kern, block, smem, (total + gpc - 1) / gpc, &grid)) return rc;
  void* ws = nullptr;
  if (!e_in_smem) {
    p.ws_stride = (size_t)n * (n + 1);
    CK(malloc_async(&ws, sizeof(cx<REAL>) * p.ws_stride * grid, st));
    p.ws = ws;
  } else {
    p.ws = nullptr; p.ws_stride = 0;
  }
  kern<<<grid, block, smem, st>>>(p);
  CK(cudaGetLastError());
  if (ws) CK(cudaFreeAsync(ws, st));
  return 0;
}

template <int MODE> int dispatch_env(const EnvParams& p, cudaStream_t st) {
  if (env_real_applies(p)) return dispatch_env_real<REAL, MODE>(p, st);
  switch (group_for_n(p.D * p.D)) {
    case 4: return launch_env_generic<4, MODE>(p, st);
    case 16: return launch_env_generic<16, MODE>(p, st);
    case 128: return launch_env_generic<128, MODE>(p, st);
    default: return launch_env_generic<256, MODE>(p, st);
  }
}

// D = 4 eigenvalue-only fast path (kernels_fp16.cuh): half a warp per problem
template <int MINB> int launch_fp16_v(FpParams p, cudaStream_t st) {
  const Fp16Layout<REAL> L = fp16_layout<REAL>(p.d);
  const int block = 128, gpc = block / 16;
  const size_t smem = L.total * gpc;
  auto kern = fp16_kernel<REAL, MINB>;
  if (int rc = allow_smem(kern, smem)) return rc;
  int grid = 1;
  if (int rc = persistent_grid(kern, block, smem, (p.N + gpc - 1) / gpc, &grid)) return rc;
  p.ws = nullptr; p.ws_stride = 0;
  kern<<<grid, block, smem, st>>>(p);
  CK(cudaGetLastError());
  return 0;
}
// quarter-warp form (kernels_fp16x8.cuh): eight lanes per problem, CTAs of 64 threads
template <int MINB> int launch_fp16x8_v(FpParams p, cudaStream_t st) {
  const Fp16Layout<REAL> L = fp16_layout<REAL>(p.d);
  const int block = 64, gpc = block / 8;
  const size_t smem = L.total * gpc;
  auto kern = fp16x8_kernel<REAL, MINB>;
  if (int rc = allow_smem(kern, smem)) return rc;
  int grid = 1;
  if (int rc = persistent_grid(kern, block, smem, (p.N + gpc - 1) / gpc, &grid)) return rc;
  p.ws = nullptr; p.ws_stride = 0;
  kern<<<grid, block, smem, st>>>(p);
  CK(cudaGetLastError());
  return 0;
}
// shared-resident form (kernels_fp16s.cuh): half a warp per problem, matrix in a swizzled shared tile
int launch_fp16s(FpParams p, cudaStream_t st) {
  const Fp16sLayout<REAL> L = fp16s_layout<REAL>();
  const int block = 128, gpc = block / 16;
  const size_t smem = L.total * gpc;
  auto kern = fp16s_kernel<REAL>;
  if (int rc = allow_smem(kern, smem)) return rc;
  int grid = 1;
  if (int rc = persistent_grid(kern, block, smem, (p.N + gpc - 1) / gpc, &grid)) return rc;
  p.ws = nullptr; p.ws_stride = 0;
  kern<<<grid, block, smem, st>>>(p);
  CK(cudaGetLastError());
  return 0;
}
template <bool FASTRSQ, bool TRIM = false> int launch_fp16s8(FpParams p, cudaStream_t st) {
  const Fp16sLayout<REAL> L = fp16s_layout<REAL>();
  const int block = 64, gpc = block / 8;
  const size_t smem = L.total * gpc;
  auto kern = fp16s8_kernel<REAL, FASTRSQ, TRIM>;
  if (int rc = allow_smem(kern, smem)) return rc;
  int grid = 1;
  if (int rc = persistent_grid(kern, block, smem, (p.N + gpc - 1) / gpc, &grid)) return rc;
  p.ws = nullptr; p.ws_stride = 0;
  kern<<<grid, block, smem, st>>>(p);
  CK(cudaGetLastError());
  return 0;
}
// packed two-kernel form (kernels_fp16p.cuh): Hessenberg reduction -> packed workspace -> QR with 18 warps per SM
int launch_fp16p(FpParams p, cudaStream_t st) {
  const int64_t CH = 1 << 20;                                  // problems per launch pair: 2.4 KB (complex128) of workspace each
  const int64_t nws = p.N < CH ? p.N : CH;
  const int block = 64, gpc = block / 8;
  const size_t smem_h = fp16s_layout<REAL>().total * gpc, smem_q = fp16p_layout<REAL>().total * gpc;
  auto kh = fp16p8_kernel<REAL, 1>;
  auto kq = fp16p8_kernel<REAL, 2>;
  if (int rc = allow_smem(kh, smem_h)) return rc;
  if (int rc = allow_smem(kq, smem_q)) return rc;
  void* ws = nullptr;
  CK(malloc_async(&ws, sizeof(cx<REAL>) * (size_t)F16P_SIZE * nws, st));
  p.ws = ws; p.ws_stride = F16P_SIZE;
  for (int64_t off = 0; off < p.N; off += CH) {
    p.pid_offset = off; p.n_chunk = p.N - off < CH ? p.N - off : CH;
    int gh = 1, gq = 1;
    if (int rc = persistent_grid(kh, block, smem_h, (p.n_chunk + gpc - 1) / gpc, &gh)) { cudaFreeAsync(ws, st); return rc; }
    if (int rc = persistent_grid(kq, block, smem_q, (p.n_chunk + gpc - 1) / gpc, &gq)) { cudaFreeAsync(ws, st); return rc; }
    kh<<<gh, block, smem_h, st>>>(p);
    kq<<<gq, block, smem_q, st>>>(p);
  }
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(ws, st);
  CK(e);
  return 0;
}
// option fp16_fast: 6 = shared-resident form, 7 = its quarter-warp variant; 1 / 2 = half-warp form with the register budget of 3 / 4 CTAs per SM;
// 3 / 4 / 5 = quarter-warp form with the register budget of 4 / 5 / 6 CTAs (of 64 threads) per SM
int launch_fp16(const FpParams& p, cudaStream_t st) {
  switch (option_get(OPT_FP16_FAST)) {
    case 2: return launch_fp16_v<4>(p, st);
    case 3: return launch_fp16x8_v<4>(p, st);
    case 4: return launch_fp16x8_v<5>(p, st);
    case 5: return launch_fp16x8_v<6>(p, st);
    case 6: return launch_fp16s(p, st);
    case 7: return launch_fp16s8<false>(p, st);
    case 8: return launch_fp16s8<true>(p, st);
    case 9: return launch_fp16p(p, st);             // packed two-kernel form of 8
    case 10: return launch_fp16s8<true, true>(p, st); // 8 with the trimmed sweep bodies      // + branch-free reciprocal square root in the sweep body
    default: return launch_fp16_v<3>(p, st);
  }
}

// D = 8 eigenvalue-only path (kernels_fp64w.cuh): one warp per 64 x 64 map, matrix in a skewed shared tile
int launch_fp64w(FpParams p, cudaStream_t st) {
  const Fp64wLayout<REAL> L = fp64w_layout<REAL>();
  const size_t smem = L.total;
  auto kern = fp64w_kernel<REAL>;
  if (int rc = allow_smem(kern, smem)) return rc;
  int grid = 1;
  if (int rc = persistent_grid(kern, 32, smem, p.N, &grid)) return rc;
  p.ws = nullptr; p.ws_stride = 0;
  kern<<<grid, 32, smem, st>>>(p);
  CK(cudaGetLastError());
  return 0;
}

// packed two-kernel form (kernels_fp64p.cuh): Hessenberg reduction -> packed workspace -> QR with twice the warps per SM
int launch_fp64p(FpParams p, cudaStream_t st) {
  const int64_t CH = 32768;                                    // problems per launch pair: 34 KB (complex128) of workspace each, 1.1 GB at most
  const int64_t nws = p.N < CH ? p.N : CH;
  const size_t smem_h = fp64w_layout<REAL>().total + sizeof(cx<REAL>) * 2 * F64_N, smem_q = fp64p_layout<REAL>().total;
  auto kh = fp64p_hess_kernel<REAL>;
  auto kq = fp64p_qr_kernel<REAL>;
  if (int rc = allow_smem(kh, smem_h)) return rc;
  if (int rc = allow_smem(kq, smem_q)) return rc;
  void* ws = nullptr;
  CK(malloc_async(&ws, sizeof(cx<REAL>) * (size_t)F64P_SIZE * nws, st));
  p.ws = ws; p.ws_stride = F64P_SIZE;
  for (int64_t off = 0; off < p.N; off += CH) {
    p.pid_offset = off; p.n_chunk = p.N - off < CH ? p.N - off : CH;
    int gh = 1, gq = 1;
    if (int rc = persistent_grid(kh, 128, smem_h, p.n_chunk, &gh)) { cudaFreeAsync(ws, st); return rc; }
    if (int rc = persistent_grid(kq, 32, smem_q, p.n_chunk, &gq)) { cudaFreeAsync(ws, st); return rc; }
    kh<<<gh, 128, smem_h, st>>>(p);
    kq<<<gq, 32, smem_q, st>>>(p);
  }
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(ws, st);
  CK(e);
  return 0;
}

// D = 2 eigenvalue-only path (kernels_fpd2.cuh): one thread per problem, registers only
template <bool VEC> int launch_fp_d2_v(FpParams p, cudaStream_t st) {
  auto kern = fp_d2_kernel<REAL, VEC>;
  int grid = 1;
  if (int rc = persistent_grid(kern, 128, 0, (p.N + 127) / 128, &grid)) return rc;
  p.ws = nullptr; p.ws_stride = 0;
  kern<<<grid, 128, 0, st>>>(p);
  CK(cudaGetLastError());
  return 0;
}
int launch_fp_d2(const FpParams& p, cudaStream_t st) {
  return p.vec ? launch_fp_d2_v<true>(p, st) : launch_fp_d2_v<false>(p, st);
}

template <int G> int launch_fp(FpParams p, cudaStream_t st) {
  const int n = p.D * p.D;
  const int h_in_smem = n <= 64;
  int block = G > 32 ? G : 128;
  if (G <= 32 && option_get(OPT_FP_BLOCK) >= 32 && option_get(OPT_FP_BLOCK) <= 128) block = option_get(OPT_FP_BLOCK) & ~31;
  const int gpc = G > 32 ? 1 : block / G;
  const FpLayout<REAL> L = fp_layout<REAL>(p.D, h_in_smem);
  const size_t smem = L.total * gpc;
  auto kern = fixed_point_kernel<REAL, G>;
  if (int rc = allow_smem(kern, smem)) return rc;
  int grid = 1;
  if (int rc = persistent_grid(kern, block, smem, (p.N + gpc - 1) / gpc, &grid)) return rc;
  void* ws = nullptr;
  if (!h_in_smem) {
    p.ws_stride = (size_t)n * (n + 1);
    CK(malloc_async(&ws, sizeof(cx<REAL>) * p.ws_stride * grid, st));
    p.ws = ws;
  } else { p.ws = nullptr; p.ws_stride = 0; }
  kern<<<grid, block, smem, st>>>(p);
  CK(cudaGetLastError());
  if (ws) CK(cudaFreeAsync(ws, st));
  return 0;
}

template <int G>
int launch_ansatz(const GateOp* dops, int nops, int nq, int64_t N, int P, const double* theta, int full, void* out,
                  int coord, const double* dshifts, int nshift, cudaStream_t st) {
  const int R = 1 << nq, nc = full ? R : R / 2;
  const size_t per = (((size_t)R * nc * sizeof(cx<REAL>) + 15) & ~size_t(15)) + (((size_t)2 * nops * sizeof(REAL) + 15) & ~size_t(15));
  const int block = G > 32 ? G : 128;
  const int gpc = G > 32 ? 1 : block / G;
  const size_t smem = per * gpc;
  auto kern = ansatz_kernel<REAL, G>;
  if (int rc = allow_smem(kern, smem)) return rc;
  int grid = 1;
  if (int rc = persistent_grid(kern, block, smem, (N * (nshift > 0 ? nshift : 1) + gpc - 1) / gpc, &grid)) return rc;
  kern<<<grid, block, smem, st>>>(dops, nops, nq, N, P, theta, full, (cx<REAL>*)out, coord, dshifts, nshift);
  CK(cudaGetLastError());
  return 0;
}
}  // namespace

int env_generic_f64(const EnvParams& p, int mode, cudaStream_t st) {
  return mode == 0 ? dispatch_env<0>(p, st) : dispatch_env<1>(p, st);
}
int fp16_debug_f64(unsigned long long* out, int reset) {
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpyFromSymbol(out, g_fp16_dbg, 4 * sizeof(unsigned long long)));
  if (reset) {
    const unsigned long long z[4] = {0, 0, 0, 0};
    CK(cudaMemcpyToSymbol(g_fp16_dbg, z, sizeof(z)));
  }
  return 0;
}
int fixed_point_f64(const FpParams& p, cudaStream_t st) {
  if (p.D == 2 && p.d <= 16 && option_get(OPT_FP_D2)) return launch_fp_d2(p, st);
  if (p.D == 4 && p.vec == nullptr && p.d <= 16 && option_get(OPT_FP16_FAST)) return launch_fp16(p, st);
  if (p.D == 8 && p.vec == nullptr && p.d <= 16 && option_get(OPT_FP64_FAST) == 2) return launch_fp64p(p, st);
  if (p.D == 8 && p.vec == nullptr && p.d <= 16 && option_get(OPT_FP64_FAST)) return launch_fp64w(p, st);
  if (p.D == 4 && option_get(OPT_FP_GROUP) == 8) return launch_fp<8>(p, st);
  if (p.D == 4 && option_get(OPT_FP_GROUP) == 4) return launch_fp<4>(p, st);
  switch (group_for_n(p.D * p.D)) {
    case 4: return launch_fp<4>(p, st);
    case 16: return launch_fp<16>(p, st);
    case 128: return launch_fp<128>(p, st);
    default: return launch_fp<256>(p, st);
  }
}
int ansatz_f64(const GateOp* dops, int nops, int nq, int64_t N, int P, const double* theta, int full, void* out,
               int coord, const double* dshifts, int nshift, cudaStream_t st) {
  const int R = 1 << nq, elems = R * (full ? R : R / 2);
  return elems <= 8 ? launch_ansatz<4>(dops, nops, nq, N, P, theta, full, out, coord, dshifts, nshift, st)
       : elems <= 64 ? launch_ansatz<16>(dops, nops, nq, N, P, theta, full, out, coord, dshifts, nshift, st)
                     : launch_ansatz<32>(dops, nops, nq, N, P, theta, full, out, coord, dshifts, nshift, st);
}
}  // namespace qmps_host
