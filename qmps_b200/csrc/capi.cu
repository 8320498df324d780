// qmps_b200 C ABI (include/qmps_b200.h): argument checks, dispatch, host-buffer
// pipelines.  No torch types here; Python reaches this through ctypes.
#include <math.h>
#include <stdio.h>
#include <mutex>

#include "api_common.cuh"
#include "kernels_misc.cuh"
#include "kernels_gemm.cuh"

using namespace qmps;
using namespace qmps_host;

static_assert(sizeof(qmps_gate_op) == sizeof(GateOp), "gate op layout");
static_assert(QMPS_G_YYPOW == G_YYPOW && QMPS_G_Z == G_Z, "gate codes");

namespace qmps_host {
std::string& last_error() { thread_local std::string e; return e; }
static int g_options[OPT_COUNT] = {1 /* d2_pdl */, 1 /* d2_ctas_per_sm (measured best: profiles/sweep_d2_r01.jsonl); 0 = occupancy */, 8 /* fp16_fast: D = 4 eigenvalue-only kernel; 0 generic, 1/2 half-warp registers, 3-5 quarter-warp registers, 6 shared-resident, 7 shared-resident quarter-warp, 8 the same with a branch-free reciprocal square root in the sweep body (measured best: profiles/exp_fp16_r02n.jsonl), 9 packed two-kernel form, 10 trimmed sweep bodies (both measured equal or slower: profiles/exp_fp16_r02q/r.jsonl) */, 1 /* env_real */, 1 /* tc_power: complex64 D % 64 == 0 on tcgen05 */, 1 /* tc_persistent */, 0, 0, -1 /* er_wide: auto */, 1 /* fp_d2: thread-per-problem D = 2 eigenvalue path */, 1 /* bw_thread: thread-per-candidate brick-wall cost */, 1 /* i8_power: complex128 D % 64 == 0 on tcgen05 kind::i8 (0: FP64 tensor pipe) */, -1 /* tc_presplit: complex64 slab images carry hi + lo planes (1), fp32 split in shared memory (0), by size (-1) */, 2 /* fp64_fast: D = 8 eigenvalue-only path; 0 generic CTA-per-problem kernel, 1 one warp per 64 x 64 map (kernels_fp64w.cuh), 2 packed two-kernel form (kernels_fp64p.cuh; measured best: profiles/exp_fp64w_r02n.jsonl) */};
std::unordered_map<LaunchKey, int, LaunchKeyHash>& occupancy_cache() { static std::unordered_map<LaunchKey, int, LaunchKeyHash> c; return c; }
std::mutex& occupancy_mutex() { static std::mutex m; return m; }
int option_get(int key) { return (key >= 0 && key < OPT_COUNT) ? g_options[key] : 0; }
}

namespace {

int env_exact_any(int d, int D, int64_t N, const void* in, int in_is_U, int assume_lc, void* eta, void* r,
                  void* C, int32_t* status, int dtype, cudaStream_t st) {
  if (N == 0) return 0;
  if (D == 2 && d == 2 && assume_lc) return env_d2(N, in, in_is_U, eta, r, C, status, dtype, st);
  EnvParams p;
  memset(&p, 0, sizeof(p));
  p.d = d; p.D = D; p.N = N; p.in = in; p.in_is_U = in_is_U; p.assume_lc = assume_lc;
  p.eta = eta; p.r = r; p.C = C; p.status = status; p.coord = -1;
  return dtype == QMPS_C128 ? env_generic_f64(p, 0, st) : env_generic_f32(p, 0, st);
}

template <typename T>
int tm_power_impl(int d, int D, int64_t N, const void* A, const void* B, void* r_io, int K, void* rayleigh,
                  cudaStream_t st) {
  if (N == 0) return 0;
  const size_t DD = (size_t)D * D;
  Scratch scratch(st);
  cx<T>* Tb = nullptr; cx<T>* Er = nullptr; T* invn = nullptr;
  CK(scratch.get(&Tb, sizeof(cx<T>) * N * d * DD));
  CK(scratch.get(&invn, sizeof(T) * N));
  cx<T>* r = (cx<T>*)r_io;
  const dim3 grid1((D + 31) / 32, (D + 31) / 32, (unsigned)(N * d)), grid2((D + 31) / 32, (D + 31) / 32, (unsigned)N);
  auto apply = [&](cx<T>* dst) {
    // T[b,s] = A[b,s] . r[b]   then   dst[b] = sum_s T[b,s] . B[b,s]^H
    zgemm_tile_kernel<T><<<grid1, 256, 0, st>>>(D, D, D, 1, (const cx<T>*)A, r, 0, d, Tb, nullptr);
    zgemm_tile_kernel<T><<<grid2, 256, 0, st>>>(D, D, D, d, Tb, (const cx<T>*)B, 1, 1, dst, nullptr);
  };
  for (int it = 0; it < K; ++it) {
    apply(r);
    inv_norm_kernel<T><<<(unsigned)N, 256, 0, st>>>((int64_t)DD, r, invn);
    scale_kernel<T><<<(unsigned)N, 256, 0, st>>>((int64_t)DD, r, invn);
  }
  if (rayleigh) {
    CK(scratch.get(&Er, sizeof(cx<T>) * N * DD));
    apply(Er);
    vdot_kernel<T><<<(unsigned)N, 256, 0, st>>>((int64_t)DD, r, Er, (cx<T>*)rayleigh);
  }
  CK(cudaGetLastError());
  return 0;
}


// complex128: DMMA tiles (kernels_gemm.cuh); two launches per application, norms carried as
// per-tile partial sums so that no separate normalisation pass touches r.
int tm_power_f64(int d, int D, int64_t N, const void* A, const void* B, void* r_io, int K, void* rayleigh,
                 cudaStream_t st) {
  if (N == 0) return 0;
  typedef cx<double> Z;
  const size_t DD = (size_t)D * D;
  const int tx = (D + ZG_TN - 1) / ZG_TN, ty = (D + ZG_TM - 1) / ZG_TM, tiles = tx * ty;
  Scratch scratch(st);
  Z* Tb = nullptr; Z* Er = nullptr; double* nrm = nullptr;
  CK(scratch.get(&Tb, sizeof(Z) * N * d * DD));
  CK(scratch.get(&nrm, sizeof(double) * N * tiles));
  if (int rc = allow_smem(zgemm_dmma_kernel<0>, ZG_SMEM_BYTES)) return rc;
  if (int rc = allow_smem(zgemm_dmma_kernel<1>, ZG_SMEM_BYTES)) return rc;
  Z* r = (Z*)r_io;
  const dim3 grid1(tx, ty, (unsigned)(N * d)), grid2(tx, ty, (unsigned)N);
  auto apply = [&](Z* dst, const double* norm_in, double* norm_out) {
    ZgParams p1;
    p1.M = D; p1.N = D; p1.K = D; p1.nsum = 1; p1.A = (const Z*)A; p1.B = r; p1.b_div = d; p1.C = Tb;
    p1.norm_in = norm_in; p1.n_in = tiles; p1.norm_out = nullptr;
    zgemm_dmma_kernel<0><<<grid1, 256, ZG_SMEM_BYTES, st>>>(p1);
    ZgParams p2;
    p2.M = D; p2.N = D; p2.K = D; p2.nsum = d; p2.A = Tb; p2.B = (const Z*)B; p2.b_div = 1; p2.C = dst;
    p2.norm_in = nullptr; p2.n_in = 0; p2.norm_out = norm_out;
    zgemm_dmma_kernel<1><<<grid2, 256, ZG_SMEM_BYTES, st>>>(p2);
  };
  for (int it = 0; it < K; ++it) apply(r, it == 0 ? nullptr : nrm, nrm);
  if (K > 0) zg_scale_kernel<<<(unsigned)N, 256, 0, st>>>((int64_t)DD, r, nrm, tiles);
  if (rayleigh) {
    CK(scratch.get(&Er, sizeof(Z) * N * DD));
    apply(Er, nullptr, nullptr);
    vdot_kernel<double><<<(unsigned)N, 256, 0, st>>>((int64_t)DD, r, Er, (Z*)rayleigh);
  }
  CK(cudaGetLastError());
  return 0;
}

// One UNNORMALISED application Y = sum_s A_s X B_s^dagger (the transfer-matrix apply itself: the building block of the
// Neumann / Krylov iterations of the large-D tangent vector).  complex128: the two DMMA launches of tm_power_f64;
// complex64: the SIMT tile kernel.
int tm_apply_f64(int d, int D, int64_t N, const void* A, const void* B, const void* X, void* Y, cudaStream_t st) {
  if (N == 0) return 0;
  typedef cx<double> Z;
  const size_t DD = (size_t)D * D;
  const int tx = (D + ZG_TN - 1) / ZG_TN, ty = (D + ZG_TM - 1) / ZG_TM;
  Scratch scratch(st);
  Z* Tb = nullptr;
  CK(scratch.get(&Tb, sizeof(Z) * N * d * DD));
  if (int rc = allow_smem(zgemm_dmma_kernel<0>, ZG_SMEM_BYTES)) return rc;
  if (int rc = allow_smem(zgemm_dmma_kernel<1>, ZG_SMEM_BYTES)) return rc;
  const dim3 grid1(tx, ty, (unsigned)(N * d)), grid2(tx, ty, (unsigned)N);
  ZgParams p1;
  p1.M = D; p1.N = D; p1.K = D; p1.nsum = 1; p1.A = (const Z*)A; p1.B = (const Z*)X; p1.b_div = d; p1.C = Tb;
  p1.norm_in = nullptr; p1.n_in = 0; p1.norm_out = nullptr;
  zgemm_dmma_kernel<0><<<grid1, 256, ZG_SMEM_BYTES, st>>>(p1);
  ZgParams p2;
  p2.M = D; p2.N = D; p2.K = D; p2.nsum = d; p2.A = Tb; p2.B = (const Z*)B; p2.b_div = 1; p2.C = (Z*)Y;
  p2.norm_in = nullptr; p2.n_in = 0; p2.norm_out = nullptr;
  zgemm_dmma_kernel<1><<<grid2, 256, ZG_SMEM_BYTES, st>>>(p2);
  CK(cudaGetLastError());
  return 0;
}
int tm_apply_f32(int d, int D, int64_t N, const void* A, const void* B, const void* X, void* Y, cudaStream_t st) {
  if (N == 0) return 0;
  typedef float T;
  const size_t DD = (size_t)D * D;
  Scratch scratch(st);
  cx<T>* Tb = nullptr;
  CK(scratch.get(&Tb, sizeof(cx<T>) * N * d * DD));
  const dim3 grid1((D + 31) / 32, (D + 31) / 32, (unsigned)(N * d)), grid2((D + 31) / 32, (D + 31) / 32, (unsigned)N);
  zgemm_tile_kernel<T><<<grid1, 256, 0, st>>>(D, D, D, 1, (const cx<T>*)A, (const cx<T>*)X, 0, d, Tb, nullptr);
  zgemm_tile_kernel<T><<<grid2, 256, 0, st>>>(D, D, D, d, Tb, (const cx<T>*)B, 1, 1, (cx<T>*)Y, nullptr);
  CK(cudaGetLastError());
  return 0;
}

}  // namespace

// ================================ exported C ABI =============================================
extern "C" {

const char* qmps_version(void) { return "qmps_b200 0.1.0 (sm_100a)"; }
const char* qmps_last_error(void) { return last_error().c_str(); }
int qmps_set_option(const char* name, int value) {
  if (!name) return fail(QMPS_ERR_ARG, "set_option: null name");
  if (!strcmp(name, "d2_pdl")) { g_options[OPT_D2_PDL] = value; return 0; }
  if (!strcmp(name, "d2_ctas_per_sm")) { g_options[OPT_D2_CTAS_PER_SM] = value; return 0; }
  if (!strcmp(name, "fp16_fast")) { g_options[OPT_FP16_FAST] = value; return 0; }
  if (!strcmp(name, "i8_power")) { g_options[OPT_I8_POWER] = value; return 0; }
  if (!strcmp(name, "tc_presplit")) { g_options[OPT_TC_PRESPLIT] = value; return 0; }
  if (!strcmp(name, "fp64_fast")) { g_options[OPT_FP64_FAST] = value; return 0; }
  if (!strcmp(name, "env_real")) { g_options[OPT_ENV_REAL] = value; return 0; }
  if (!strcmp(name, "tc_power")) { g_options[OPT_TC_POWER] = value; return 0; }
  if (!strcmp(name, "tc_persistent")) { g_options[OPT_TC_PERSISTENT] = value; return 0; }
  if (!strcmp(name, "fp_group")) { g_options[OPT_FP_GROUP] = value; return 0; }
  if (!strcmp(name, "fp_block")) { g_options[OPT_FP_BLOCK] = value; return 0; }
  if (!strcmp(name, "er_wide")) { g_options[OPT_ER_WIDE] = value; return 0; }
  if (!strcmp(name, "fp_d2")) { g_options[OPT_FP_D2] = value; return 0; }
  if (!strcmp(name, "bw_thread")) { g_options[OPT_BW_THREAD] = value; return 0; }
  return fail(QMPS_ERR_ARG, std::string("set_option: unknown option ") + name);
}
int qmps_debug_counters(unsigned long long* out4, int reset) {
  if (!out4) return fail(QMPS_ERR_ARG, "debug_counters: null output");
  return fp16_debug_f64(out4, reset);
}
int qmps_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int qmps_unitary_to_tensor(int D, int64_t N, const void* U, void* A, int dtype, void* stream) {
  if (D < 1 || N < 0 || (!U && N) || (!A && N)) return fail(QMPS_ERR_ARG, "unitary_to_tensor: bad arguments");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = N * 2 * D * D;
  const int grid = (int)((total + 255) / 256 < (int64_t)sm_count() * 16 ? (total + 255) / 256 : (int64_t)sm_count() * 16);
  if (dtype == QMPS_C128) u2t_kernel<double><<<grid, 256, 0, st>>>(D, N, (const cx<double>*)U, (cx<double>*)A);
  else u2t_kernel<float><<<grid, 256, 0, st>>>(D, N, (const cx<float>*)U, (cx<float>*)A);
  CK(cudaGetLastError());
  return 0;
}

int qmps_tensor_to_unitary(int d, int D, int64_t N, const void* A, void* U, int dtype, void* stream) {
  if (d < 1 || D < 1 || N < 0 || (!A && N) || (!U && N)) return fail(QMPS_ERR_ARG, "tensor_to_unitary: bad arguments");
  if (d * D > 64) return fail(QMPS_ERR_UNSUPPORTED, "tensor_to_unitary: d*D > 64 not supported");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int m = d * D;
  const size_t elems = (size_t)m * D * 2 + (size_t)m * m + D;
  const int grid = (int)(N < (int64_t)sm_count() * 4 ? N : (int64_t)sm_count() * 4);
  if (dtype == QMPS_C128) {
    const size_t smem = elems * sizeof(cx<double>);
    if (int rc = allow_smem(t2u_kernel<double>, smem)) return rc;
    t2u_kernel<double><<<grid, 128, smem, st>>>(d, D, N, (const cx<double>*)A, (cx<double>*)U);
  } else {
    const size_t smem = elems * sizeof(cx<float>);
    if (int rc = allow_smem(t2u_kernel<float>, smem)) return rc;
    t2u_kernel<float><<<grid, 128, smem, st>>>(d, D, N, (const cx<float>*)A, (cx<float>*)U);
  }
  CK(cudaGetLastError());
  return 0;
}

int qmps_environment_to_unitary(int n, int64_t N, const void* v, void* V, int dtype, void* stream) {
  if (n < 1 || N < 0 || (!v && N) || (!V && N)) return fail(QMPS_ERR_ARG, "environment_to_unitary: bad arguments");
  if (n > 4096) return fail(QMPS_ERR_UNSUPPORTED, "environment_to_unitary: n > 4096 not supported");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)(N < (int64_t)sm_count() * 4 ? N : (int64_t)sm_count() * 4);
  if (dtype == QMPS_C128) {
    const size_t smem = (size_t)n * sizeof(cx<double>);
    if (int rc = allow_smem(env2u_kernel<double>, smem)) return rc;
    env2u_kernel<double><<<grid, 256, smem, st>>>(n, N, (const cx<double>*)v, (cx<double>*)V);
  } else {
    const size_t smem = (size_t)n * sizeof(cx<float>);
    if (int rc = allow_smem(env2u_kernel<float>, smem)) return rc;
    env2u_kernel<float><<<grid, 256, smem, st>>>(n, N, (const cx<float>*)v, (cx<float>*)V);
  }
  CK(cudaGetLastError());
  return 0;
}

int qmps_env_exact(int d, int D, int64_t N, const void* in, int in_is_full_U, int assume_left_canonical,
                   void* eta, void* r, void* C, int32_t* status, int dtype, void* stream) {
  if (N < 0 || (!in && N)) return fail(QMPS_ERR_ARG, "env_exact: bad arguments");
  if (!is_pow2(D) || D < 1 || D > 16) return fail(QMPS_ERR_UNSUPPORTED, "env_exact: D must be 1, 2, 4, 8 or 16");
  if (d < 1 || d > 4) return fail(QMPS_ERR_UNSUPPORTED, "env_exact: d must be 1..4");
  if (in_is_full_U && d != 2) return fail(QMPS_ERR_ARG, "env_exact: unitary input implies d = 2");
  if (dtype != QMPS_C128 && dtype != QMPS_C64) return fail(QMPS_ERR_ARG, "env_exact: bad dtype");
  return env_exact_any(d, D, N, in, in_is_full_U, assume_left_canonical, eta, r, C, status, dtype, (cudaStream_t)stream);
}

int qmps_fixed_point(int d, int D, int64_t NA, const void* A, int64_t NB, const void* B, int pair_mode, int left,
                     void* eta, void* vec, void* cost, void* echo, void* fid, int32_t* status, int dtype,
                     void* stream) {
  return qmps_fixed_point_ex(d, D, NA, A, NB, B, pair_mode, left, QMPS_GAUGE_ZGEEV, eta, vec, cost, echo, fid, status,
                             dtype, stream);
}

int qmps_fixed_point_ex(int d, int D, int64_t NA, const void* A, int64_t NB, const void* B, int pair_mode, int left,
                        int vec_gauge, void* eta, void* vec, void* cost, void* echo, void* fid, int32_t* status,
                        int dtype, void* stream) {
  if (vec_gauge != QMPS_GAUGE_TRACE && vec_gauge != QMPS_GAUGE_ZGEEV) return fail(QMPS_ERR_ARG, "fixed_point: bad vec_gauge");
  if (NA < 0 || NB < 0 || (!A && NA) || (!B && NB)) return fail(QMPS_ERR_ARG, "fixed_point: bad arguments");
  if (D < 1 || D > 16) return fail(QMPS_ERR_UNSUPPORTED, "fixed_point: D must be 1..16");
  if (d < 1 || d > 16) return fail(QMPS_ERR_UNSUPPORTED, "fixed_point: d must be 1..16");
  if (pair_mode == 0 && !(NA == NB || NA == 1 || NB == 1)) return fail(QMPS_ERR_ARG, "fixed_point: batch sizes do not broadcast");
  if (NA == 0 || NB == 0) return 0;
  FpParams p;
  memset(&p, 0, sizeof(p));
  p.d = d; p.D = D; p.NA = NA; p.NB = NB; p.A = A; p.B = B; p.pair_mode = pair_mode; p.left = left;
  if (pair_mode < 0 || pair_mode > 2) return fail(QMPS_ERR_ARG, "fixed_point: pair_mode must be 0, 1 or 2");
  p.N = pair_mode != 0 ? NA * NB : (NA > NB ? NA : NB);
  p.vec_gauge = vec_gauge;
  p.eta = eta; p.vec = vec; p.cost = cost; p.echo = echo; p.fid = fid; p.status = status;
  if (dtype == QMPS_C128) return fixed_point_f64(p, (cudaStream_t)stream);
  if (dtype == QMPS_C64) return fixed_point_f32(p, (cudaStream_t)stream);
  return fail(QMPS_ERR_ARG, "fixed_point: bad dtype");
}

int qmps_merge(int d1, int d2, int D, int64_t NA, const void* A, int64_t NB, const void* B, int64_t NW,
               const void* W, void* M, int dtype, void* stream) {
  if (NA < 1 || NB < 1 || !A || !B || !M || (W && NW < 1)) return fail(QMPS_ERR_ARG, "merge: bad arguments");
  if (d1 < 1 || d2 < 1 || D < 1 || d1 * d2 * D * D > 8192) return fail(QMPS_ERR_UNSUPPORTED, "merge: block too large");
  int64_t N = NA > NB ? NA : NB;
  if (W && NW > N) N = NW;
  if ((NA != 1 && NA != N) || (NB != 1 && NB != N) || (W && NW != 1 && NW != N))
    return fail(QMPS_ERR_ARG, "merge: batch sizes do not broadcast (each of NA, NB, NW must be 1 or N)");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)(N < (int64_t)sm_count() * 8 ? N : (int64_t)sm_count() * 8);
  if (dtype == QMPS_C128) {
    const size_t smem = W ? sizeof(cx<double>) * d1 * d2 * D * D : 0;
    if (int rc = allow_smem(merge_kernel<double>, smem)) return rc;
    merge_kernel<double><<<grid, 128, smem, st>>>(d1, d2, D, NA, (const cx<double>*)A, NB, (const cx<double>*)B, NW,
                                                  (const cx<double>*)W, N, (cx<double>*)M);
  } else {
    const size_t smem = W ? sizeof(cx<float>) * d1 * d2 * D * D : 0;
    if (int rc = allow_smem(merge_kernel<float>, smem)) return rc;
    merge_kernel<float><<<grid, 128, smem, st>>>(d1, d2, D, NA, (const cx<float>*)A, NB, (const cx<float>*)B, NW,
                                                 (const cx<float>*)W, N, (cx<float>*)M);
  }
  CK(cudaGetLastError());
  return 0;
}

int qmps_ansatz(const qmps_gate_op* ops, int nops, int nq, int64_t N, int P, const double* theta, int full_unitary,
                void* out, int dtype, void* stream) {
  if (!ops || nops < 0 || nq < 1 || nq > 5 || N < 0 || (!theta && N && P) || (!out && N))
    return fail(QMPS_ERR_ARG, "ansatz: bad arguments (nq must be 1..5)");
  for (int k = 0; k < nops; ++k)
    if (ops[k].param >= P || ops[k].q0 < 0 || ops[k].q0 >= nq || ops[k].q1 < 0 || ops[k].q1 >= nq)
      return fail(QMPS_ERR_ARG, "ansatz: gate references a parameter or qubit out of range");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  GateOp* dops = nullptr;
  if (int rc = to_device_async((const GateOp*)ops, (size_t)nops, &dops, st)) return rc;
  int rc = dtype == QMPS_C128 ? ansatz_f64(dops, nops, nq, N, P, theta, full_unitary, out, -1, nullptr, 0, st)
                              : ansatz_f32(dops, nops, nq, N, P, theta, full_unitary, out, -1, nullptr, 0, st);
  if (dops) CK(cudaFreeAsync(dops, st));
  return rc;
}

int qmps_energy_theta(const qmps_gate_op* ops, int nops, int nq, int64_t N, int P, const double* theta,
                      const void* hmat, int coord, const double* shifts, int nshift, void* energy,
                      int32_t* status, int dtype, void* stream) {
  if (!ops || nops < 1 || nq < 2 || nq > 5 || N < 0 || !theta || !hmat || !energy || nshift < 0 || (nshift && !shifts))
    return fail(QMPS_ERR_ARG, "energy_theta: bad arguments (nq must be 2..5)");
  if (nshift && (coord < 0 || coord >= P)) return fail(QMPS_ERR_ARG, "energy_theta: coord out of range");
  for (int k = 0; k < nops; ++k)
    if (ops[k].param >= P || ops[k].q0 < 0 || ops[k].q0 >= nq || ops[k].q1 < 0 || ops[k].q1 >= nq)
      return fail(QMPS_ERR_ARG, "energy_theta: gate references a parameter or qubit out of range");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  GateOp* dops = nullptr;
  double* dsh = nullptr;
  if (int rc = to_device_async((const GateOp*)ops, (size_t)nops, &dops, st)) return rc;
  if (int rc = to_device_async(shifts, (size_t)nshift, &dsh, st)) { if (dops) cudaFreeAsync(dops, st); return rc; }
  int rc = 0;
  const int D = 1 << (nq - 1);
  if (D == 2 && nops <= d2_max_ops()) {
    rc = energy_d2_theta(dops, nops, N, P, theta, hmat, coord, dsh, nshift, energy, status, dtype, st);
  } else if (D == 8 && dtype == QMPS_C128 && option_get(OPT_ENV_REAL) && (option_get(OPT_ER_WIDE) < 0 || option_get(OPT_ER_WIDE) == 3)) {
    // phase split (measured: profiles/exp_split_r02*.json): the ansatz runs as its own high-occupancy kernel and leaves
    // the tensors in HBM (2 KB per evaluation), the 200-register solve kernel starts from them.  Chunked so that the
    // scratch stays below 512 MiB.
    const int S = nshift > 0 ? nshift : 1;
    const int64_t chunk = ((int64_t)1 << 18) / S;
    Scratch scratch(st);
    cx<double>* At = nullptr;
    const int64_t first = N < chunk ? N : chunk;
    if (scratch.get(&At, sizeof(cx<double>) * (size_t)first * S * 128) != cudaSuccess) rc = fail(QMPS_ERR_CUDA, "energy_theta: scratch allocation failed");
    for (int64_t off = 0; off < N && !rc; off += chunk) {
      const int64_t cnt = (N - off < chunk) ? (N - off) : chunk;
      rc = ansatz_f64(dops, nops, nq, cnt, P, theta + off * P, 0, At, coord, dsh, nshift, st);
      if (rc) break;
      EnvParams p;
      memset(&p, 0, sizeof(p));
      p.d = 2; p.D = D; p.N = cnt * S; p.in = At; p.assume_lc = 1; p.status = status ? status + off * S : nullptr; p.coord = -1;
      p.hmat = hmat; p.energy = (double*)energy + off * S; p.two_site = 0;
      rc = env_generic_f64(p, 1, st);
    }
  } else {
    EnvParams p;
    memset(&p, 0, sizeof(p));
    p.d = 2; p.D = D; p.N = N; p.assume_lc = 1; p.status = status;
    p.ops = dops; p.nops = nops; p.nq = nq; p.P = P; p.theta = theta; p.coord = coord; p.shifts = dsh; p.nshift = nshift;
    p.hmat = hmat; p.energy = energy; p.two_site = 0;
    rc = dtype == QMPS_C128 ? env_generic_f64(p, 1, st) : env_generic_f32(p, 1, st);
  }
  if (!rc) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) rc = fail(QMPS_ERR_CUDA, cudaGetErrorString(e)); }
  if (dops) cudaFreeAsync(dops, st);
  if (dsh) cudaFreeAsync(dsh, st);
  return rc;
}

int qmps_energy_tensor(int D, int64_t N, const void* in, int two_site, const void* hmat, void* energy,
                       int32_t* status, int dtype, void* stream) {
  if (N < 0 || (!in && N) || !hmat || (!energy && N)) return fail(QMPS_ERR_ARG, "energy_tensor: bad arguments");
  if (!is_pow2(D) || D > 16) return fail(QMPS_ERR_UNSUPPORTED, "energy_tensor: D must be 1, 2, 4, 8 or 16");
  if (N == 0) return 0;
  EnvParams p;
  memset(&p, 0, sizeof(p));
  p.d = two_site ? 4 : 2; p.D = D; p.N = N; p.in = in; p.assume_lc = 1; p.status = status; p.coord = -1;
  p.hmat = hmat; p.energy = energy; p.two_site = two_site;
  return dtype == QMPS_C128 ? env_generic_f64(p, 1, (cudaStream_t)stream) : env_generic_f32(p, 1, (cudaStream_t)stream);
}

int qmps_rotosolve_fit(int64_t N, int nshift, const double* cost, double* theta_star, double* fit, double* theta_io,
                       int P, int coord, void* stream) {
  if (N < 0 || (nshift != 3 && nshift != 6) || (!cost && N)) return fail(QMPS_ERR_ARG, "rotosolve_fit: nshift must be 3 or 6");
  if (theta_io && (coord < 0 || coord >= P)) return fail(QMPS_ERR_ARG, "rotosolve_fit: coord out of range");
  if (N == 0) return 0;
  const int grid = (int)((N + 127) / 128 < (int64_t)sm_count() * 8 ? (N + 127) / 128 : (int64_t)sm_count() * 8);
  rotosolve_fit_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(N, nshift, cost, theta_star, fit, theta_io, P, coord);
  CK(cudaGetLastError());
  return 0;
}

int qmps_tm_apply(int d, int D, int64_t N, const void* A, const void* B, const void* X, void* Y, int dtype, void* stream) {
  if (d < 1 || D < 1 || N < 0 || (N && (!A || !B || !X || !Y))) return fail(QMPS_ERR_ARG, "tm_apply: bad arguments");
  if (dtype != QMPS_C128 && dtype != QMPS_C64) return fail(QMPS_ERR_ARG, "tm_apply: bad dtype");
  if (X == Y) return fail(QMPS_ERR_ARG, "tm_apply: X and Y must not alias");
  const int64_t chunk = 65535 / d;
  if (chunk < 1) return fail(QMPS_ERR_UNSUPPORTED, "tm_apply: d > 65535");
  const size_t csz = dtype == QMPS_C128 ? 16 : 8, DD = (size_t)D * D;
  for (int64_t n0 = 0; n0 < N; n0 += chunk) {
    const int64_t n = (N - n0 < chunk) ? N - n0 : chunk;
    const char* a = (const char*)A + csz * (size_t)n0 * d * DD;
    const char* b = (const char*)B + csz * (size_t)n0 * d * DD;
    const char* x = (const char*)X + csz * (size_t)n0 * DD;
    char* y = (char*)Y + csz * (size_t)n0 * DD;
    const int rc = dtype == QMPS_C128 ? tm_apply_f64(d, D, n, a, b, x, y, (cudaStream_t)stream)
                                      : tm_apply_f32(d, D, n, a, b, x, y, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return 0;
}

int qmps_tm_power(int d, int D, int64_t N, const void* A, const void* B, void* r_io, int K, void* rayleigh,
                  int dtype, void* stream) {
  if (d < 1 || D < 1 || N < 0 || K < 0 || (N && (!A || !B || !r_io))) return fail(QMPS_ERR_ARG, "tm_power: bad arguments");
  if (dtype != QMPS_C128 && dtype != QMPS_C64) return fail(QMPS_ERR_ARG, "tm_power: bad dtype");
  // problems are independent: batches beyond the grid z-limit (N * d <= 65535) run as consecutive
  // chunks on the same stream
  const int64_t chunk = 65535 / d;
  if (chunk < 1) return fail(QMPS_ERR_UNSUPPORTED, "tm_power: d > 65535");
  const size_t csz = dtype == QMPS_C128 ? 16 : 8, DD = (size_t)D * D;
  for (int64_t n0 = 0; n0 < N; n0 += chunk) {
    const int64_t n = (N - n0 < chunk) ? N - n0 : chunk;
    const char* a = (const char*)A + csz * (size_t)n0 * d * DD;
    const char* b = (const char*)B + csz * (size_t)n0 * d * DD;
    char* r = (char*)r_io + csz * (size_t)n0 * DD;
    void* ray = rayleigh ? (void*)((char*)rayleigh + csz * (size_t)n0) : nullptr;
    int rc;
    if (dtype == QMPS_C128 && tm_power_i8_applies(d, D, n)) rc = tm_power_i8(d, D, n, a, b, r, K, ray, (cudaStream_t)stream);
    else if (dtype == QMPS_C128) rc = tm_power_f64(d, D, n, a, b, r, K, ray, (cudaStream_t)stream);
    else if (tm_power_tc_applies(d, D, n)) rc = tm_power_tc(d, D, n, a, b, r, K, ray, (cudaStream_t)stream);
    else rc = tm_power_impl<float>(d, D, n, a, b, r, K, ray, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return 0;
}

int qmps_cgemm_c64_tc(int64_t batch, int nsum, int M, int N, int K, const void* X, const void* Y, int conj_y, void* C,
                      void* stream) {
  if (batch < 0 || nsum < 1 || (batch && (!X || !Y || !C))) return fail(QMPS_ERR_ARG, "cgemm_c64_tc: bad arguments");
  return cgemm_c64_tc(batch, nsum, M, N, K, X, Y, conj_y, C, (cudaStream_t)stream);
}

int qmps_zgemm_c128_i8(int64_t batch, int M, int N, int K, const void* X, const void* Y, int conj_y, void* C, void* stream) {
  if (batch < 0 || (batch && (!X || !Y || !C))) return fail(QMPS_ERR_ARG, "zgemm_c128_i8: bad arguments");
  return zgemm_c128_i8(batch, M, N, K, X, Y, conj_y, C, (cudaStream_t)stream);
}

int qmps_argmin(int64_t N, const double* cost, int64_t index_offset, double* best_cost, int64_t* best_index,
                void* stream) {
  if (N < 0 || (!cost && N) || !best_cost || !best_index) return fail(QMPS_ERR_ARG, "argmin: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int grid = (int)((N + 255) / 256);
  if (grid > sm_count() * 4) grid = sm_count() * 4;
  if (grid < 1) grid = 1;
  Scratch scratch(st);
  double* bc = nullptr; int64_t* bi = nullptr; unsigned int* ctr = nullptr;
  CK(scratch.get(&bc, sizeof(double) * grid));
  CK(scratch.get(&bi, sizeof(int64_t) * grid));
  CK(scratch.get(&ctr, sizeof(unsigned int)));
  CK(cudaMemsetAsync(ctr, 0, sizeof(unsigned int), st));
  argmin_kernel<<<grid, 256, 0, st>>>(N, cost, index_offset, bc, bi, ctr, best_cost, best_index);
  CK(cudaGetLastError());
  return 0;
}

int qmps_loschmidt_rate(int64_t NT, const double* t, double g0, double g1, double* out, void* stream) {
  if (NT < 0 || (NT && (!t || !out))) return fail(QMPS_ERR_ARG, "loschmidt_rate: bad arguments");
  if (NT == 0) return 0;
  int grid = (int)((NT + 3) / 4);
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  loschmidt_rate_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(NT, t, g0, g1, out);
  CK(cudaGetLastError());
  return 0;
}

// ---- host-buffer pipeline: chunked H2D -> solve -> D2H on rotating streams ---------------------
int qmps_env_exact_host(int d, int D, int64_t N, const void* in, int in_is_full_U, int assume_left_canonical,
                        void* eta, void* r, void* C, int32_t* status, int dtype, int device) {
  if (N < 0 || (!in && N)) return fail(QMPS_ERR_ARG, "env_exact_host: bad arguments");
  if (!is_pow2(D) || D > 16 || d < 1 || d > 4) return fail(QMPS_ERR_UNSUPPORTED, "env_exact_host: unsupported d/D");
  if (in_is_full_U && d != 2) return fail(QMPS_ERR_ARG, "env_exact_host: unitary input implies d = 2");
  if (N == 0) return 0;
  CK(cudaSetDevice(device));
  const size_t csz = dtype == QMPS_C128 ? 16 : 8;
  const size_t in_per = csz * (in_is_full_U ? (size_t)4 * D * D : (size_t)d * D * D);
  const size_t mat_per = csz * (size_t)D * D;
  const size_t out_per = (eta ? csz : 0) + (r ? mat_per : 0) + (C ? mat_per : 0) + (status ? 4 : 0);
  // chunk so that one chunk moves ~16 MiB in; at least 3 chunks in flight for overlap
  int64_t chunk = (int64_t)((16u << 20) / in_per);
  if (chunk < 1) chunk = 1;
  if (chunk > N) chunk = N;
  const int NS = 3;
  struct Slot { cudaStream_t st; char* din; char* dout; };
  // staging slots are per device; callers on different devices do not serialise on each other
  static std::mutex mu[64];
  static Slot slots[64][NS];
  static size_t cap_in[64] = {0}, cap_out[64] = {0};
  if (device < 0 || device >= 64) return fail(QMPS_ERR_ARG, "env_exact_host: bad device");
  std::lock_guard<std::mutex> lock(mu[device]);
  Slot* sl = slots[device];
  const size_t need_in = (size_t)chunk * in_per, need_out = (size_t)chunk * (out_per ? out_per : 1);
  if (cap_in[device] < need_in || cap_out[device] < need_out) {
    cap_in[device] = cap_out[device] = 0;      // a failed re-allocation must not leave stale capacities / pointers
    for (int k = 0; k < NS; ++k) {
      if (sl[k].din) { cudaFree(sl[k].din); sl[k].din = nullptr; }
      if (sl[k].dout) { cudaFree(sl[k].dout); sl[k].dout = nullptr; }
      if (!sl[k].st) CK(cudaStreamCreateWithFlags(&sl[k].st, cudaStreamNonBlocking));
      CK(cudaMalloc((void**)&sl[k].din, need_in));
      CK(cudaMalloc((void**)&sl[k].dout, need_out));
    }
    cap_in[device] = need_in; cap_out[device] = need_out;
  }
  int rc = 0;
  int k = 0;
  for (int64_t off = 0; off < N && !rc; off += chunk, k = (k + 1) % NS) {
    const int64_t cnt = (N - off < chunk) ? (N - off) : chunk;
    Slot& s = sl[k];
    CK(cudaMemcpyAsync(s.din, (const char*)in + (size_t)off * in_per, (size_t)cnt * in_per, cudaMemcpyHostToDevice, s.st));
    char* o = s.dout;
    char* d_eta = nullptr; char* d_r = nullptr; char* d_C = nullptr; char* d_st = nullptr;
    if (eta) { d_eta = o; o += (size_t)cnt * csz; }
    if (r) { d_r = o; o += (size_t)cnt * mat_per; }
    if (C) { d_C = o; o += (size_t)cnt * mat_per; }
    if (status) { d_st = o; }
    rc = env_exact_any(d, D, cnt, s.din, in_is_full_U, assume_left_canonical, d_eta, d_r, d_C, (int32_t*)d_st, dtype, s.st);
    if (rc) break;
    if (eta) CK(cudaMemcpyAsync((char*)eta + (size_t)off * csz, d_eta, (size_t)cnt * csz, cudaMemcpyDeviceToHost, s.st));
    if (r) CK(cudaMemcpyAsync((char*)r + (size_t)off * mat_per, d_r, (size_t)cnt * mat_per, cudaMemcpyDeviceToHost, s.st));
    if (C) CK(cudaMemcpyAsync((char*)C + (size_t)off * mat_per, d_C, (size_t)cnt * mat_per, cudaMemcpyDeviceToHost, s.st));
    if (status) CK(cudaMemcpyAsync((char*)status + (size_t)off * 4, d_st, (size_t)cnt * 4, cudaMemcpyDeviceToHost, s.st));
  }
  for (int q = 0; q < NS; ++q) if (sl[q].st) CK(cudaStreamSynchronize(sl[q].st));
  return rc;
}

// ---- the scalar drop-in in ONE call (qmps/tools.py:176-182 get_env_exact): U -> V on host buffers -----------------
// H2D of the unitaries, environment solve (eta, r, Cholesky C), environment_to_unitary(C), D2H of V and the status
// words, one stream synchronisation.  The Python mirror used to make six API calls with three host syncs per solve.
int qmps_get_env_exact_host(int D, int64_t N, const void* U, void* V, int32_t* status, int dtype, int device) {
  if (N < 0 || (N && (!U || !V))) return fail(QMPS_ERR_ARG, "get_env_exact_host: bad arguments");
  if (!is_pow2(D) || D > 16) return fail(QMPS_ERR_UNSUPPORTED, "get_env_exact_host: unsupported D");
  if (device < 0 || device >= 64) return fail(QMPS_ERR_ARG, "get_env_exact_host: bad device");
  if (N == 0) return 0;
  CK(cudaSetDevice(device));
  const size_t csz = dtype == QMPS_C128 ? 16 : 8;
  const int n = D * D;
  const size_t u_per = csz * 4 * n, c_per = csz * n, v_per = csz * (size_t)n * n;
  struct Slot { cudaStream_t st; char* buf; size_t cap; };
  static std::mutex mu[64];
  static Slot slots[64];
  std::lock_guard<std::mutex> lock(mu[device]);
  Slot& s = slots[device];
  const size_t need = (size_t)N * (u_per + c_per + v_per + 16);
  if (!s.st) CK(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
  if (s.cap < need) {
    if (s.buf) { cudaFree(s.buf); s.buf = nullptr; s.cap = 0; }
    CK(cudaMalloc((void**)&s.buf, need));
    s.cap = need;
  }
  char* dU = s.buf; char* dC = dU + (size_t)N * u_per; char* dV = dC + (size_t)N * c_per; char* dS = dV + (size_t)N * v_per;
  CK(cudaMemcpyAsync(dU, U, (size_t)N * u_per, cudaMemcpyHostToDevice, s.st));
  if (int rc = env_exact_any(2, D, N, dU, 1, 1, nullptr, nullptr, dC, (int32_t*)dS, dtype, s.st)) return rc;
  if (int rc = qmps_environment_to_unitary(n, N, dC, dV, dtype, (void*)s.st)) return rc;
  CK(cudaMemcpyAsync(V, dV, (size_t)N * v_per, cudaMemcpyDeviceToHost, s.st));
  if (status) CK(cudaMemcpyAsync(status, dS, (size_t)N * 4, cudaMemcpyDeviceToHost, s.st));
  CK(cudaStreamSynchronize(s.st));
  return 0;
}

// ---- packed D = 2 outputs: 64 B per solve instead of 148 B (the host link is what bounds the host-buffer path) ------
int qmps_env_exact_packed(int64_t N, const void* in, int in_is_full_U, void* packed, void* stream) {
  if (N < 0 || (N && (!in || !packed))) return fail(QMPS_ERR_ARG, "env_exact_packed: bad arguments");
  return env_d2_packed(N, in, in_is_full_U, packed, (cudaStream_t)stream);
}
int qmps_env_exact_packed_host(int64_t N, const void* in, int in_is_full_U, void* packed, int device) {
  if (N < 0 || (N && (!in || !packed))) return fail(QMPS_ERR_ARG, "env_exact_packed_host: bad arguments");
  if (device < 0 || device >= 64) return fail(QMPS_ERR_ARG, "env_exact_packed_host: bad device");
  if (N == 0) return 0;
  CK(cudaSetDevice(device));
  const size_t in_per = in_is_full_U ? 256 : 128, out_per = 64;
  int64_t chunk = (int64_t)((16u << 20) / in_per);
  if (chunk > N) chunk = N;
  const int NS = 3;
  struct Slot { cudaStream_t st; char* din; char* dout; };
  static std::mutex mu[64];
  static Slot slots[64][NS];
  static size_t cap[64] = {0};
  std::lock_guard<std::mutex> lock(mu[device]);
  Slot* sl = slots[device];
  if (cap[device] < (size_t)chunk) {
    cap[device] = 0;
    for (int k = 0; k < NS; ++k) {
      if (sl[k].din) { cudaFree(sl[k].din); sl[k].din = nullptr; }
      if (sl[k].dout) { cudaFree(sl[k].dout); sl[k].dout = nullptr; }
      if (!sl[k].st) CK(cudaStreamCreateWithFlags(&sl[k].st, cudaStreamNonBlocking));
      CK(cudaMalloc((void**)&sl[k].din, (size_t)chunk * 256));
      CK(cudaMalloc((void**)&sl[k].dout, (size_t)chunk * out_per));
    }
    cap[device] = (size_t)chunk;
  }
  int rc = 0, k = 0;
  for (int64_t off = 0; off < N && !rc; off += chunk, k = (k + 1) % NS) {
    const int64_t cnt = (N - off < chunk) ? (N - off) : chunk;
    Slot& s = sl[k];
    CK(cudaMemcpyAsync(s.din, (const char*)in + (size_t)off * in_per, (size_t)cnt * in_per, cudaMemcpyHostToDevice, s.st));
    rc = env_d2_packed(cnt, s.din, in_is_full_U, s.dout, s.st);
    if (rc) break;
    CK(cudaMemcpyAsync((char*)packed + (size_t)off * out_per, s.dout, (size_t)cnt * out_per, cudaMemcpyDeviceToHost, s.st));
  }
  for (int q = 0; q < NS; ++q) if (sl[q].st) CK(cudaStreamSynchronize(sl[q].st));
  return rc;
}

}  // extern "C"
