// qmps_b200 C ABI, two-layer brick-wall iMPS family (SURVEY 8(f)-4;
// new_tdvp/ClassicalTDVPStripped.py:228-533, 777-790).  Own translation unit.
#include "api_common.cuh"
#include "kernels_bw.cuh"

using namespace qmps;
using namespace qmps_host;

namespace {

constexpr int BW_G = 16;          // lanes per problem (two problems per warp)
constexpr int BW_BLOCK = 128;

template <typename T> int launch_bw(const BwParams& p, cudaStream_t st) {
  const int gpc = BW_BLOCK / BW_G;
  const size_t smem = bw_work_bytes<T>(BW_G) * gpc;
  auto kern = bw_kernel<T, BW_G>;
  if (int rc = allow_smem(kern, smem)) return rc;
  int grid = 1;
  if (int rc = persistent_grid(kern, BW_BLOCK, smem, (p.N + gpc - 1) / gpc, &grid)) return rc;
  kern<<<grid, BW_BLOCK, smem, st>>>(p);
  CK(cudaGetLastError());
  return 0;
}

int run_bw(const BwParams& p, int dtype, void* stream) {
  return dtype == QMPS_C128 ? launch_bw<double>(p, (cudaStream_t)stream) : launch_bw<float>(p, (cudaStream_t)stream);
}

bool bcast_ok(int64_t n, int64_t N) { return n == 1 || n == N; }

// one ket state, one W, many candidates: thread-per-candidate kernel behind a one-warp precompute of chi
template <typename T> int launch_bw_cost_thread(const BwParams& p, cudaStream_t st) {
  cx<T>* chi = nullptr;
  CK(malloc_async((void**)&chi, sizeof(cx<T>) * 64, st));
  bw_chi_kernel<T><<<1, 32, 0, st>>>((const cx<T>*)p.U1, (const cx<T>*)p.U2, (const cx<T>*)p.W, chi);
  auto kern = bw_cost_thread_kernel<T>;
  int grid = 1;
  if (int rc = persistent_grid(kern, 128, 0, (p.N + 127) / 128, &grid)) return rc;
  kern<<<grid, 128, 0, st>>>(p, chi);
  CK(cudaGetLastError());
  CK(cudaFreeAsync(chi, st));
  return 0;
}

template <typename T> int launch_bw_env_thread(const BwParams& p, cudaStream_t st) {
  auto kern = bw_env_thread_kernel<T>;
  int grid = 1;
  if (int rc = persistent_grid(kern, 128, 0, (p.N + 127) / 128, &grid)) return rc;
  kern<<<grid, 128, 0, st>>>(p);
  CK(cudaGetLastError());
  return 0;
}

int check_common(const char* who, int64_t N, int64_t NK, const void* U1, const void* U2, int dtype) {
  if (N < 0) return fail(QMPS_ERR_ARG, std::string(who) + ": negative batch");
  if (dtype != QMPS_C128 && dtype != QMPS_C64) return fail(QMPS_ERR_ARG, std::string(who) + ": bad dtype");
  if (N && (!U1 || !U2)) return fail(QMPS_ERR_ARG, std::string(who) + ": null ket unitaries");
  if (N && !bcast_ok(NK, N)) return fail(QMPS_ERR_ARG, std::string(who) + ": NK must be 1 or N");
  return 0;
}

}  // namespace

extern "C" {

int qmps_bw_environment(int side, int64_t N, int64_t NK, const void* U1, const void* U2, int64_t NB,
                        const void* U1_, const void* U2_, int bra_undaggered, void* mat, void* eta,
                        void* vec, int32_t* status, int dtype, void* stream) {
  if (int rc = check_common("bw_environment", N, NK, U1, U2, dtype)) return rc;
  if (side != 0 && side != 1) return fail(QMPS_ERR_ARG, "bw_environment: side must be 0 (right) or 1 (left)");
  if (N && (!U1_ || !U2_ || !bcast_ok(NB, N))) return fail(QMPS_ERR_ARG, "bw_environment: bad bra unitaries");
  if (N == 0) return 0;
  BwParams p;
  memset(&p, 0, sizeof(p));
  p.mode = BW_ENV; p.side = side; p.bra_undaggered = bra_undaggered; p.N = N; p.NK = NK; p.NB = NB;
  p.U1 = U1; p.U2 = U2; p.B1 = U1_; p.B2 = U2_; p.mat = mat; p.eta = eta; p.vec = vec; p.status = status;
  if (option_get(OPT_BW_THREAD))
    return dtype == QMPS_C128 ? launch_bw_env_thread<double>(p, (cudaStream_t)stream)
                              : launch_bw_env_thread<float>(p, (cudaStream_t)stream);
  return run_bw(p, dtype, stream);
}

int qmps_bw_env_apply(int64_t N, int64_t NK, const void* U1, const void* U2, int64_t NB, const void* U1_,
                      const void* U2_, int bra_undaggered, int64_t NM, const void* M, void* out, int dtype,
                      void* stream) {
  if (int rc = check_common("bw_env_apply", N, NK, U1, U2, dtype)) return rc;
  if (N && (!U1_ || !U2_ || !M || !out || !bcast_ok(NB, N) || !bcast_ok(NM, N)))
    return fail(QMPS_ERR_ARG, "bw_env_apply: bad arguments");
  if (N == 0) return 0;
  BwParams p;
  memset(&p, 0, sizeof(p));
  p.mode = BW_APPLY; p.bra_undaggered = bra_undaggered; p.N = N; p.NK = NK; p.NB = NB; p.NM = NM;
  p.U1 = U1; p.U2 = U2; p.B1 = U1_; p.B2 = U2_; p.Mr = M; p.vec = out;
  return run_bw(p, dtype, stream);
}

int qmps_bw_expectation(int64_t N, int64_t NK, const void* U1, const void* U2, int op_qubits, int64_t NO,
                        const void* O, void* out, int dtype, void* stream) {
  if (int rc = check_common("bw_expectation", N, NK, U1, U2, dtype)) return rc;
  if (op_qubits != 2 && op_qubits != 4) return fail(QMPS_ERR_UNSUPPORTED, "bw_expectation: operator on 2 or 4 qubits");
  if (N && (!O || !out || !bcast_ok(NO, N))) return fail(QMPS_ERR_ARG, "bw_expectation: bad arguments");
  if (N == 0) return 0;
  BwParams p;
  memset(&p, 0, sizeof(p));
  p.mode = BW_EXPECT; p.mbits = op_qubits; p.N = N; p.NK = NK; p.NW = NO;
  p.U1 = U1; p.U2 = U2; p.W = O; p.real_out = out;
  return run_bw(p, dtype, stream);
}

int qmps_bw_overlap(int64_t N, int64_t NK, const void* U1, const void* U2, int64_t NB, const void* U1_,
                    const void* U2_, int bra_undaggered, int64_t NM, const void* Mr, const void* Ml, int64_t NW,
                    const void* W, void* overlap, int dtype, void* stream) {
  if (int rc = check_common("bw_overlap", N, NK, U1, U2, dtype)) return rc;
  if (N && (!U1_ || !U2_ || !Mr || !Ml || !W || !overlap || !bcast_ok(NB, N) || !bcast_ok(NM, N) || !bcast_ok(NW, N)))
    return fail(QMPS_ERR_ARG, "bw_overlap: bad arguments");
  if (N == 0) return 0;
  BwParams p;
  memset(&p, 0, sizeof(p));
  p.mode = BW_OVERLAP; p.bra_undaggered = bra_undaggered; p.N = N; p.NK = NK; p.NB = NB; p.NM = NM; p.NW = NW;
  p.U1 = U1; p.U2 = U2; p.B1 = U1_; p.B2 = U2_; p.Mr = Mr; p.Ml = Ml; p.W = W; p.overlap = overlap;
  return run_bw(p, dtype, stream);
}

int qmps_bw_evolve_cost(int64_t N, int64_t NK, const void* U1, const void* U2, int64_t NB, const void* V1,
                        const void* V2, int64_t NW, const void* W, void* cost, void* overlap, void* eta, void* Mr,
                        int32_t* status, int dtype, void* stream) {
  if (int rc = check_common("bw_evolve_cost", N, NK, U1, U2, dtype)) return rc;
  if (N && (!V1 || !V2 || !W || !bcast_ok(NB, N) || !bcast_ok(NW, N)))
    return fail(QMPS_ERR_ARG, "bw_evolve_cost: bad arguments");
  if (N == 0) return 0;
  BwParams p;
  memset(&p, 0, sizeof(p));
  p.mode = BW_COST; p.bra_undaggered = 1; p.N = N; p.NK = NK; p.NB = NB; p.NW = NW;
  p.U1 = U1; p.U2 = U2; p.B1 = V1; p.B2 = V2; p.W = W; p.real_out = cost; p.overlap = overlap; p.eta = eta;
  p.vec = Mr; p.status = status;
  if (NK == 1 && NW == 1 && option_get(OPT_BW_THREAD))
    return dtype == QMPS_C128 ? launch_bw_cost_thread<double>(p, (cudaStream_t)stream)
                              : launch_bw_cost_thread<float>(p, (cudaStream_t)stream);
  return run_bw(p, dtype, stream);
}

}  // extern "C"
