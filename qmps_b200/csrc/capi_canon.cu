// qmps_b200 C ABI, canonical forms and local expectation values.
// Own translation unit; composes the exported solvers (qmps_fixed_point, qmps_env_exact) with
// the O(d D^3) gauge / expectation kernels of kernels_canon.cuh on one stream, no host sync.
#include "api_common.cuh"
#include "kernels_canon.cuh"

using namespace qmps;
using namespace qmps_host;

namespace {

__global__ void status_merge_kernel(int64_t N, int32_t* __restrict__ dst, const int32_t* __restrict__ src) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    if (dst[i] == ST_OK) dst[i] = src[i];
}

int canon_grid(int64_t N) {
  const int64_t blocks = (N + CANON_WARPS - 1) / CANON_WARPS;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

template <typename T> int launch_gauge(const GaugeParams& p, cudaStream_t st) {
  const size_t smem = gauge_smem_per_warp<T>(p.d, p.D) * CANON_WARPS;
  if (int rc = allow_smem(gauge_kernel<T>, smem)) return rc;
  gauge_kernel<T><<<canon_grid(p.N), CANON_WARPS * 32, smem, st>>>(p);
  CK(cudaGetLastError());
  return 0;
}

template <typename T> int launch_expect(const ExpectParams& p, cudaStream_t st) {
  const size_t smem = expect_smem_per_warp<T>(p.d, p.D) * CANON_WARPS;
  if (int rc = allow_smem(expect_kernel<T>, smem)) return rc;
  expect_kernel<T><<<canon_grid(p.N), CANON_WARPS * 32, smem, st>>>(p);
  CK(cudaGetLastError());
  return 0;
}

int check_shape(const char* who, int d, int D, int64_t N, int dtype) {
  if (N < 0) return fail(QMPS_ERR_ARG, std::string(who) + ": negative batch");
  if (D < 1 || D > 16) return fail(QMPS_ERR_UNSUPPORTED, std::string(who) + ": D must be 1..16");
  if (d < 1 || d > 4) return fail(QMPS_ERR_UNSUPPORTED, std::string(who) + ": d must be 1..4");
  if (dtype != QMPS_C128 && dtype != QMPS_C64) return fail(QMPS_ERR_ARG, std::string(who) + ": bad dtype");
  return 0;
}

}  // namespace

extern "C" {

int qmps_gauge_transform(int d, int D, int64_t N, const void* A, const void* X, int x_kind, const void* eta,
                         void* A_out, void* G_out, int32_t* status, int dtype, void* stream) {
  if (int rc = check_shape("gauge_transform", d, D, N, dtype)) return rc;
  if (N && (!A || !X || !A_out)) return fail(QMPS_ERR_ARG, "gauge_transform: null array");
  if (x_kind != 0 && x_kind != 1) return fail(QMPS_ERR_ARG, "gauge_transform: x_kind must be 0 or 1");
  if (N == 0) return 0;
  GaugeParams p;
  memset(&p, 0, sizeof(p));
  p.d = d; p.D = D; p.N = N; p.A = A; p.X = X; p.eta = eta; p.x_kind = x_kind; p.A_out = A_out; p.G_out = G_out;
  p.status = status; p.keep_status = 0;
  return dtype == QMPS_C128 ? launch_gauge<double>(p, (cudaStream_t)stream) : launch_gauge<float>(p, (cudaStream_t)stream);
}

int qmps_left_canonicalise(int d, int D, int64_t N, const void* A, void* AL, void* eta, void* L, int32_t* status,
                           int dtype, void* stream) {
  if (int rc = check_shape("left_canonicalise", d, D, N, dtype)) return rc;
  if (N && (!A || !AL)) return fail(QMPS_ERR_ARG, "left_canonicalise: null array");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t csz = dtype == QMPS_C128 ? 16 : 8;
  void* lvec = nullptr; void* eta_tmp = nullptr;
  CK(malloc_async(&lvec, csz * (size_t)N * D * D, st));
  if (!eta) { CK(malloc_async(&eta_tmp, csz * (size_t)N, st)); }
  void* e = eta ? eta : eta_tmp;
  // l must come out Hermitian: the trace gauge (a positive multiple of a Hermitian PD matrix has tr > 0)
  int rc = qmps_fixed_point_ex(d, D, N, A, N, A, 0, 1, QMPS_GAUGE_TRACE, e, lvec, nullptr, nullptr, nullptr, status,
                               dtype, stream);
  if (!rc) {
    GaugeParams p;
    memset(&p, 0, sizeof(p));
    p.d = d; p.D = D; p.N = N; p.A = A; p.X = lvec; p.eta = e; p.x_kind = 0; p.A_out = AL; p.G_out = L;
    p.status = status; p.keep_status = 1;
    rc = dtype == QMPS_C128 ? launch_gauge<double>(p, st) : launch_gauge<float>(p, st);
  }
  cudaFreeAsync(lvec, st);
  if (eta_tmp) cudaFreeAsync(eta_tmp, st);
  return rc;
}

int qmps_mixed_canonical(int d, int D, int64_t N, const void* A, int assume_left_canonical, void* AL, void* AR,
                         void* C, void* eta, int32_t* status, int dtype, void* stream) {
  if (int rc = check_shape("mixed_canonical", d, D, N, dtype)) return rc;
  if (N && !A) return fail(QMPS_ERR_ARG, "mixed_canonical: null input");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t csz = dtype == QMPS_C128 ? 16 : 8;
  const size_t a_bytes = csz * (size_t)N * d * D * D, m_bytes = csz * (size_t)N * D * D;
  void* al_tmp = nullptr; void* c_tmp = nullptr; int32_t* st2 = nullptr;
  int rc = 0;
  const void* al = A;
  if (!assume_left_canonical) {
    if (!AL) { CK(malloc_async(&al_tmp, a_bytes, st)); }
    void* dst = AL ? AL : al_tmp;
    rc = qmps_left_canonicalise(d, D, N, A, dst, eta, nullptr, status, dtype, stream);
    al = dst;
  } else if (AL && AL != A) {
    CK(cudaMemcpyAsync(AL, A, a_bytes, cudaMemcpyDeviceToDevice, st));
  }
  if (!rc && (AR || C)) {
    if (!C) { CK(malloc_async(&c_tmp, m_bytes, st)); }
    void* cc = C ? C : c_tmp;
    int32_t* s2 = status;
    if (status && !assume_left_canonical) { CK(malloc_async((void**)&st2, sizeof(int32_t) * (size_t)N, st)); s2 = st2; }
    rc = qmps_env_exact(d, D, N, al, 0, 1, assume_left_canonical ? eta : nullptr, nullptr, cc, s2, dtype, stream);
    if (!rc && st2) {
      status_merge_kernel<<<canon_grid(N), 128, 0, st>>>(N, status, st2);
      if (cudaGetLastError() != cudaSuccess) rc = fail(QMPS_ERR_CUDA, "mixed_canonical: status merge launch failed");
    }
    if (!rc && AR) {
      GaugeParams p;
      memset(&p, 0, sizeof(p));
      p.d = d; p.D = D; p.N = N; p.A = al; p.X = cc; p.x_kind = 1; p.A_out = AR;
      rc = dtype == QMPS_C128 ? launch_gauge<double>(p, st) : launch_gauge<float>(p, st);
    }
  } else if (!rc && assume_left_canonical && (eta || status)) {
    rc = qmps_env_exact(d, D, N, al, 0, 1, eta, nullptr, nullptr, status, dtype, stream);
  }
  if (al_tmp) cudaFreeAsync(al_tmp, st);
  if (c_tmp) cudaFreeAsync(c_tmp, st);
  if (st2) cudaFreeAsync(st2, st);
  return rc;
}

int qmps_expectation(int d, int D, int64_t N, const void* A, const void* r, const void* lvec, const void* eta,
                     int nops, const void* ops, void* out, int dtype, void* stream) {
  if (int rc = check_shape("expectation", d, D, N, dtype)) return rc;
  if (nops < 0 || (N && nops && (!A || !r || !ops || !out))) return fail(QMPS_ERR_ARG, "expectation: null array");
  if (N == 0 || nops == 0) return 0;
  ExpectParams p;
  memset(&p, 0, sizeof(p));
  p.d = d; p.D = D; p.nops = nops; p.N = N; p.A = A; p.r = r; p.lvec = lvec; p.eta = eta; p.ops = ops; p.out = out;
  return dtype == QMPS_C128 ? launch_expect<double>(p, (cudaStream_t)stream) : launch_expect<float>(p, (cudaStream_t)stream);
}

}  // extern "C"
