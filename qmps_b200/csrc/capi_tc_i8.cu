// qmps_b200: host side of the tcgen05 kind::i8 complex128 path (kernels_tc_i8.cuh) -- the batched complex
// product from exact int8 slice products and the large-D power method built on it.  Own translation unit.
#include "api_common.cuh"
#include "kernels_tc_i8.cuh"

using namespace qmps;
namespace qmps_host {
namespace {

typedef cx<double> Z;

int64_t ximg_bytes(int64_t nmat, int R, int K) { return nmat * (int64_t)(R / tci8::XROWS) * (K / tci8::KS) * tci8::X_SLAB; }
int64_t yimg_bytes(int64_t nmat, int R, int K) { return nmat * (int64_t)(R / tci8::YROWS) * (K / tci8::KS) * tci8::Y_SLAB; }

int launch_slice(int64_t nmat, int R, int K, const Z* in, int64_t mstride, int64_t rstride, int kin, int64_t kstride,
                 int64_t kostride, int is_y, const double* norm_in, int n_in, int a_div, unsigned char* img, int* ex,
                 cudaStream_t st) {
  if (nmat * R == 0) return 0;
  int G = 1;
  while (G < 32 && G < K / 16) G *= 2;                      // lanes per row: the power of two >= K / 16 chunks (<= 32)
  const int64_t warps = (nmat * R + (32 / G) - 1) / (32 / G);
  int64_t blocks = (warps + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (K / 16 <= G)
    tci8::slice_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(nmat, R, K, in, mstride, rstride, kin, kstride, kostride, is_y, norm_in,
                                                               n_in, a_div, img, ex, G);
  else
    tci8::slice_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(nmat, R, K, in, mstride, rstride, kin, kstride, kostride, is_y, norm_in,
                                                                n_in, a_div, img, ex, G);
  CK(cudaGetLastError());
  return 0;
}

int launch_tile(const tci8::Params& p, cudaStream_t st) {
  const int64_t total = (int64_t)p.batch * p.nrbX * p.ncbY;
  if (total == 0) return 0;
  if (int rc = allow_smem(tci8::zgemm_i8_kernel, tci8::SMEM_BYTES)) return rc;
  int64_t grid = total;
  const int64_t cap = sm_count();
  if (grid > cap) grid = cap;
  tci8::zgemm_i8_kernel<<<(unsigned)grid, tci8::THREADS, tci8::SMEM_BYTES, st>>>(p);
  CK(cudaGetLastError());
  return 0;
}

}  // namespace

bool i8_shape_ok(int M, int N, int K) {
  return M > 0 && N > 0 && K > 0 && M % tci8::XROWS == 0 && N % tci8::YROWS == 0 && K % tci8::KS == 0 && K < (1 << 19);
}

// C[b] = X[b] . op(Y[b / y_div]);  X [batch][M][K], Y [batch / y_div][N][K], C [batch][M][N], all complex128
int zgemm_c128_i8(int64_t batch, int M, int N, int K, const void* X, const void* Y, int conj_y, void* Cout, cudaStream_t st) {
  if (batch == 0) return 0;
  if (!i8_shape_ok(M, N, K)) return fail(QMPS_ERR_UNSUPPORTED, "zgemm_c128_i8: M must be a multiple of 64, N of 32, K of 64");
  if (batch > (int64_t)1 << 30) return fail(QMPS_ERR_UNSUPPORTED, "zgemm_c128_i8: batch too large");
  Scratch scratch(st);
  unsigned char *xi = nullptr, *yi = nullptr; int *ex = nullptr, *ey = nullptr;
  CK(scratch.get(&xi, ximg_bytes(batch, M, K)));
  CK(scratch.get(&yi, yimg_bytes(batch, N, K)));
  CK(scratch.get(&ex, sizeof(int) * batch * M * 2));
  CK(scratch.get(&ey, sizeof(int) * batch * N * 2));
  if (int rc = launch_slice(batch, M, K, (const Z*)X, (int64_t)M * K, K, K, 1, 0, 0, nullptr, 0, 1, xi, ex, st)) return rc;
  if (int rc = launch_slice(batch, N, K, (const Z*)Y, (int64_t)N * K, K, K, 1, 0, 1, nullptr, 0, 1, yi, ey, st)) return rc;
  tci8::Params p;
  memset(&p, 0, sizeof(p));
  p.X = xi; p.Y = yi; p.ex_x = ex; p.ex_y = ey; p.nkb = K / tci8::KS; p.nrbX = M / tci8::XROWS; p.ncbY = N / tci8::YROWS;
  p.y_div = 1; p.batch = (int)batch; p.conj_y = conj_y; p.out_c = (Z*)Cout;
  return launch_tile(p, st);
}

bool tm_power_i8_applies(int d, int D, int64_t N) {
  // option i8_power: 0 = never (FP64 tensor pipe, DMMA), 1 / 2 = whenever the shapes allow.  Measured
  // (profiles/exp_i8_r02h.jsonl, algorithmic TFLOP/s, i8 vs DMMA): D = 64 21.6 vs 21.4, 128 38.2 vs 25.3, 256 61.4 vs 27.7.
  const int mode = option_get(OPT_I8_POWER);
  if (!mode) return false;
  return D >= 64 && D % 64 == 0 && d >= 1 && (int64_t)d * D < (1 << 19) && N * d < ((int64_t)1 << 30);
}

static __global__ void __launch_bounds__(256)
i8_scale_kernel(int64_t len, Z* __restrict__ r, const double* __restrict__ norm, int n) {
  double s = 0.0;
  for (int j = 0; j < n; ++j) s += norm[(size_t)blockIdx.x * n + j];
  const double a = rsqrt(s);
  Z* p = r + (size_t)blockIdx.x * len;
  for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.y * blockDim.x) p[i] = p[i] * a;
}
static __global__ void i8_sum_partials_kernel(int64_t nb, const Z* __restrict__ part, int n, Z* __restrict__ out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  Z s = mk<double>(0.0, 0.0);
  for (int j = 0; j < n; ++j) s = s + part[b * n + j];
  out[b] = s;
}

// K normalised applications r <- sum_s A_s r B_s^dagger / |.|_F in complex128 on tcgen05 kind::i8.
//   stage 1   T[b,s] = A[b,s] . (r[b] / |r[b]|)      X = A_s (rows i, K = k),  Y = r^T (rows j, K = k)
//   stage 2   r'[b]  = sum_s T[b,s] . B[b,s]^dagger   X = [T_0 | T_1 ...] (rows i, K = (s, j)),  Y = [B_0 | B_1 ...] (rows l)
// The sum over s is a concatenation along K under one row exponent, so it accumulates exactly in the integers.
int tm_power_i8(int d, int D, int64_t N, const void* A, const void* B, void* r_io, int K, void* rayleigh, cudaStream_t st) {
  if (N == 0) return 0;
  const int64_t DD = (int64_t)D * D;
  const int tiles2 = (D / tci8::XROWS) * (D / tci8::YROWS);
  Scratch scratch(st);
  unsigned char *Ai = nullptr, *Bi = nullptr, *Ti = nullptr, *Ri = nullptr;
  int *eA = nullptr, *eB = nullptr, *eT = nullptr, *eR = nullptr;
  Z *Tb = nullptr, *Er = nullptr, *dots = nullptr, *Rt = nullptr; double* nrm = nullptr;
  CK(scratch.get(&Ai, ximg_bytes(N * d, D, D)));
  CK(scratch.get(&Bi, yimg_bytes(N, D, d * D)));
  CK(scratch.get(&Ti, ximg_bytes(N, D, d * D)));
  CK(scratch.get(&Ri, yimg_bytes(N, D, D)));
  CK(scratch.get(&eA, sizeof(int) * N * d * D * 2));
  CK(scratch.get(&eB, sizeof(int) * N * D * 2));
  CK(scratch.get(&eT, sizeof(int) * N * D * 2));
  CK(scratch.get(&eR, sizeof(int) * N * D * 2));
  CK(scratch.get(&Tb, sizeof(Z) * N * d * DD));
  CK(scratch.get(&nrm, sizeof(double) * N * tiles2));
  CK(scratch.get(&Rt, sizeof(Z) * N * DD));
  Z* r = (Z*)r_io;
  // constant operands, sliced once: A_s[i][k] as X; B[b][l][(s, j)] = B_s[l][j] as Y
  if (int rc = launch_slice(N * d, D, D, (const Z*)A, DD, D, D, 1, 0, 0, nullptr, 0, 1, Ai, eA, st)) return rc;
  if (int rc = launch_slice(N, D, d * D, (const Z*)B, d * DD, D, D, 1, DD, 1, nullptr, 0, 1, Bi, eB, st)) return rc;
  // src_t != nullptr: r is available transposed (rows j, K = k contiguous: coalesced slicing), else strided reads of src
  auto apply = [&](const Z* src, const Z* src_t, const double* norm_in, double* norm_out, Z* out_c, Z* out_ct, const Z* dot_with,
                   Z* dot_out) -> int {
    // r^T: rows j, K = k, element (j, k) = r[k][j]; the 1 / |r| of the previous application is folded into the slicing
    if (src_t) { if (int rc = launch_slice(N, D, D, src_t, DD, D, D, 1, 0, 1, norm_in, tiles2, 1, Ri, eR, st)) return rc; }
    else if (int rc = launch_slice(N, D, D, src, DD, 1, D, D, 0, 1, norm_in, tiles2, 1, Ri, eR, st)) return rc;
    tci8::Params p1;
    memset(&p1, 0, sizeof(p1));
    p1.X = Ai; p1.Y = Ri; p1.ex_x = eA; p1.ex_y = eR; p1.nkb = D / tci8::KS; p1.nrbX = D / tci8::XROWS; p1.ncbY = D / tci8::YROWS;
    p1.y_div = d; p1.batch = (int)(N * d); p1.conj_y = 0; p1.out_c = Tb;
    if (int rc = launch_tile(p1, st)) return rc;
    // [T_0 | T_1 | ...]: rows i, K = (s, j)
    if (int rc = launch_slice(N, D, d * D, Tb, d * DD, D, D, 1, DD, 0, nullptr, 0, 1, Ti, eT, st)) return rc;
    tci8::Params p2;
    memset(&p2, 0, sizeof(p2));
    p2.X = Ti; p2.Y = Bi; p2.ex_x = eT; p2.ex_y = eB; p2.nkb = d * D / tci8::KS; p2.nrbX = D / tci8::XROWS; p2.ncbY = D / tci8::YROWS;
    p2.y_div = 1; p2.batch = (int)N; p2.conj_y = 1; p2.out_c = out_c; p2.out_ct = out_ct; p2.norm_out = norm_out; p2.dot_with = dot_with; p2.dot_out = dot_out;
    return launch_tile(p2, st);
  };
  for (int it = 0; it < K; ++it)
    if (int rc = apply(r, it == 0 ? nullptr : Rt, it == 0 ? nullptr : nrm, nrm, it == K - 1 ? r : nullptr, Rt, nullptr, nullptr)) return rc;
  if (K > 0) i8_scale_kernel<<<dim3((unsigned)N, (unsigned)((DD + 4095) / 4096 < 64 ? (DD + 4095) / 4096 : 64)), 256, 0, st>>>(DD, r, nrm, tiles2);
  if (rayleigh) {
    CK(scratch.get(&Er, sizeof(Z) * N * DD));
    CK(scratch.get(&dots, sizeof(Z) * N * tiles2));
    if (int rc = apply(r, nullptr, nullptr, nullptr, Er, nullptr, r, dots)) return rc;
    i8_sum_partials_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(N, dots, tiles2, (Z*)rayleigh);
  }
  CK(cudaGetLastError());
  return 0;
}

}  // namespace qmps_host
