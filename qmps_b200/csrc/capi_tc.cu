// qmps_b200: host side of the tcgen05 (3xTF32) complex64 path -- the batched complex product
// and the large-D power method built on it (kernels_tc.cuh).  Own translation unit.
#include "api_common.cuh"
#include "kernels_tc.cuh"

using namespace qmps;
namespace qmps_host {
namespace {

typedef cx<float> C;

// Image mode (option tc_presplit: -1 = by size, 0 / 1 force): fp32 images split into hi / lo in shared memory halve
// the global traffic, which is what bounds small K (D = 64: +33 %); at large K the kernel is bound by shared-memory
// operand bandwidth and the extra split traffic costs more than it saves (D = 256: -14 %), so there the images carry
// both planes as in round 1 (profiles/configs5_r02*.jsonl).
int presplit_for(int K) {
  const int v = option_get(OPT_TC_PRESPLIT);
  return v >= 0 ? (v != 0) : (K >= 192);
}
int64_t image_bytes(int64_t nmat, int R, int K, int presplit) {
  return nmat * (int64_t)(R / tc::ROWS) * (K / tc::KS) * (presplit ? tc::SLAB_BYTES : tc::IMG_SLAB_BYTES);
}

int launch_pack(int64_t nmat, int R, int K, const C* in, int64_t mstride, int64_t rstride, int64_t kstride,
                unsigned char* img, int presplit, cudaStream_t st) {
  const int64_t total = nmat * (int64_t)R * K;
  if (total == 0) return 0;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  tc::pack_kernel<<<(unsigned)blocks, 256, 0, st>>>(nmat, R, K, in, mstride, rstride, kstride, img, presplit);
  CK(cudaGetLastError());
  return 0;
}

int launch_tile(const tc::Params& p, cudaStream_t st) {
  const int64_t total = (int64_t)p.batch * p.nrbX * p.nrbY;
  if (total == 0) return 0;
  if (int rc = allow_smem(tc::cgemm_tc_kernel, tc::SMEM_BYTES)) return rc;
  int64_t grid = total;
  if (option_get(OPT_TC_PERSISTENT)) {
    const int64_t cap = sm_count();
    if (grid > cap) grid = cap;
  }
  tc::cgemm_tc_kernel<<<(unsigned)grid, tc::THREADS, tc::SMEM_BYTES, st>>>(p);
  CK(cudaGetLastError());
  return 0;
}

}  // namespace

bool tc_shape_ok(int M, int N, int K) {
  return M > 0 && N > 0 && K > 0 && M % tc::ROWS == 0 && N % tc::ROWS == 0 && K % tc::KS == 0;
}

// C[b] = sum_t X[b][t] . op(Y[b][t]);  X [batch][nsum][M][K], Y [batch][nsum][N][K], C [batch][M][N]
int cgemm_c64_tc(int64_t batch, int nsum, int M, int N, int K, const void* X, const void* Y, int conj_y, void* Cout,
                 cudaStream_t st) {
  if (batch == 0) return 0;
  if (!tc_shape_ok(M, N, K) || nsum < 1) return fail(QMPS_ERR_UNSUPPORTED, "cgemm_c64_tc: M, N must be multiples of 64 and K of 32");
  if (batch * nsum > (int64_t)1 << 30) return fail(QMPS_ERR_UNSUPPORTED, "cgemm_c64_tc: batch too large");
  unsigned char *xi = nullptr, *yi = nullptr;
  const int ps = presplit_for(K);
  CK(malloc_async((void**)&xi, image_bytes(batch * nsum, M, K, ps), st));
  CK(malloc_async((void**)&yi, image_bytes(batch * nsum, N, K, ps), st));
  if (int rc = launch_pack(batch * nsum, M, K, (const C*)X, (int64_t)M * K, K, 1, xi, ps, st)) return rc;
  if (int rc = launch_pack(batch * nsum, N, K, (const C*)Y, (int64_t)N * K, K, 1, yi, ps, st)) return rc;
  tc::Params p;
  memset(&p, 0, sizeof(p));
  p.presplit = ps;
  p.X = xi; p.Y = yi; p.nsum = nsum; p.nkb = K / tc::KS; p.nrbX = M / tc::ROWS; p.nrbY = N / tc::ROWS;
  p.y_div = 1; p.batch = (int)batch; p.conj_y = conj_y; p.a_div = 1; p.out_c = (C*)Cout;
  if (int rc = launch_tile(p, st)) return rc;
  CK(cudaFreeAsync(xi, st));
  CK(cudaFreeAsync(yi, st));
  return 0;
}

bool tm_power_tc_applies(int d, int D, int64_t N) {
  return option_get(OPT_TC_POWER) && D >= 64 && D % 64 == 0 && d >= 1 && N * d < ((int64_t)1 << 30);
}

// K normalised applications r <- sum_s A_s r B_s^dagger / |.|_F in complex64 on tcgen05.
int tm_power_tc(int d, int D, int64_t N, const void* A, const void* B, void* r_io, int K, void* rayleigh,
                cudaStream_t st) {
  if (N == 0) return 0;
  const int nrb = D / tc::ROWS, nkb = D / tc::KS, tiles = nrb * nrb;
  const int64_t DD = (int64_t)D * D;
  unsigned char *Ai = nullptr, *Bi = nullptr, *Ti = nullptr, *Ri = nullptr;
  float* nrm = nullptr; C* dots = nullptr;
  const int ps = presplit_for(D);
  CK(malloc_async((void**)&Ai, image_bytes(N * d, D, D, ps), st));
  CK(malloc_async((void**)&Bi, image_bytes(N * d, D, D, ps), st));
  CK(malloc_async((void**)&Ti, image_bytes(N * d, D, D, ps), st));
  CK(malloc_async((void**)&Ri, image_bytes(N, D, D, ps), st));
  CK(malloc_async((void**)&nrm, sizeof(float) * N * tiles, st));
  C* r = (C*)r_io;
  // A_s[i][k]: rows i, K = k;  B_s[l][j]: rows l, K = j;  r^T: rows j, K = k  (element (row j, k) = r[k][j])
  if (int rc = launch_pack(N * d, D, D, (const C*)A, DD, D, 1, Ai, ps, st)) return rc;
  if (int rc = launch_pack(N * d, D, D, (const C*)B, DD, D, 1, Bi, ps, st)) return rc;
  if (int rc = launch_pack(N, D, D, r, DD, 1, D, Ri, ps, st)) return rc;
  auto apply = [&](const float* norm_in, float* norm_out, C* out_c, const C* dot_with, C* dot_out) -> int {
    tc::Params p1;
    memset(&p1, 0, sizeof(p1));
    p1.presplit = ps;
    p1.X = Ai; p1.Y = Ri; p1.nsum = 1; p1.nkb = nkb; p1.nrbX = nrb; p1.nrbY = nrb; p1.y_div = d;
    p1.batch = (int)(N * d); p1.conj_y = 0; p1.norm_in = norm_in; p1.n_in = tiles; p1.a_div = d;
    p1.out_img = Ti; p1.out_mode = 1; p1.out_nrb = nrb; p1.out_nkb = nkb;
    if (int rc = launch_tile(p1, st)) return rc;
    tc::Params p2;
    memset(&p2, 0, sizeof(p2));
    p2.presplit = ps;
    p2.X = Ti; p2.Y = Bi; p2.nsum = d; p2.nkb = nkb; p2.nrbX = nrb; p2.nrbY = nrb; p2.y_div = 1;
    p2.batch = (int)N; p2.conj_y = 1; p2.a_div = 1; p2.norm_out = norm_out;
    p2.out_img = Ri; p2.out_mode = 2; p2.out_nrb = nrb; p2.out_nkb = nkb;
    p2.out_c = out_c; p2.dot_with = dot_with; p2.dot_out = dot_out;
    return launch_tile(p2, st);
  };
  for (int it = 0; it < K; ++it)
    if (int rc = apply(it == 0 ? nullptr : nrm, nrm, it == K - 1 ? r : nullptr, nullptr, nullptr)) return rc;
  if (K > 0) tc::scale_by_norm_kernel<<<(unsigned)N, 256, 0, st>>>(DD, r, nrm, tiles);
  if (rayleigh) {
    CK(malloc_async((void**)&dots, sizeof(C) * N * tiles, st));
    // the image of r still holds the un-normalised r'; alpha from the same partial norms normalises it
    if (int rc = apply(K > 0 ? nrm : nullptr, nullptr, nullptr, r, dots)) return rc;
    tc::sum_partials_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(N, dots, tiles, (C*)rayleigh);
    CK(cudaFreeAsync(dots, st));
  }
  CK(cudaGetLastError());
  CK(cudaFreeAsync(Ai, st));
  CK(cudaFreeAsync(Bi, st));
  CK(cudaFreeAsync(Ti, st));
  CK(cudaFreeAsync(Ri, st));
  CK(cudaFreeAsync(nrm, st));
  return 0;
}

}  // namespace qmps_host
