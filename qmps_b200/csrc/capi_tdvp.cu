// qmps_b200 C ABI, classical iTDVP: tangent vectors, Euler / RK4 steps and Loschmidt-rate
// trajectories of single-site uniform MPS, batched over N independent states.  One CTA per problem;
// the n x (n+1) linear systems (n = D^2) live in shared memory up to D = 8 and in a global workspace above.
#include "api_common.cuh"
#include "tdvp.cuh"

using namespace qmps;
using namespace qmps_host;

namespace qmps {
template <typename T> struct TdvpLayout { size_t A, E, x, r, K, rinv, Cc, Ci, Hl, AA, C, G, step, done, red, total; };
// e_in_smem = 0: the n x (n+1) linear system lives in a global workspace
template <typename T> QMPS_HD TdvpLayout<T> tdvp_layout(int d, int D, int G, int e_in_smem) {
  TdvpLayout<T> L;
  Bump b;
  const size_t n = (size_t)D * D, cs = sizeof(cx<T>);
  L.A = b.take(cs * d * n);
  L.E = b.take(e_in_smem ? cs * n * (n + 1) : 0);
  L.x = b.take(cs * n);
  L.r = b.take(cs * n);
  L.K = b.take(cs * n);
  L.rinv = b.take(cs * n);
  L.Cc = b.take(cs * n);
  L.Ci = b.take(cs * n);
  L.Hl = b.take(cs * n);
  L.AA = b.take(cs * d * d * n);
  L.C = b.take(cs * d * d * n);
  L.G = b.take(cs * d * n);
  L.step = b.take(sizeof(int) * n);
  L.done = b.take(sizeof(int) * n);
  L.red = b.take(sizeof(T) * G);
  L.total = (b.off + 127) & ~size_t(127);
  return L;
}

}  // namespace qmps

namespace {

struct TdvpParams {
  int d, D, imaginary;
  int64_t N;
  const void* A; const void* h;
  void* out; void* energy; int32_t* status;
  void* ws; size_t ws_stride;
};

template <typename T, int G>
__global__ void __launch_bounds__(G)
tdvp_tangent_kernel(TdvpParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Grp g; g.lane = threadIdx.x; g.size = G; g.mask = 0xffffffffu; g.cta = 1;
  const int d = p.d, D = p.D, n = D * D;
  const int e_in_smem = (p.ws == nullptr);
  const TdvpLayout<T> L = tdvp_layout<T>(d, D, G, e_in_smem);
  unsigned char* base = smem_raw;
  cx<T>* A = reinterpret_cast<cx<T>*>(base + L.A);
  cx<T>* E = e_in_smem ? reinterpret_cast<cx<T>*>(base + L.E) : reinterpret_cast<cx<T>*>(p.ws) + (size_t)blockIdx.x * p.ws_stride;
  for (int64_t pid = blockIdx.x; pid < p.N; pid += gridDim.x) {
    const cx<T>* src = reinterpret_cast<const cx<T>*>(p.A) + pid * (size_t)(d * n);
    for (int q = g.lane; q < d * n; q += g.size) A[q] = src[q];
    g.sync();
    T en;
    const int st = tdvp_tangent_problem<T>(
        g, A, reinterpret_cast<const cx<T>*>(p.h), p.imaginary, d, D, E, reinterpret_cast<cx<T>*>(base + L.x),
        reinterpret_cast<cx<T>*>(base + L.r), reinterpret_cast<cx<T>*>(base + L.K), reinterpret_cast<cx<T>*>(base + L.rinv),
        reinterpret_cast<cx<T>*>(base + L.Cc), reinterpret_cast<cx<T>*>(base + L.Ci), reinterpret_cast<cx<T>*>(base + L.Hl),
        reinterpret_cast<cx<T>*>(base + L.AA), reinterpret_cast<cx<T>*>(base + L.C), reinterpret_cast<cx<T>*>(base + L.G),
        reinterpret_cast<int*>(base + L.step), reinterpret_cast<int*>(base + L.done), reinterpret_cast<T*>(base + L.red),
        reinterpret_cast<cx<T>*>(p.out) + pid * (size_t)(d * n), &en);
    if (g.lane == 0) {
      if (p.energy) reinterpret_cast<T*>(p.energy)[pid] = en;
      if (p.status) p.status[pid] = st;
    }
    g.sync();
  }
}

// out = sqrt|eta| L^-1 B L, one warp per problem
template <typename T>
__global__ void __launch_bounds__(128)
tdvp_gauge_back_kernel(int d, int D, int64_t N, const cx<T>* __restrict__ B, const cx<T>* __restrict__ Lm,
                       const cx<T>* __restrict__ eta, cx<T>* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, DD = D * D;
  Grp g; g.lane = threadIdx.x & 31; g.size = 32; g.mask = 0xffffffffu; g.cta = 0;
  cx<T>* sB = reinterpret_cast<cx<T>*>(smem_raw) + (size_t)warp * (d * DD + 3 * DD);
  for (int64_t n = (int64_t)blockIdx.x * 4 + warp; n < N; n += (int64_t)gridDim.x * 4) {
    const T scale = sqrt(cabs(eta[n]));
    gauge_back_problem<T>(g, B + n * (size_t)(d * DD), Lm + n * (size_t)DD, scale, d, D, sB, sB + d * DD, sB + d * DD + DD,
                          sB + d * DD + 2 * DD, out + n * (size_t)(d * DD));
  }
}

// y = a + c * x (elementwise over complex arrays), and the RK4 combination
template <typename T>
__global__ void axpy_kernel(int64_t n, const cx<T>* __restrict__ a, const cx<T>* __restrict__ x, T c, cx<T>* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = mk<T>(a[i].re + c * x[i].re, a[i].im + c * x[i].im);
}
template <typename T>
__global__ void rk4_combine_kernel(int64_t n, const cx<T>* __restrict__ a, const cx<T>* __restrict__ k1, const cx<T>* __restrict__ k2,
                                   const cx<T>* __restrict__ k3, const cx<T>* __restrict__ k4, T dt, cx<T>* __restrict__ y) {
  const T c = dt / T(6);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = mk<T>(a[i].re + c * (k1[i].re + T(2) * k2[i].re + T(2) * k3[i].re + k4[i].re),
                 a[i].im + c * (k1[i].im + T(2) * k2[i].im + T(2) * k3[i].im + k4[i].im));
}

template <typename T, int G> int launch_tangent(TdvpParams p, cudaStream_t st) {
  const int n = p.D * p.D;
  const int e_in_smem = n <= 64;
  const TdvpLayout<T> L = tdvp_layout<T>(p.d, p.D, G, e_in_smem);
  auto kern = tdvp_tangent_kernel<T, G>;
  if (L.total > 220 * 1024) return fail(QMPS_ERR_UNSUPPORTED, "tdvp: d^2 D^2 too large for one CTA's shared memory");
  if (int rc = allow_smem(kern, L.total)) return rc;
  int grid = 1;
  if (int rc = persistent_grid(kern, G, L.total, p.N, &grid)) return rc;
  Scratch scratch(st);
  cx<T>* ws = nullptr;
  if (!e_in_smem) {
    p.ws_stride = (size_t)n * (n + 1);
    CK(scratch.get(&ws, sizeof(cx<T>) * p.ws_stride * grid));
    p.ws = ws;
  } else { p.ws = nullptr; p.ws_stride = 0; }
  kern<<<grid, G, L.total, st>>>(p);
  CK(cudaGetLastError());
  return 0;
}

template <typename T> int tangent_any(const TdvpParams& p, cudaStream_t st) {
  return p.D * p.D <= 64 ? launch_tangent<T, 128>(p, st) : launch_tangent<T, 256>(p, st);
}

int check(const char* who, int d, int D, int64_t N, int dtype) {
  if (N < 0) return fail(QMPS_ERR_ARG, std::string(who) + ": negative batch");
  if (D < 1 || D > 16) return fail(QMPS_ERR_UNSUPPORTED, std::string(who) + ": D must be 1..16");
  if (d < 1 || d > 4) return fail(QMPS_ERR_UNSUPPORTED, std::string(who) + ": d must be 1..4");
  if (dtype != QMPS_C128 && dtype != QMPS_C64) return fail(QMPS_ERR_ARG, std::string(who) + ": bad dtype");
  return 0;
}

int ew_grid(int64_t n) { int64_t b = (n + 255) / 256; const int64_t cap = (int64_t)sm_count() * 8; return (int)(b < cap ? (b < 1 ? 1 : b) : cap); }

// dA/dt of tensors in any gauge: canonicalise, tangent, gauge back.  Scratch: AL, dAL [N d D^2], L [N D^2], eta [N].
template <typename T>
int dadt_any(int d, int D, int64_t N, const void* A, const void* h, int imaginary, void* dA, void* energy, int32_t* status,
             int dtype, void* AL, void* dAL, void* Lm, void* eta, cudaStream_t st) {
  if (int rc = qmps_left_canonicalise(d, D, N, A, AL, eta, Lm, status, dtype, st)) return rc;
  TdvpParams p;
  memset(&p, 0, sizeof(p));
  p.d = d; p.D = D; p.imaginary = imaginary; p.N = N; p.A = AL; p.h = h; p.out = dAL; p.energy = energy; p.status = nullptr;
  if (int rc = tangent_any<T>(p, st)) return rc;
  const size_t smem = sizeof(cx<T>) * (size_t)(d * D * D + 3 * D * D) * 4;
  if (int rc = allow_smem(tdvp_gauge_back_kernel<T>, smem)) return rc;
  tdvp_gauge_back_kernel<T><<<ew_grid(N * 64), 128, smem, st>>>(d, D, N, (const cx<T>*)dAL, (const cx<T>*)Lm, (const cx<T>*)eta, (cx<T>*)dA);
  CK(cudaGetLastError());
  return 0;
}

template <typename T>
int evolve_impl(int d, int D, int64_t N, void* A_io, const void* h, double dt, int n_steps, int method, int imaginary, void* traj,
                void* rates, void* energy, int32_t* status, int dtype, cudaStream_t st) {
  const size_t tsz = (size_t)d * D * D, cs = sizeof(cx<T>), rs = sizeof(T);
  const int64_t ne = (int64_t)N * tsz;
  Scratch scratch(st);
  cx<T> *A0 = nullptr, *AL = nullptr, *dAL = nullptr, *Lm = nullptr, *eta = nullptr, *stage = nullptr, *k[4] = {nullptr, nullptr, nullptr, nullptr};
  CK(scratch.get(&A0, cs * ne)); CK(scratch.get(&AL, cs * ne)); CK(scratch.get(&dAL, cs * ne));
  CK(scratch.get(&Lm, cs * N * D * D)); CK(scratch.get(&eta, cs * N)); CK(scratch.get(&stage, cs * ne));
  for (int q = 0; q < (method == 1 ? 4 : 1); ++q) CK(scratch.get(&k[q], cs * ne));
  cx<T>* A = (cx<T>*)A_io;
  // the trajectory starts from the canonical form of the input (as the reference's loops do)
  if (int rc = qmps_left_canonicalise(d, D, N, A, stage, eta, nullptr, status, dtype, st)) return rc;
  CK(cudaMemcpyAsync(A, stage, cs * ne, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(A0, A, cs * ne, cudaMemcpyDeviceToDevice, st));
  auto record = [&](int step) -> int {
    if (traj) CK(cudaMemcpyAsync((char*)traj + (size_t)step * cs * ne, A, cs * ne, cudaMemcpyDeviceToDevice, st));
    if (rates)
      if (int rc = qmps_fixed_point(d, D, N, A, N, A0, 0, 0, nullptr, nullptr, nullptr, (char*)rates + (size_t)step * rs * N, nullptr,
                                    nullptr, dtype, st)) return rc;
    return 0;
  };
  if (int rc = record(0)) return rc;
  const int eg = ew_grid(ne);
  for (int step = 1; step <= n_steps; ++step) {
    void* e_out = energy ? (char*)energy + (size_t)(step - 1) * rs * N : nullptr;
    if (method == 1) {
      // scripts/classical_time_evolution.py:22-26
      if (int rc = dadt_any<T>(d, D, N, A, h, imaginary, k[0], e_out, nullptr, dtype, AL, dAL, Lm, eta, st)) return rc;
      axpy_kernel<T><<<eg, 256, 0, st>>>(ne, A, k[0], (T)(dt / 2), stage);
      if (int rc = dadt_any<T>(d, D, N, stage, h, imaginary, k[1], nullptr, nullptr, dtype, AL, dAL, Lm, eta, st)) return rc;
      axpy_kernel<T><<<eg, 256, 0, st>>>(ne, A, k[1], (T)(dt / 2), stage);
      if (int rc = dadt_any<T>(d, D, N, stage, h, imaginary, k[2], nullptr, nullptr, dtype, AL, dAL, Lm, eta, st)) return rc;
      axpy_kernel<T><<<eg, 256, 0, st>>>(ne, A, k[2], (T)dt, stage);
      if (int rc = dadt_any<T>(d, D, N, stage, h, imaginary, k[3], nullptr, nullptr, dtype, AL, dAL, Lm, eta, st)) return rc;
      rk4_combine_kernel<T><<<eg, 256, 0, st>>>(ne, A, k[0], k[1], k[2], k[3], (T)dt, stage);
    } else {
      // Trajectory.eulerint (qmps/loschmidts/mps_loschmidts.py:22)
      if (int rc = dadt_any<T>(d, D, N, A, h, imaginary, k[0], e_out, nullptr, dtype, AL, dAL, Lm, eta, st)) return rc;
      axpy_kernel<T><<<eg, 256, 0, st>>>(ne, A, k[0], (T)dt, stage);
    }
    if (int rc = qmps_left_canonicalise(d, D, N, stage, A, eta, nullptr, status, dtype, st)) return rc;
    if (int rc = record(step)) return rc;
  }
  CK(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" {

int qmps_tdvp_tangent(int d, int D, int64_t N, const void* AL, const void* h, int imaginary, void* dA, void* energy,
                      int32_t* status, int dtype, void* stream) {
  if (int rc = check("tdvp_tangent", d, D, N, dtype)) return rc;
  if (N && (!AL || !h || !dA)) return fail(QMPS_ERR_ARG, "tdvp_tangent: null array");
  if (N == 0) return 0;
  TdvpParams p;
  memset(&p, 0, sizeof(p));
  p.d = d; p.D = D; p.imaginary = imaginary; p.N = N; p.A = AL; p.h = h; p.out = dA; p.energy = energy; p.status = status;
  return dtype == QMPS_C128 ? tangent_any<double>(p, (cudaStream_t)stream) : tangent_any<float>(p, (cudaStream_t)stream);
}

int qmps_tdvp_dadt(int d, int D, int64_t N, const void* A, const void* h, int imaginary, void* dA, void* energy,
                   int32_t* status, int dtype, void* stream) {
  if (int rc = check("tdvp_dadt", d, D, N, dtype)) return rc;
  if (N && (!A || !h || !dA)) return fail(QMPS_ERR_ARG, "tdvp_dadt: null array");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t cs = dtype == QMPS_C128 ? 16 : 8, ne = (size_t)N * d * D * D;
  Scratch scratch(st);
  char *AL = nullptr, *dAL = nullptr, *Lm = nullptr, *eta = nullptr;
  CK(scratch.get(&AL, cs * ne)); CK(scratch.get(&dAL, cs * ne)); CK(scratch.get(&Lm, cs * N * D * D)); CK(scratch.get(&eta, cs * N));
  return dtype == QMPS_C128 ? dadt_any<double>(d, D, N, A, h, imaginary, dA, energy, status, dtype, AL, dAL, Lm, eta, st)
                            : dadt_any<float>(d, D, N, A, h, imaginary, dA, energy, status, dtype, AL, dAL, Lm, eta, st);
}

int qmps_tdvp_evolve(int d, int D, int64_t N, void* A_io, const void* h, double dt, int n_steps, int method, int imaginary,
                     void* traj, void* rates, void* energy, int32_t* status, int dtype, void* stream) {
  if (int rc = check("tdvp_evolve", d, D, N, dtype)) return rc;
  if (N && (!A_io || !h)) return fail(QMPS_ERR_ARG, "tdvp_evolve: null array");
  if (n_steps < 0 || (method != 0 && method != 1)) return fail(QMPS_ERR_ARG, "tdvp_evolve: n_steps >= 0, method 0 (Euler) or 1 (RK4)");
  if (N == 0) return 0;
  return dtype == QMPS_C128 ? evolve_impl<double>(d, D, N, A_io, h, dt, n_steps, method, imaginary, traj, rates, energy, status, dtype, (cudaStream_t)stream)
                            : evolve_impl<float>(d, D, N, A_io, h, dt, n_steps, method, imaginary, traj, rates, energy, status, dtype, (cudaStream_t)stream);
}

}  // extern "C"
