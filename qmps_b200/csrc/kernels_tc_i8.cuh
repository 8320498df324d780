// qmps_b200 large-D transfer-matrix application in complex128 on the 5th-generation tensor cores
// (BASELINE config 5; qmps.ipynb cells 29-32; SURVEY 8(d) cfg 5).
//
// tcgen05.mma has no FP64 kind.  complex128 results to the 1e-10 parity bar come from EXACT integer
// arithmetic on kind::i8 (Ozaki-style slicing; accuracy prototyped in tools/ozaki_prototype.py):
//   * every real operand row is scaled by a power of two so that |x| < 1/2 and cut into six signed 7-bit slices,
//         x = 2^e sum_{p<6} q_p 2^{-7(p+1)},   |q_p| <= 64  (int8),   remainder < 2^-43 |row|max;
//   * a slice product q_i^X . q_j^Y is exact in int32 (K 64^2 < 2^31 for K < 2^19); only the 21 pairs with
//     i + j < 6 are formed, and all pairs of one level t = i + j accumulate in ONE int32 TMEM accumulator, so a
//     tile needs six accumulators (6 x 64 columns of the 512);
//   * the epilogue converts the six integers to FP64, weights them 2^{-7(t+2)}, applies the row / column
//     exponents and combines the four real products of the complex one -- all in FP64 registers.
// Measured relative error of one application: 4e-12 (6 slices), the figure the numpy prototype gives.
//
// Complex -> real as in kernels_tc.cuh: an X block of 64 complex rows is 128 plane rows [re ; im], a Y block
// of 32 complex rows is 64 plane rows, one UMMA of shape M = 128, N = 64, K = 32 (bytes) yields RR|RI / IR|II.
//
// Memory: operands live in global memory as "slice images" -- per (matrix, row block, K slab of 64) three planes,
// plane p holding slices 2p (bytes 0..63 of each 128-byte row) and 2p + 1 (bytes 64..127), each plane laid out
// exactly as the K-major SWIZZLE_128B UMMA tile.  A slab (48 KB of X + 24 KB of Y) moves with two bulk copies of
// the TMA engine (cp.async.bulk + mbarrier complete_tx); no tensor map.
//
// Kernel: persistent, warp-specialised, 192 threads, 1 CTA per SM: warp 0 = TMA producer (2-stage ring, 72 KB per
// stage), warp 1 = single-thread MMA issuer (42 UMMAs per slab), warps 2-9 = epilogue (tcgen05.ld of the six
// accumulators, FP64 recombination, lane-pair exchange through shared memory, norms / dot products, stores): two warps
// per TMEM lane quarter, each taking half of the tile's columns -- the FP64 recombination is the critical path of a tile
// (profiles/ncu_i8gemm_r02f.txt), so it gets eight warps instead of four.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "core.cuh"
#include "kernels_tc.cuh"

namespace qmps {
namespace tci8 {

constexpr int NSL = 6;                          // slices per real number
constexpr int XROWS = 64;                       // complex rows per X block (M = 128 plane rows)
constexpr int YROWS = 32;                       // complex rows per Y block (N = 64 plane rows)
constexpr int KS = 64;                          // K elements per slab
constexpr int X_PLANE = 128 * 128, Y_PLANE = 64 * 128;
constexpr int X_SLAB = 3 * X_PLANE, Y_SLAB = 3 * Y_PLANE;
constexpr int STAGE_BYTES = X_SLAB + Y_SLAB;    // 73 728
constexpr int NSTAGE = 2;
constexpr int XCH_BYTES = 128 * 32 * 8;         // 32 doubles per epilogue thread
constexpr int THREADS = 320;                     // producer warp, MMA warp, eight epilogue warps
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + XCH_BYTES + 1024;

// byte offset of (plane row, byte column) inside a SWIZZLE_128B plane
QMPS_HD uint32_t img_off(int prow, int bcol) {
  return (uint32_t)((prow >> 3) * 1024 + (prow & 7) * 128 + ((((bcol >> 4) ^ prow) & 7) << 4) + (bcol & 15));
}

#if defined(__CUDACC__)

// ---- slicing: FP64 complex matrices -> slice images + per-plane-row exponents -------------------------------
// element (row, k) of matrix m is  in[m * mstride + row * rstride + (k / kin) * kostride + (k % kin) * kstride]
// (kin < K concatenates several matrices along K under ONE row exponent: the sum over the physical index of
// stage 2 then accumulates exactly in the integer accumulators).  is_y: blocks of 32 rows instead of 64.
// scale (optional): value multiplied by rsqrt(sum_j norm_in[(m / a_div) * n_in + j]) before slicing.
// One warp per (matrix, complex row).  ex[(m * R + row) * 2 + part]: exponent e, x ~ 2^e sum q_p 2^{-7(p+1)}.
template <bool SINGLE>
static __global__ void __launch_bounds__(256, 2)
slice_kernel(int64_t nmat, int R, int K, const cx<double>* __restrict__ in, int64_t mstride, int64_t rstride, int kin,
             int64_t kstride, int64_t kostride, int is_y, const double* __restrict__ norm_in, int n_in, int a_div,
             unsigned char* __restrict__ img, int* __restrict__ ex, int G) {
  // G lanes per complex row (a power of two <= 32, <= K / 16): a K = 64 row keeps 4 lanes busy, so 8 rows share a warp
  const int lane = threadIdx.x & 31, gl = lane & (G - 1), rpw = 32 / G;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int brows = is_y ? YROWS : XROWS;
  const int nrb = R / brows, nkb = K / KS, nchunk = K / 16;
  const int64_t slab = is_y ? Y_SLAB : X_SLAB, plane = is_y ? Y_PLANE : X_PLANE;
  const int64_t nrows = nmat * R;
  for (int64_t w0 = warp * rpw; w0 < nrows; w0 += nwarps * rpw) {
    int64_t w = w0 + lane / G;
    const bool live = w < nrows;
    if (!live) w = nrows - 1;
    const int64_t m = w / R;
    const int row = (int)(w - m * R);
    double alpha = 1.0;
    if (norm_in) {
      const double* ni = norm_in + (m / a_div) * n_in;
      double s = 0.0;
      for (int j = 0; j < n_in; ++j) s += ni[j];
      alpha = rsqrt(s);
    }
    const cx<double>* src = in + m * mstride + row * rstride;
    double mre = 0.0, mim = 0.0;
    // one chunk per lane (K <= 512): the 16 elements are loaded ONCE, all 16 loads in flight, and stay in registers
    // for the slicing below (the kernel is bound by global-load latency: profiles/ncu_slice_r02e.txt)
    constexpr bool single = SINGLE;                  // nchunk <= G, decided by the launcher
    cx<double> zc[16];
    if (single) {
      const int c = gl < nchunk ? gl : nchunk - 1;
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const int k = c * 16 + t;
        zc[t] = src[(k / kin) * kostride + (k % kin) * kstride];
      }
#pragma unroll
      for (int t = 0; t < 16; ++t) { mre = fmax(mre, fabs(zc[t].re)); mim = fmax(mim, fabs(zc[t].im)); }
    } else {
      for (int c = gl; c < nchunk; c += G) {
#pragma unroll 4
        for (int t = 0; t < 16; ++t) {
          const int k = c * 16 + t;
          const cx<double> z = src[(k / kin) * kostride + (k % kin) * kstride];
          mre = fmax(mre, fabs(z.re)); mim = fmax(mim, fabs(z.im));
        }
      }
    }
    for (int o = G >> 1; o > 0; o >>= 1) {
      mre = fmax(mre, __shfl_xor_sync(0xffffffffu, mre, o));
      mim = fmax(mim, __shfl_xor_sync(0xffffffffu, mim, o));
    }
    int ere = 0, eim = 0;
    if (mre * alpha > 0.0) { frexp(mre * alpha, &ere); ere += 1; }      // |x| / 2^e < 1/2
    if (mim * alpha > 0.0) { frexp(mim * alpha, &eim); eim += 1; }
    if (live && gl == 0) { ex[(m * R + row) * 2] = ere; ex[(m * R + row) * 2 + 1] = eim; }
    // Fixed point, not floating point: v = round(x 2^(42 - e)) as a 64-bit integer by the magic-number add (one DFMA;
    // |v| <= 2^41).  Adding the bias B = 64 (1 + 128 + ... + 128^5) makes every balanced base-128 digit d_p = u_p - 64
    // non-negative, so the six slices are plain bit fields of v + B -- no carry chain: a funnel shift and a mask per
    // digit, four digits per 32-bit word, and one SWAR byte-wise subtraction of 64 per word.  (The first version cut the
    // slices with rint / double->int conversions and the second with a serial carry chain; at ~110 integer
    // instructions per real number the slicing passes cost as much as the GEMMs: profiles/launches_i8_r02*.csv.)
    const double MAGIC = 6755399441055744.0;                             // 1.5 * 2^52: ulp = 1
    constexpr long long BIAS = 64ll * ((1ll << 42) - 1) / 127;           // sum_p 64 * 128^p
    const long long KSUB = 0x4338000000000000ll - BIAS;                  // bits(MAGIC) - B
    const double sre = ldexp(alpha, 42 - ere), sim = ldexp(alpha, 42 - eim);
    const int rb = row / brows, i = row - rb * brows;
    for (int c = gl; c < nchunk && live; c += G) {                       // 16 consecutive k -> one 16-byte chunk per slice
      const int k0 = c * 16, kb = k0 / KS, kk = k0 - kb * KS;
      unsigned char* base = img + (((m * nrb + rb) * nkb + kb) * slab);
      unsigned long long ur[16], ui[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const int k = k0 + t;
        const cx<double> z = single ? zc[t] : src[(k / kin) * kostride + (k % kin) * kstride];
        ur[t] = (unsigned long long)(__double_as_longlong(fma(z.re, sre, MAGIC)) - KSUB);     // v + B >= 0
        ui[t] = (unsigned long long)(__double_as_longlong(fma(z.im, sim, MAGIC)) - KSUB);
      }
#pragma unroll
      for (int p = 0; p < NSL; ++p) {
        const int sh = 7 * (NSL - 1 - p);                                // slice p has weight 128^(5 - p) in v
        uint32_t wr[4] = {0, 0, 0, 0}, wi[4] = {0, 0, 0, 0};
        if (p > 0) {
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            wr[t >> 2] |= ((uint32_t)(ur[t] >> sh) & 127u) << (8 * (t & 3));
            wi[t >> 2] |= ((uint32_t)(ui[t] >> sh) & 127u) << (8 * (t & 3));
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {                                  // bytes u in [0, 127] -> (u - 64) as int8
            wr[q] = ((wr[q] | 0x80808080u) - 0x40404040u) ^ 0x80808080u;
            wi[q] = ((wi[q] | 0x80808080u) - 0x40404040u) ^ 0x80808080u;
          }
        } else {                                                         // top digit: u <= 128, no mask
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            wr[t >> 2] |= ((uint32_t)((int)(ur[t] >> sh) - 64) & 0xffu) << (8 * (t & 3));
            wi[t >> 2] |= ((uint32_t)((int)(ui[t] >> sh) - 64) & 0xffu) << (8 * (t & 3));
          }
        }
        const int bcol = (p & 1) * 64 + kk;
        unsigned char* pl = base + (p >> 1) * plane;
        *reinterpret_cast<uint4*>(pl + img_off(i, bcol)) = make_uint4(wr[0], wr[1], wr[2], wr[3]);
        *reinterpret_cast<uint4*>(pl + img_off(brows + i, bcol)) = make_uint4(wi[0], wi[1], wi[2], wi[3]);
      }
    }
  }
}

// 2^e as a double (|e| <= 1022), and an exact int32 -> double conversion without the conversion pipe:
// the double with high word 0x43300000 and low word L is 2^52 + L
__device__ __forceinline__ double exp2i(int e) {
  e = e < -1022 ? -1022 : (e > 1023 ? 1023 : e);
  return __hiloint2double((1023 + e) << 20, 0);
}
__device__ __forceinline__ double i2d(int v) {
  return __hiloint2double(0x43300000, (int)((unsigned)v ^ 0x80000000u)) - 4503601774854144.0;   // 2^52 + 2^31
}

// D[tmem] (+)= A[smem] . B[smem]^T, kind::i8 (int8 x int8 -> int32), issued by one thread
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32_i(uint32_t taddr, int (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = (int)r[i];
}

__device__ __forceinline__ void tmem_ld16_i(uint32_t taddr, int (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = (int)r[i];
}
// the same without the wait: several loads in flight, one tmem_ld_wait() before the values are used
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, int (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// exact int64 -> double for |x| < 2^51 without the conversion pipe
__device__ __forceinline__ double ll2d(long long x) {
  return __longlong_as_double(x + 0x4338000000000000ll) - 6755399441055744.0;
}
__device__ __forceinline__ void epi_bar8() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// instruction descriptor: D = S32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), both K-major,
// N = 64 ((64 >> 3) << 17), M = 128 ((128 >> 4) << 24)
constexpr uint32_t IDESC_I8_128x64 = (2u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

struct Params {
  const unsigned char* X;      // slice images, matrix index bz
  const unsigned char* Y;      // slice images, matrix index bz / y_div
  const int* ex_x;             // [batch][M][2]
  const int* ex_y;             // [batch / y_div][N][2]
  int nkb, nrbX, ncbY, y_div, batch, conj_y;
  double* norm_out;            // [bz][tile] partial sums of |C|^2
  cx<double>* out_c;           // interleaved C[bz][M][N]
  cx<double>* out_ct;          // interleaved transpose C^T[bz][N][M] (what the next application slices as r^T)
  const cx<double>* dot_with;  // dot_out[bz][tile] = sum conj(dot_with[bz][i][l]) C[i][l]
  cx<double>* dot_out;
};

static __global__ void __launch_bounds__(THREADS, 1)
zgemm_i8_kernel(Params p) {
  using namespace tc;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t s_bar[2 * NSTAGE + 2];
  __shared__ uint32_t s_tmem;
  __shared__ double s_red[8][4];
  __shared__ double s_cs[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t dyn0 = smem_u32(smem_dyn);
  const uint32_t ring = (dyn0 + 1023u) & ~1023u;
  double2* xch = reinterpret_cast<double2*>(smem_dyn + (ring - dyn0) + NSTAGE * STAGE_BYTES);
  const uint32_t bar0 = smem_u32(s_bar);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  const uint32_t tfull_bar = bar0 + 8u * (2 * NSTAGE), tempty_bar = bar0 + 8u * (2 * NSTAGE + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1); mbar_init(tempty_bar, 256);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&s_tmem), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  const int tiles_per = p.nrbX * p.ncbY;
  const int64_t total = (int64_t)p.batch * tiles_per;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int64_t bz = tile / tiles_per;
        const int rem = (int)(tile - bz * tiles_per), rbx = rem / p.ncbY, cby = rem - rbx * p.ncbY;
        for (int kb = 0; kb < p.nkb; ++kb) {
          const unsigned char* xs = p.X + (((bz * p.nrbX + rbx) * p.nkb + kb) * (int64_t)X_SLAB);
          const unsigned char* ys = p.Y + ((((bz / p.y_div) * p.ncbY + cby) * p.nkb + kb) * (int64_t)Y_SLAB);
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          bulk_g2s(ring + stage * STAGE_BYTES, xs, X_SLAB, full_bar(stage));
          bulk_g2s(ring + stage * STAGE_BYTES + X_SLAB, ys, Y_SLAB, full_bar(stage));
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0, accphase = 0;
      for (int64_t tile = blockIdx.x; tile < total; tile += gridDim.x) {
        mbar_wait(tempty_bar, accphase ^ 1u);          // the epilogue has drained the six accumulators
        tc_fence_after();
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t xs = ring + stage * STAGE_BYTES, ys = xs + X_SLAB;
#pragma unroll
          for (int ks = 0; ks < KS / 32; ++ks) {
#pragma unroll
            for (int t = 0; t < NSL; ++t) {
#pragma unroll
              for (int i = 0; i <= t; ++i) {
                const int j = t - i;
                const uint64_t xd = smem_desc(xs + (i >> 1) * X_PLANE + (i & 1) * 64 + ks * 32);
                const uint64_t yd = smem_desc(ys + (j >> 1) * Y_PLANE + (j & 1) * 64 + ks * 32);
                umma_i8(tmem_base + (uint32_t)t * 64u, xd, yd, IDESC_I8_128x64, (uint32_t)((kb | ks | i) != 0));
              }
            }
          }
          umma_commit(empty_bar(stage));
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar);
        accphase ^= 1u;
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int ew = warp - 2, h = ew >> 2;      // h: which half of the tile's 32 complex columns
    const int prow = 32 * q + lane;            // plane row of X = TMEM lane
    const int tid_e = h * 128 + prow;
    const int i = prow & 63, upper = prow >> 6;
    const int M = p.nrbX * XROWS, N = p.ncbY * YROWS;
    uint32_t accphase = 0;
    for (int64_t tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int64_t bz = tile / tiles_per;
      const int rem = (int)(tile - bz * tiles_per), rbx = rem / p.ncbY, cby = rem - rbx * p.ncbY;
      const int row = rbx * XROWS + i;
      const double rs = exp2i(p.ex_x[(bz * M + row) * 2 + upper] - 28);
      // column scales 2^ey of this tile, once per tile: s_cs[c] for Y re rows (c < 32) and im rows (c >= 32)
      epi_bar8();                              // everybody is done with the previous tile's s_cs / exchange buffer
      if (tid_e < 64) {
        const int* eyp = p.ex_y + ((bz / p.y_div) * N + cby * YROWS) * 2;
        s_cs[tid_e] = exp2i(eyp[2 * (tid_e & 31) + (tid_e >> 5)]);
      }
      mbar_wait(tfull_bar, accphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(16 * h);
      // acc[c], c < 16: (my X plane row) . (Y re row 16 h + c);  acc[16 + c]: . (Y im row 16 h + c).
      // The six levels are combined EXACTLY in 64-bit integers first (three levels per integer, IMAD.WIDE), so an output
      // costs two int64 -> double bit tricks and one DFMA instead of six conversions and six DFMAs:
      //     sum_t v_t 2^-7(t+2) = 2^-28 (hi + lo 2^-21),  hi = v0 2^14 + v1 2^7 + v2,  lo = v3 2^14 + v4 2^7 + v5
      double acc[32];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        int v[NSL][16];
#pragma unroll
        for (int t = 0; t < NSL; ++t) tmem_ld16_nowait(taddr + (uint32_t)t * 64u + (uint32_t)half * 32u, v[t]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const long long hi = (long long)v[0][c] * 16384 + ((long long)v[1][c] * 128 + (long long)v[2][c]);
          const long long lo = (long long)v[3][c] * 16384 + ((long long)v[4][c] * 128 + (long long)v[5][c]);
          acc[16 * half + c] = fma(ll2d(lo), 4.76837158203125e-07, ll2d(hi));        // 2^-21
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar);                 // this thread no longer reads the accumulators
      accphase ^= 1u;
      epi_bar8();                              // s_cs is complete
#pragma unroll
      for (int c = 0; c < 16; ++c) { acc[c] *= rs * s_cs[16 * h + c]; acc[16 + c] *= rs * s_cs[32 + 16 * h + c]; }   // rs carries the 2^-28
      // lower thread (Xr row): acc = [RR | RI];  upper thread (Xi row): acc = [IR | II]  (16 complex columns each).
      // lower keeps local columns 0..7 and gives RR, RI of columns 8..15; upper the other way round.
      // (static register indices on both sides of every select: a run-time offset would push acc[] to local memory)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        xch[g * 256 + tid_e] = upper ? make_double2(acc[2 * g], acc[2 * g + 1]) : make_double2(acc[8 + 2 * g], acc[8 + 2 * g + 1]);
        xch[(4 + g) * 256 + tid_e] = upper ? make_double2(acc[16 + 2 * g], acc[16 + 2 * g + 1])
                                           : make_double2(acc[24 + 2 * g], acc[24 + 2 * g + 1]);
      }
      epi_bar8();
      const int partner = tid_e ^ 64, keep0 = 16 * h + (upper ? 8 : 0);
      const double sgn = p.conj_y ? 1.0 : -1.0;
      double cre[8], cim[8], ssq = 0.0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const double2 p1 = xch[g * 256 + partner], p2 = xch[(4 + g) * 256 + partner];
        const double P1[2] = {p1.x, p1.y}, P2[2] = {p2.x, p2.y};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = 2 * g + e;
          double rr, ri, ir, ii;
          if (!upper) { rr = acc[c]; ri = acc[16 + c]; ir = P1[e]; ii = P2[e]; }
          else { ir = acc[8 + c]; ii = acc[24 + c]; rr = P1[e]; ri = P2[e]; }
          // no conj: Cr = RR - II, Ci = RI + IR;   conj(Y): Cr = RR + II, Ci = IR - RI
          cre[c] = rr + sgn * ii;
          cim[c] = ir - sgn * ri;
          ssq += cre[c] * cre[c] + cim[c] * cim[c];
        }
      }
      const int tile_in = rbx * p.ncbY + cby;
      const int col0 = cby * YROWS + keep0;
      if (p.out_c) {
        double2* o = reinterpret_cast<double2*>(p.out_c + ((bz * M + row) * (int64_t)N + col0));
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = make_double2(cre[c], cim[c]);
      }
      if (p.out_ct) {                          // lanes = consecutive rows: every store instruction writes 512 contiguous bytes
        double2* o = reinterpret_cast<double2*>(p.out_ct + ((bz * N + col0) * (int64_t)M + row));
#pragma unroll
        for (int c = 0; c < 8; ++c) o[(int64_t)c * M] = make_double2(cre[c], cim[c]);
      }
      double dr = 0.0, di = 0.0;
      if (p.dot_with) {
        const double2* w = reinterpret_cast<const double2*>(p.dot_with + ((bz * M + row) * (int64_t)N + col0));
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const double2 z = w[c];
          dr += z.x * cre[c] + z.y * cim[c];
          di += z.x * cim[c] - z.y * cre[c];
        }
      }
      if (p.norm_out || p.dot_out) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
          dr += __shfl_xor_sync(0xffffffffu, dr, o);
          di += __shfl_xor_sync(0xffffffffu, di, o);
        }
        if (lane == 0) { s_red[ew][0] = ssq; s_red[ew][1] = dr; s_red[ew][2] = di; }
        epi_bar8();
        if (tid_e == 0) {
          double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int k = 0; k < 8; ++k) { s0 += s_red[k][0]; s1 += s_red[k][1]; s2 += s_red[k][2]; }
          if (p.norm_out) p.norm_out[bz * tiles_per + tile_in] = s0;
          if (p.dot_out) p.dot_out[bz * tiles_per + tile_in] = mk<double>(s1, s2);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

#endif  // __CUDACC__

}  // namespace tci8
}  // namespace qmps
