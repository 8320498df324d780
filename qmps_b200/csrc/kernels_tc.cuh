// qmps_b200 large-D transfer-matrix application in complex64 on the 5th-generation tensor cores
// (BASELINE config 5; qmps.ipynb cells 29-32; SURVEY 8(d) cfg 5):
//
//     stage 1   T[b,s] = alpha_b * A[b,s] . r[b]                X = A_s   (rows i, K = k),  Y = r^T (rows j, K = k)
//     stage 2   r'[b]  = sum_s T[b,s] . B[b,s]^dagger           X = T_s   (rows i, K = j),  Y = B_s (rows l, K = j)
//
// Both stages are the same batched complex product  C = sum_t X_t . Y_t^T  (or Y_t^H)  with X and Y
// stored K-major, so one kernel serves both.
//
// Arithmetic.  tcgen05.mma has no FP32 kind; kind::tf32 keeps 10 mantissa bits.  Every real
// operand is split x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi), and each product is
// issued as hi.hi + hi.lo + lo.hi (3xTF32, the dropped lo.lo term is 2^-22 relative): FP32-grade
// results (tolerance 1e-5 of the complex64 mode) at one third of the TF32 tensor rate.
//
// Complex -> real.  An operand block of 64 complex rows is stored as 128 plane rows
// [re rows 0..63 ; im rows 64..127].  One UMMA of shape M = 128, N = 128, K = 8 on a pair of
// blocks gives all four real products at once in a 128-lane x 128-column TMEM accumulator:
//        lanes  0..63 :  [ Xr.Yr^T | Xr.Yi^T ]          (RR | RI)
//        lanes 64..127:  [ Xi.Yr^T | Xi.Yi^T ]          (IR | II)
// The epilogue exchanges half of each row between the lane pairs (i, i + 64) through shared
// memory and forms  Cr = RR -+ II,  Ci = RI +- IR.
//
// Memory.  Operands live in global memory as "slab images": for every (matrix, row block of 64,
// K slab of 32) ONE 16 KB plane of fp32 values laid out EXACTLY as the UMMA K-major
// SWIZZLE_128B shared-memory tile (8-row x 128-byte atoms, 16-byte chunk c of row r stored at
// chunk c ^ (r & 7), 1024 bytes between 8-row groups).  A slab therefore moves with ONE bulk
// copy of the TMA engine (cp.async.bulk, 16 KB, mbarrier complete_tx) and needs no tensor map.
// The hi / lo TF32 split happens in SHARED memory after the copy (four splitter warps, element-wise and in place,
// so the swizzle is untouched): round 1 stored both planes in global memory and D = 64 was bound by that traffic
// (2.6x the algorithmic bytes, profiles/ncu_tc_canon_r01e.txt).
// The epilogue writes its result directly as the slab image the next stage reads (natural for
// T, transposed for r'), so r never leaves the image form between applications.
//
// Kernel structure (persistent, warp-specialised, 320 threads, 1 CTA per SM):
//     warp 0   lane 0: producer  -- bulk copies (2 x 16 KB) into a 3-stage ring (64 KB per stage once split)
//     warps 6-9      : splitter  -- fp32 plane -> hi (in place) + lo planes, fence.proxy.async, hand-over to the issuer
//     warp 1   lane 0: MMA issuer -- 4 K-steps x 3 UMMAs per slab, tcgen05.commit frees the stage
//     warps 2-5      : epilogue  -- tcgen05.ld, pair exchange, scale / store, norms
// Two TMEM accumulators (2 x 128 columns) let the epilogue of tile n overlap the MMAs of n + 1.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "core.cuh"

namespace qmps {
namespace tc {

constexpr int ROWS = 64;                       // complex rows per operand block
constexpr int KS = 32;                         // K elements per slab (= one 128-byte swizzle row)
constexpr int PLANE_BYTES = 128 * 128;         // 128 plane rows x 128 B
constexpr int IMG_SLAB_BYTES = PLANE_BYTES;    // a slab in GLOBAL memory: one fp32 plane (round 2: half the image traffic)
constexpr int SLAB_BYTES = 2 * PLANE_BYTES;    // a slab in SHARED memory: hi + lo TF32 planes, split after the bulk copy
constexpr int STAGE_BYTES = 2 * SLAB_BYTES;    // X slab + Y slab
constexpr int NSTAGE = 3;                        // 3 x 64 KB in flight per CTA (D = 64 is bound by image traffic)
constexpr int XCH_BYTES = 128 * 64 * 4;
constexpr int THREADS = 320;                     // producer, MMA issuer, 4 epilogue warps, 4 splitter warps
constexpr int TMEM_COLS = 256;
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + XCH_BYTES + 1024;   // + slack for the 1024-byte alignment

// byte offset of element (plane row, k) inside a 16 KB plane
QMPS_HD uint32_t img_off(int prow, int k) {
  return (uint32_t)((prow >> 3) * 1024 + (prow & 7) * 128 + ((((k >> 2) ^ prow) & 7) << 4) + (k & 3) * 4);
}

#if defined(__CUDACC__)

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  const float rem = x - hi;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(rem));
  lo = __uint_as_float(l);
}

// ---- pack: interleaved complex64 matrices -> slab images --------------------------------------
// element (row, k) of matrix m is in[m * mstride + row * rstride + k * kstride]
static __global__ void __launch_bounds__(256)
pack_kernel(int64_t nmat, int R, int K, const cx<float>* __restrict__ in, int64_t mstride, int64_t rstride,
            int64_t kstride, unsigned char* __restrict__ img, int presplit) {
  const int nrb = R / ROWS, nkb = K / KS;
  const int64_t total = nmat * (int64_t)R * K;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(e % K);
    const int64_t q = e / K;
    const int row = (int)(q % R);
    const int64_t m = q / R;
    const cx<float> z = in[m * mstride + row * rstride + k * kstride];
    const int rb = row / ROWS, i = row % ROWS, kb = k / KS, kk = k % KS;
    unsigned char* base = img + (((m * nrb + rb) * nkb + kb) * (int64_t)(presplit ? SLAB_BYTES : IMG_SLAB_BYTES));
    if (presplit) {
      float hi, lo;
      split_tf32(z.re, hi, lo);
      *reinterpret_cast<float*>(base + img_off(i, kk)) = hi;
      *reinterpret_cast<float*>(base + PLANE_BYTES + img_off(i, kk)) = lo;
      split_tf32(z.im, hi, lo);
      *reinterpret_cast<float*>(base + img_off(64 + i, kk)) = hi;
      *reinterpret_cast<float*>(base + PLANE_BYTES + img_off(64 + i, kk)) = lo;
    } else {
      *reinterpret_cast<float*>(base + img_off(i, kk)) = z.re;
      *reinterpret_cast<float*>(base + img_off(64 + i, kk)) = z.im;
    }
  }
}

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// A wait that cannot hang the GPU: after ~4 s the kernel traps (the host sees a launch failure).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
    if ((spin & 1023u) == 1023u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000ll) __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, kind::tf32, issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive columns: thread l of the warp receives lane (base + l)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
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
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100 format):
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1)
//   [32,46) stride byte offset >> 4 (1024 B between 8-row groups)   [46,48) version = 1
//   [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major,
// N = 128 (16 << 17), M = 128 (8 << 24)
constexpr uint32_t IDESC_TF32_128x128 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

struct Params {
  const unsigned char* X;      // slab images, matrix index  bz * nsum + t
  const unsigned char* Y;      // slab images, matrix index (bz / y_div) * nsum + t
  int nsum, nkb, nrbX, nrbY, y_div, batch, conj_y;
  int presplit;                // 1: global images hold hi + lo planes (32 KB per slab, no in-kernel split): the large-K mode

  const float* norm_in; int n_in, a_div;   // alpha = rsqrt(sum_j norm_in[(bz / a_div) * n_in + j]); nullptr -> 1
  float* norm_out;             // [bz][tile] partial sums of |C|^2 after alpha
  unsigned char* out_img;      // next-stage image of C (matrix index bz)
  int out_mode;                // 1: rows = C rows, K = C columns;  2: rows = C columns, K = C rows
  int out_nrb, out_nkb;
  cx<float>* out_c;            // interleaved C[bz][M][N]
  const cx<float>* dot_with;   // dot_out[bz][tile] = sum conj(dot_with[bz][i][l]) C[i][l]
  cx<float>* dot_out;
};

static __global__ void __launch_bounds__(THREADS, 1)
cgemm_tc_kernel(Params p) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t s_bar[3 * NSTAGE + 4];
  __shared__ uint32_t s_tmem;
  __shared__ float s_red[4][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t dyn0 = smem_u32(smem_dyn);
  const uint32_t ring = (dyn0 + 1023u) & ~1023u;
  float4* xch = reinterpret_cast<float4*>(smem_dyn + (ring - dyn0) + NSTAGE * STAGE_BYTES);
  const uint32_t bar0 = smem_u32(s_bar);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * NSTAGE + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * NSTAGE + 2 + a); };
  auto conv_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 4 + s); };   // the splitter warps are done with stage s

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); mbar_init(conv_bar(s), 128); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&s_tmem), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  const int tiles_per = p.nrbX * p.nrbY;
  const int64_t total = (int64_t)p.batch * tiles_per;
  const int iters = p.nsum * p.nkb;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int64_t bz = tile / tiles_per;
        const int rem = (int)(tile - bz * tiles_per), rbx = rem / p.nrbY, rby = rem - rbx * p.nrbY;
        for (int it = 0; it < iters; ++it) {
          const int t = it / p.nkb, kb = it - t * p.nkb;
          const int isb = p.presplit ? SLAB_BYTES : IMG_SLAB_BYTES;
          const unsigned char* xs = p.X + ((((bz * p.nsum + t) * p.nrbX + rbx) * p.nkb + kb) * (int64_t)isb);
          const unsigned char* ys = p.Y + (((((bz / p.y_div) * p.nsum + t) * p.nrbY + rby) * p.nkb + kb) * (int64_t)isb);
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), 2 * isb);
          bulk_g2s(ring + stage * STAGE_BYTES, xs, isb, full_bar(stage));                 // fp32 plane: lands where hi will be
          bulk_g2s(ring + stage * STAGE_BYTES + SLAB_BYTES, ys, isb, full_bar(stage));
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t accphase = 0;
      for (int64_t tile = blockIdx.x; tile < total; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), accphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 128u;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(p.presplit ? full_bar(stage) : conv_bar(stage), phase);
          tc_fence_after();
          const uint32_t xs = ring + stage * STAGE_BYTES, ys = xs + SLAB_BYTES;
#pragma unroll
          for (int ks = 0; ks < KS / 8; ++ks) {
            const uint64_t xh = smem_desc(xs + ks * 32), xl = smem_desc(xs + PLANE_BYTES + ks * 32);
            const uint64_t yh = smem_desc(ys + ks * 32), yl = smem_desc(ys + PLANE_BYTES + ks * 32);
            umma_tf32(d_tmem, xl, yh, IDESC_TF32_128x128, (it | ks) != 0);
            umma_tf32(d_tmem, xh, yl, IDESC_TF32_128x128, 1u);
            umma_tf32(d_tmem, xh, yh, IDESC_TF32_128x128, 1u);
          }
          umma_commit(empty_bar(stage));
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) accphase ^= 1u;
      }
    }
    __syncwarp();
  } else if (warp >= 6) {
    // splitter warps: the bulk copy delivered fp32 planes; split every element x = hi + lo (TF32) in place -- hi stays
    // where it landed, lo goes to the plane behind it -- then hand the stage to the MMA issuer.  Element-wise, so the
    // SWIZZLE_128B layout is preserved.
    const int tcv = threadIdx.x - 192;         // 0 .. 127
    unsigned char* ring_g = smem_dyn + (ring - dyn0);
    int stage = 0; uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < total && !p.presplit; tile += gridDim.x) {
      for (int it = 0; it < iters; ++it) {
        mbar_wait(full_bar(stage), phase);
#pragma unroll
        for (int op = 0; op < 2; ++op) {
          float4* hi4 = reinterpret_cast<float4*>(ring_g + stage * STAGE_BYTES + op * SLAB_BYTES);
          float4* lo4 = reinterpret_cast<float4*>(ring_g + stage * STAGE_BYTES + op * SLAB_BYTES + PLANE_BYTES);
#pragma unroll
          for (int j = 0; j < PLANE_BYTES / 16 / 128; ++j) {
            const int idx = tcv + 128 * j;
            const float4 x = hi4[idx];
            float4 h, l;
            split_tf32(x.x, h.x, l.x); split_tf32(x.y, h.y, l.y); split_tf32(x.z, h.z, l.z); split_tf32(x.w, h.w, l.w);
            hi4[idx] = h; lo4[idx] = l;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        mbar_arrive(conv_bar(stage));
        if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int prow = 32 * q + lane;            // plane row = TMEM lane
    const int i = prow & 63, upper = prow >> 6;
    const int M = p.nrbX * ROWS, N = p.nrbY * ROWS;
    int acc = 0; uint32_t accphase = 0;
    for (int64_t tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int64_t bz = tile / tiles_per;
      const int rem = (int)(tile - bz * tiles_per), rbx = rem / p.nrbY, rby = rem - rbx * p.nrbY;
      float alpha = 1.0f;
      if (p.norm_in) {
        const float* ni = p.norm_in + (bz / p.a_div) * p.n_in;
        float s = 0.0f;
        for (int j = 0; j < p.n_in; ++j) s += ni[j];
        alpha = rsqrtf(s);
      }
      mbar_wait(tfull_bar(acc), accphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)acc * 128u;
      const uint32_t give0 = upper ? 0u : 32u, keep0 = upper ? 32u : 0u;
      float a[32], b[32];
      epi_bar();                               // the exchange buffer is free again
      tmem_ld32(taddr + give0, a);
#pragma unroll
      for (int g = 0; g < 8; ++g) xch[g * 128 + prow] = make_float4(a[4 * g], a[4 * g + 1], a[4 * g + 2], a[4 * g + 3]);
      tmem_ld32(taddr + give0 + 64u, a);
#pragma unroll
      for (int g = 0; g < 8; ++g) xch[(8 + g) * 128 + prow] = make_float4(a[4 * g], a[4 * g + 1], a[4 * g + 2], a[4 * g + 3]);
      epi_bar();
      tmem_ld32(taddr + keep0, a);
      tmem_ld32(taddr + keep0 + 64u, b);
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));            // this thread no longer reads the accumulator
      acc ^= 1;
      if (acc == 0) accphase ^= 1u;
      // lower thread (rows of Xr): a = RR, b = RI, partner gave P1 = IR, P2 = II   (columns 0..31)
      // upper thread (rows of Xi): a = IR, b = II, partner gave P1 = RR, P2 = RI   (columns 32..63)
      const int partner = prow ^ 64;
      const float sgn = p.conj_y ? 1.0f : -1.0f;    // Cr = RR + sgn * II ... Ci = -sgn * RI ... see below
      float ssq = 0.0f;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 p1 = xch[g * 128 + partner], p2 = xch[(8 + g) * 128 + partner];
        const float P1[4] = {p1.x, p1.y, p1.z, p1.w}, P2[4] = {p2.x, p2.y, p2.z, p2.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = 4 * g + e;
          float rr, ri, ir, ii;
          if (!upper) { rr = a[c]; ri = b[c]; ir = P1[e]; ii = P2[e]; }
          else { ir = a[c]; ii = b[c]; rr = P1[e]; ri = P2[e]; }
          // no conj: Cr = RR - II, Ci = RI + IR;   conj(Y): Cr = RR + II, Ci = IR - RI
          const float cr = (rr + sgn * ii) * alpha;
          const float ci = (ir - sgn * ri) * alpha;
          a[c] = cr; b[c] = ci;
          ssq += cr * cr + ci * ci;
        }
      }
      const int tile_in = rbx * p.nrbY + rby;
      const int row = rbx * ROWS + i, col0 = rby * ROWS + 32 * upper;
      if (p.out_c) {
        float4* o = reinterpret_cast<float4*>(p.out_c + ((bz * M + row) * (int64_t)N + col0));
#pragma unroll
        for (int g = 0; g < 16; ++g) o[g] = make_float4(a[2 * g], b[2 * g], a[2 * g + 1], b[2 * g + 1]);
      }
      float dr = 0.0f, di = 0.0f;
      if (p.dot_with) {
        const float4* w = reinterpret_cast<const float4*>(p.dot_with + ((bz * M + row) * (int64_t)N + col0));
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          const float4 z = w[g];
          dr += z.x * a[2 * g] + z.y * b[2 * g] + z.z * a[2 * g + 1] + z.w * b[2 * g + 1];
          di += z.x * b[2 * g] - z.y * a[2 * g] + z.z * b[2 * g + 1] - z.w * a[2 * g + 1];
        }
      }
      if (p.out_img && p.out_mode == 1) {
        // rows = C rows (block rbx), K = C columns: my 32 columns are exactly K slab 2*rby + upper
        unsigned char* base = p.out_img + ((((bz * p.out_nrb + rbx) * p.out_nkb) + 2 * rby + upper) * (int64_t)(p.presplit ? SLAB_BYTES : IMG_SLAB_BYTES));
        unsigned char* rre = base + (i >> 3) * 1024 + (i & 7) * 128;
        unsigned char* rim = rre + 8 * 1024;                 // plane row 64 + i
        if (p.presplit) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int pos = ((g ^ i) & 7) << 4;
            float h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_tf32(a[4 * g + e], h[e], l[e]);
            *reinterpret_cast<float4*>(rre + pos) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(rre + PLANE_BYTES + pos) = make_float4(l[0], l[1], l[2], l[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) split_tf32(b[4 * g + e], h[e], l[e]);
            *reinterpret_cast<float4*>(rim + pos) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(rim + PLANE_BYTES + pos) = make_float4(l[0], l[1], l[2], l[3]);
          }
        } else {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int pos = ((g ^ i) & 7) << 4;
            *reinterpret_cast<float4*>(rre + pos) = make_float4(a[4 * g], a[4 * g + 1], a[4 * g + 2], a[4 * g + 3]);
            *reinterpret_cast<float4*>(rim + pos) = make_float4(b[4 * g], b[4 * g + 1], b[4 * g + 2], b[4 * g + 3]);
          }
        }
      } else if (p.out_img && p.out_mode == 2) {
        // rows = C columns (block rby), K = C rows: my row i sits at k = i & 31 of K slab 2*rbx + (i >> 5)
        unsigned char* base = p.out_img + ((((bz * p.out_nrb + rby) * p.out_nkb) + 2 * rbx + (i >> 5)) * (int64_t)(p.presplit ? SLAB_BYTES : IMG_SLAB_BYTES));
        const int k = i & 31;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int n = 32 * upper + c;
          const uint32_t ore = img_off(n, k), oim = img_off(64 + n, k);
          if (p.presplit) {
            float h, l;
            split_tf32(a[c], h, l);
            *reinterpret_cast<float*>(base + ore) = h;
            *reinterpret_cast<float*>(base + PLANE_BYTES + ore) = l;
            split_tf32(b[c], h, l);
            *reinterpret_cast<float*>(base + oim) = h;
            *reinterpret_cast<float*>(base + PLANE_BYTES + oim) = l;
          } else {
            *reinterpret_cast<float*>(base + ore) = a[c];
            *reinterpret_cast<float*>(base + oim) = b[c];
          }
        }
      }
      if (p.norm_out || p.dot_out) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
          dr += __shfl_xor_sync(0xffffffffu, dr, o);
          di += __shfl_xor_sync(0xffffffffu, di, o);
        }
        if (lane == 0) { s_red[q][0] = ssq; s_red[q][1] = dr; s_red[q][2] = di; }
        epi_bar();
        if (prow == 0) {
          // fixed summation order: deterministic
          const float s = (s_red[0][0] + s_red[1][0]) + (s_red[2][0] + s_red[3][0]);
          if (p.norm_out) p.norm_out[bz * tiles_per + tile_in] = s;
          if (p.dot_out)
            p.dot_out[bz * tiles_per + tile_in] = mk<float>((s_red[0][1] + s_red[1][1]) + (s_red[2][1] + s_red[3][1]),
                                                            (s_red[0][2] + s_red[1][2]) + (s_red[2][2] + s_red[3][2]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// r[b] *= 1/sqrt(sum_j norm[b*n + j])
static __global__ void __launch_bounds__(256)
scale_by_norm_kernel(int64_t len, cx<float>* __restrict__ r, const float* __restrict__ norm, int n) {
  float s = 0.0f;
  for (int j = 0; j < n; ++j) s += norm[(size_t)blockIdx.x * n + j];
  const float a = rsqrtf(s);
  cx<float>* p = r + (size_t)blockIdx.x * len;
  for (int64_t i = threadIdx.x; i < len; i += blockDim.x) p[i] = p[i] * a;
}
// out[b] = sum_j part[b*n + j]
static __global__ void sum_partials_kernel(int64_t nb, const cx<float>* __restrict__ part, int n, cx<float>* __restrict__ out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  cx<float> s = mk<float>(0.f, 0.f);
  for (int j = 0; j < n; ++j) s = s + part[b * n + j];
  out[b] = s;
}

#endif  // __CUDACC__

}  // namespace tc
}  // namespace qmps
