// Measures the tcgen05 issue peaks the D >= 64 contraction kernels are bounded by on this B200:
//   kind::tf32 (cgemm_tc_kernel), kind::i8 (zgemm_i8_kernel) and kind::f16 with bf16 operands (cross-check against the
//   driver's cuBLAS bf16 figure in MEASURED_PEAKS.json).
// One CTA per SM; one thread issues back-to-back UMMAs of shape M = 128, N = 256, K = 32 bytes on operands that
// never leave shared memory, alternating between two TMEM accumulators; no loads, no epilogue.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I qmps_b200/csrc -o tools/peaks_tc tools/peaks_tc.cu
// Output: one JSON object on stdout.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "kernels_tc.cuh"

#define CKM(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

using namespace qmps::tc;

// KIND 0: tf32, 1: i8, 2: bf16 (kind::f16)
template <int KIND>
__device__ __forceinline__ void umma_any(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else if (KIND == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) umma_peak_kernel(int iters, unsigned seed) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const uint32_t dyn0 = smem_u32(smem_dyn);
  const uint32_t base = (dyn0 + 1023u) & ~1023u;
  unsigned char* ops = smem_dyn + (base - dyn0);
  // operands: A = 128 rows x 128 B (16 KB), B = 256 rows x 128 B (32 KB); small finite values of the operand type
  unsigned s = seed + threadIdx.x * 2654435761u + blockIdx.x;
  for (int i = threadIdx.x; i < (48 * 1024) / 4; i += blockDim.x) {
    s = s * 1664525u + 1013904223u;
    uint32_t w;
    if (KIND == 0) w = __float_as_uint(((int)(s >> 20) - 2048) * (1.0f / 2048.0f)) & 0xffffe000u;
    else if (KIND == 1) w = s & 0x3f3f3f3fu;
    else { const uint32_t h = __float_as_uint(((int)(s >> 20) - 2048) * (1.0f / 2048.0f)) >> 16; w = h | (h << 16); }
    reinterpret_cast<uint32_t*>(ops)[i] = w;
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&s_bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&s_tmem), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  // D type: F32 (1 << 4) or S32 (2 << 4); A/B type: TF32 = 2, S8 = 1 (kind::i8), BF16 = 1 (kind::f16)
  const uint32_t idesc = (KIND == 0 ? ((1u << 4) | (2u << 7) | (2u << 10)) : KIND == 1 ? ((2u << 4) | (1u << 7) | (1u << 10))
                                                                                       : ((1u << 4) | (1u << 7) | (1u << 10))) |
                         ((256u >> 3) << 17) | ((128u >> 4) << 24);
  if (threadIdx.x == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {                                   // four K steps of 32 B inside the 128-byte swizzle row
        const uint64_t ad = smem_desc(base + ks * 32), bd = smem_desc(base + 16384 + ks * 32);
        umma_any<KIND>(tmem + (uint32_t)(it & 1) * 256u, ad, bd, idesc, (uint32_t)(it > 1 || ks > 0));
      }
    }
    umma_commit(smem_u32(&s_bar));
    mbar_wait(smem_u32(&s_bar), 0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <typename F> float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  CKM(cudaEventCreate(&e0)); CKM(cudaEventCreate(&e1));
  f(); f();
  CKM(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CKM(cudaEventRecord(e0));
    f();
    CKM(cudaEventRecord(e1));
    CKM(cudaEventSynchronize(e1));
    float ms; CKM(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

template <int KIND> double run(int sms, int iters, int kelems) {
  const int smem = 48 * 1024 + 1024;
  CKM(cudaFuncSetAttribute(umma_peak_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const float ms = time_ms([&] { umma_peak_kernel<KIND><<<sms, 128, smem>>>(iters, 12345u); CKM(cudaGetLastError()); }, 5);
  const double ops = (double)sms * iters * 4.0 * 2.0 * 128.0 * 256.0 * kelems;
  return ops / (ms * 1e-3) / 1e12;
}

int main() {
  int sms = 0; CKM(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int iters = 1 << 16;
  const double tf32 = run<0>(sms, iters, 8), i8 = run<1>(sms, iters, 32), bf16 = run<2>(sms, iters, 16);
  printf("{\"tf32_tcgen05_tflops\": %.1f, \"i8_tcgen05_tops\": %.1f, \"bf16_tcgen05_tflops\": %.1f, "
         "\"how_tc\": \"tools/peaks_tc.cu: %d CTAs, one thread issuing %d x 4 tcgen05.mma cta_group::1 M128 N256 K32B on resident shared-memory operands, two TMEM accumulators (best of 5)\"}\n",
         tf32, i8, bf16, sms, iters);
  return 0;
}
