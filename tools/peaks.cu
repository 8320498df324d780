// Measures the issue peaks the small-D kernels are bounded by on this B200:
//   FP64 FMA (DFMA), FP64 tensor (DMMA m8n8k4), FP32 FMA (FFMA), and a plain HBM copy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peaks tools/peaks.cu
// Output: one JSON object on stdout (written to profiles/ by tools/gpu_round.sh).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <typename T, int ILP>
__global__ void __launch_bounds__(256) fma_kernel(T* out, int iters, T a, T b) {
  T acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = T(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = acc[i] * a + b;
  }
  T s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == T(123456789)) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  double a = 1.0000001, b = 0.9999999;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  if (s == 123456789.0) out[0] = s;
}

__global__ void copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

template <typename F> float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  double* out; CK(cudaMalloc(&out, 1 << 20));
  const int iters = 4096, grid = sms * 8, block = 256;
  constexpr int ILP = 8;
  float ms64 = time_ms([&] { fma_kernel<double, ILP><<<grid, block>>>(out, iters, 1.0000001, 1e-9); }, 5);
  float ms32 = time_ms([&] { fma_kernel<float, ILP><<<grid, block>>>((float*)out, iters, 1.0000001f, 1e-9f); }, 5);
  float msmm = time_ms([&] { dmma_kernel<ILP><<<grid, block>>>(out, iters); }, 5);
  const double nfma = (double)grid * block * iters * ILP;
  const double tf64 = 2.0 * nfma / (ms64 * 1e-3) / 1e12, tf32 = 2.0 * nfma / (ms32 * 1e-3) / 1e12;
  // one m8n8k4 per warp = 8*8*4 FMA = 512 flops
  const double tfmm = (double)grid * (block / 32) * iters * ILP * 512.0 / (msmm * 1e-3) / 1e12;
  const size_t n = (size_t)1 << 28;  // 4 GiB in, 4 GiB out
  double2 *a, *b; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
  CK(cudaMemset(a, 1, n * 16));
  float msc = time_ms([&] { copy_kernel<<<sms * 16, 512>>>(a, b, n); }, 5);
  const double gbs = 2.0 * n * 16 / (msc * 1e-3) / 1e9;
  printf("{\"sms\": %d, \"fp64_fma_tflops\": %.2f, \"fp64_dmma_tflops\": %.2f, \"fp32_fma_tflops\": %.2f, \"copy_gbs\": %.1f, "
         "\"how\": \"tools/peaks.cu: %d CTAs x 256 thr, ILP 8 independent FMA chains x %d iters (best of 5); mma.sync.m8n8k4.f64; 4 GiB double2 grid-stride copy\"}\n",
         sms, tf64, tfmm, tf32, gbs, grid, iters);
  return 0;
}
