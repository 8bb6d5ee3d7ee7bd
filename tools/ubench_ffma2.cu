// Micro-benchmark: FP32 FMA issue rate on sm_100a -- scalar FFMA (3-register form) vs packed FFMA2
// (fma.rn.f32x2), 8 independent accumulator chains per thread, enough warps to hide the 4-cycle latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_ffma2 tools/ubench_ffma2.cu && /tmp/ubench_ffma2
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, CH = 8;

__global__ void k_ffma(float* out, float a, float b) {
  float acc[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  float x = a + threadIdx.x * 1e-6f, y = b + threadIdx.x * 1e-7f;   // per-thread operands: the 3-register form
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = fmaf(acc[i], x, y);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__global__ void k_ffma2(float* out, float a, float b) {
  unsigned long long acc[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = pack2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  const unsigned long long x = pack2(a + threadIdx.x * 1e-6f, a - threadIdx.x * 1e-6f), y = pack2(b + threadIdx.x * 1e-7f, b - threadIdx.x * 1e-7f);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = fma2(acc[i], x, y);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// MUFU rate: ex2 + rcp chains
__global__ void k_mufu(float* out, float a) {
  float acc[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = a + threadIdx.x * 1e-4f + i * 0.01f;
  for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      float e;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(acc[i]));
      asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(acc[i]) : "f"(e));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) f();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int blocks = sms * 4, threads = 512;
  float* out; cudaMalloc(&out, (size_t)blocks * threads * sizeof(float));
  const double n = (double)blocks * threads * ITERS * CH;
  float t1 = time_ms([&] { k_ffma<<<blocks, threads>>>(out, 0.999f, 1e-3f); });
  float t2 = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 0.999f, 1e-3f); });
  float t3 = time_ms([&] { k_mufu<<<blocks, threads>>>(out, 0.5f); });
  const double clk = khz * 1e3;
  printf("{\"sms\": %d, \"clock_mhz\": %.0f, \"ffma_ms\": %.4f, \"ffma_tflops\": %.2f, \"ffma_fma_per_clk_per_sm\": %.1f, "
         "\"ffma2_ms\": %.4f, \"ffma2_tflops\": %.2f, \"ffma2_fma_per_clk_per_sm\": %.1f, "
         "\"mufu_ms\": %.4f, \"mufu_per_clk_per_sm\": %.2f, \"err\": \"%s\"}\n",
         sms, khz / 1e3, t1, 2 * n / t1 / 1e9, n / (t1 * 1e-3) / clk / sms, t2, 4 * n / t2 / 1e9, 2 * n / (t2 * 1e-3) / clk / sms,
         t3, (double)blocks * threads * (ITERS / 4) * CH * 2 / (t3 * 1e-3) / clk / sms, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
