// microbench_fma.cu -- what can the B200 FMA pipes actually issue?  (round-1 evidence for the
// design of the fused-pass gate arithmetic; diagnostics only)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench_fma tools/microbench_fma.cu
//
// Variants: FFMA with three register operands, FFMA with one constant-bank operand,
// packed fma.rn.f32x2 (FFMA2), DFMA; 8 independent accumulator chains per thread, 1024 threads
// per SM x 148 SMs, enough iterations to swamp launch overhead.
#include <cuda_runtime.h>

#include <cstdio>

struct Consts { float c[64]; double d[64]; };

template <int MODE>
__global__ void __launch_bounds__(256) k_fma(float* out, int iters, float a0, float b0, const __grid_constant__ Consts cs) {
  float acc[8];
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i] = threadIdx.x * 1e-3f + i; a[i] = a0 + i * 1e-3f; b[i] = b0 + i * 1e-4f; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(a[(i + r) & 7]), "f"(b[i]));  // scalar, 3 regs
        if (MODE == 2) acc[i] = __fmaf_rn(a[(i + r) & 7], b[i], acc[i]);          // compiler's choice (may pack into FFMA2)
        if (MODE == 1) acc[i] = __fmaf_rn(acc[i], cs.c[r * 8 + i], b[i]);          // constant-bank operand
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_fma2(float* out, int iters, float a0, float b0) {
  unsigned long long acc[8], a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float2 v = make_float2(threadIdx.x * 1e-3f + i, 1.f + i);
    float2 x = make_float2(a0 + i * 1e-3f, a0 - i * 1e-3f);
    float2 y = make_float2(b0 + i * 1e-4f, b0 - i * 1e-4f);
    acc[i] = *reinterpret_cast<unsigned long long*>(&v);
    a[i] = *reinterpret_cast<unsigned long long*>(&x);
    b[i] = *reinterpret_cast<unsigned long long*>(&y);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(a[(i + r) & 7]), "l"(b[i]));
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float2 v = *reinterpret_cast<float2*>(&acc[i]);
    s += v.x + v.y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dfma(float* out, int iters, double a0, double b0) {
  double acc[8], a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i] = threadIdx.x * 1e-3 + i; a[i] = a0 + i * 1e-3; b[i] = b0 + i * 1e-4; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = __fma_rn(a[(i + r) & 7], b[i], acc[i]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = float(s);
}

__global__ void __launch_bounds__(256) k_dfma_const(float* out, int iters, double b0, const __grid_constant__ Consts cs) {
  double acc[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i] = threadIdx.x * 1e-3 + i; b[i] = b0 + i * 1e-4; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = __fma_rn(acc[i], cs.d[r * 8 + i], b[i]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = float(s);
}

template <typename F>
static double time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 4, threads = 256, iters = 20000;
  float* out;
  cudaMalloc(&out, size_t(blocks) * threads * sizeof(float));
  Consts cs;
  for (int i = 0; i < 64; ++i) { cs.c[i] = 1.0f - 1e-6f * i; cs.d[i] = 1.0 - 1e-9 * i; }
  const double fmas = double(blocks) * threads * double(iters) * 64.0;
  double ms;
  ms = time_ms([&] { k_fma<0><<<blocks, threads>>>(out, iters, 0.999f, 1e-3f, cs); });
  printf("{\"what\": \"FFMA 3-reg\", \"ms\": %.3f, \"TFMA_per_s\": %.2f, \"TFLOPs\": %.2f}\n", ms, fmas / ms / 1e9, 2 * fmas / ms / 1e9);
  ms = time_ms([&] { k_fma<2><<<blocks, threads>>>(out, iters, 0.999f, 1e-3f, cs); });
  printf("{\"what\": \"FFMA C++ (compiler may emit FFMA2)\", \"ms\": %.3f, \"TFMA_per_s\": %.2f, \"TFLOPs\": %.2f}\n", ms, fmas / ms / 1e9, 2 * fmas / ms / 1e9);
  ms = time_ms([&] { k_fma<1><<<blocks, threads>>>(out, iters, 0.999f, 1e-3f, cs); });
  printf("{\"what\": \"FFMA const-bank operand\", \"ms\": %.3f, \"TFMA_per_s\": %.2f, \"TFLOPs\": %.2f}\n", ms, fmas / ms / 1e9, 2 * fmas / ms / 1e9);
  ms = time_ms([&] { k_fma2<<<blocks, threads>>>(out, iters, 0.999f, 1e-3f); });
  printf("{\"what\": \"FFMA2 (fma.rn.f32x2)\", \"ms\": %.3f, \"TFMA_per_s\": %.2f, \"TFLOPs\": %.2f}\n", ms, 2 * fmas / ms / 1e9, 4 * fmas / ms / 1e9);
  ms = time_ms([&] { k_dfma<<<blocks, threads>>>(out, iters / 4, 0.999, 1e-3); });
  printf("{\"what\": \"DFMA\", \"ms\": %.3f, \"TFMA_per_s\": %.2f, \"TFLOPs\": %.2f}\n", ms, fmas / 4 / ms / 1e9, 2 * fmas / 4 / ms / 1e9);
  ms = time_ms([&] { k_dfma_const<<<blocks, threads>>>(out, iters / 4, 1e-3, cs); });
  printf("{\"what\": \"DFMA const-bank operand\", \"ms\": %.3f, \"TFMA_per_s\": %.2f, \"TFLOPs\": %.2f}\n", ms, fmas / 4 / ms / 1e9, 2 * fmas / 4 / ms / 1e9);
  cudaFree(out);
  return 0;
}
