// microbench_tcgen05.cu -- ROUND-2 PREPARATION, NOT YET RUN ON HARDWARE (written at the end of round 1 after
// the GPU budget was spent; it compiles for sm_100a, nothing here is measured or used by the product).
//
// Question it is meant to answer first thing in round 2: can `tcgen05.mma kind::tf32` carry the 3xTF32 gate
// arithmetic of hq_mma.cuh with the tile's amplitudes as the A operand?  For that it
//   1. checks D[128 x N] = A[128 x 8] * B[N x 8]^T with A and B in the K-major, no-swizzle canonical layout
//      (16-byte units: A(r, c) at unit c * LBO + (r / 8) * SBO + r % 8, the layout DESIGN.md section 6 derives
//      for "rows = groups, K-chunks = the gate's target unit bits"), N = 8, 16, 32;
//   2. times back-to-back issue of such MMAs from one thread (cycles per MMA; the model in
//      B300_MICROARCH.md says max(M,128) * N / 256 cycles, i.e. 4 cycles at N = 8).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/microbench_tcgen05 tools/microbench_tcgen05.cu
//
// Descriptor encodings follow cute/arch/mma_sm100_desc.hpp (UMMA::SmemDescriptor, UMMA::InstrDescriptor) and
// cute/atom/mma_traits_sm100.hpp (make_umma_desc, K-major SWIZZLE_NONE: LBO = K-chunk stride, SBO = 8-row stride).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// shared-memory matrix descriptor: start address, LBO, SBO in 16-byte units; version 1 (Blackwell); no swizzle
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_units, uint32_t sbo_units) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr >> 4) & 0x3fffu);           // bits [0, 14)
  d |= uint64_t(lbo_units & 0x3fffu) << 16;            // bits [16, 30)
  d |= uint64_t(sbo_units & 0x3fffu) << 32;            // bits [32, 46)
  d |= uint64_t(1) << 46;                              // version_ = 1
  return d;                                            // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, dense, M = 128, N
__host__ __device__ inline uint32_t make_instr_desc(uint32_t n) {
  uint32_t d = 0;
  d |= 1u << 4;             // c_format = F32
  d |= 2u << 7;             // a_format = TF32
  d |= 2u << 10;            // b_format = TF32
  d |= (n >> 3) << 17;      // n_dim
  d |= (128u >> 4) << 24;   // m_dim
  return d;
}

__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u));
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// one CTA of 128 threads.  A: 128 x 8 floats (row-major in global), B: N x 8, D: 128 x N
template <int N>
__global__ void __launch_bounds__(128) k_check(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                               int iters, long long* cycles) {
  __shared__ __align__(128) float4 sA[2 * 128];     // unit (r, c) at c * 128 + r
  __shared__ __align__(128) float4 sB[2 * N];       // unit (n, c) at c * N + n
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int u = tid; u < 256; u += 128) {
    const int c = u / 128, r = u % 128;
    sA[u] = make_float4(A[r * 8 + 4 * c], A[r * 8 + 4 * c + 1], A[r * 8 + 4 * c + 2], A[r * 8 + 4 * c + 3]);
  }
  for (int u = tid; u < 2 * N; u += 128) {
    const int c = u / N, n = u % N;
    sB[u] = make_float4(B[n * 8 + 4 * c], B[n * 8 + 4 * c + 1], B[n * 8 + 4 * c + 2], B[n * 8 + 4 * c + 3]);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  // generic-proxy writes of the operands must be visible to the async proxy the tensor core reads through
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(32u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_holder;

  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint64_t da = make_smem_desc(smem_u32(sA), /*LBO*/ 128, /*SBO*/ 8);
    const uint64_t db = make_smem_desc(smem_u32(sB), /*LBO*/ N, /*SBO*/ 8);
    const uint32_t idesc = make_instr_desc(N);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) mma_tf32_ss(tmem, da, db, idesc, i > 0 ? 1u : 0u);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(smem_u32(&bar), 0);
  if (tid == 0) {
    t1 = clock64();
    if (cycles) cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // epilogue: warp w owns TMEM lanes 32 w .. 32 w + 31 = rows; 8 columns per load
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem + (uint32_t(warp * 32) << 16) + uint32_t(c0);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (blockIdx.x == 0)
      for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u));
}

static float tf32_trunc(float x) {
  uint32_t b;
  memcpy(&b, &x, 4);
  b &= 0xffffe000u;
  memcpy(&x, &b, 4);
  return x;
}

template <int N>
static void run(int sms) {
  std::vector<float> hA(128 * 8), hB(N * 8), hD(128 * N);
  for (size_t i = 0; i < hA.size(); ++i) hA[i] = float((int(i) * 7) % 13 - 6);          // exact in TF32
  for (size_t i = 0; i < hB.size(); ++i) hB[i] = float((int(i) * 5) % 11 - 5);
  float *dA, *dB, *dD;
  long long* dC;
  CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dD, hD.size() * 4));
  CK(cudaMalloc(&dC, sizeof(long long) * size_t(sms)));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  k_check<N><<<1, 128>>>(dA, dB, dD, 1, nullptr);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  double err = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < 8; ++k) s += double(tf32_trunc(hA[m * 8 + k])) * tf32_trunc(hB[n * 8 + k]);
      err = fmax(err, fabs(s - hD[m * N + n]));
    }
  const int iters = 4096;
  k_check<N><<<sms, 128>>>(dA, dB, dD, iters, dC);
  CK(cudaDeviceSynchronize());
  std::vector<long long> hC(static_cast<size_t>(sms));
  CK(cudaMemcpy(hC.data(), dC, sizeof(long long) * size_t(sms), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (long long c : hC) avg += double(c);
  avg /= sms;
  printf("{\"test\": \"tcgen05_tf32_m128_k8\", \"n\": %d, \"max_abs_err\": %.3e, \"ok\": %s, \"cycles_per_mma\": %.2f, "
         "\"mac_per_clk_per_sm\": %.1f}\n", N, err, err == 0 ? "true" : "false", avg / iters, 128.0 * N * 8 / (avg / iters));
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("{\"device\": \"%s\", \"note\": \"round-2 preparation tool, first run\"}\n", prop.name);
  run<8>(prop.multiProcessorCount);
  run<16>(prop.multiProcessorCount);
  run<32>(prop.multiProcessorCount);
  return 0;
}
