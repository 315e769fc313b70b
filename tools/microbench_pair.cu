// microbench_pair.cu -- does applying TWO k = 2 complex64 gates per shared-memory round trip pay?
// (round-1 evidence for the paired fast slots of the tile kernel; diagnostics only, JSON lines)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hybridq_b200/csrc \
//        -o tools/bin/microbench_pair tools/microbench_pair.cu
//
// A 64 KiB tile (4096 units of 2 amplitudes) sits in shared memory, no global traffic in the timed part.
//   single  one gate per round trip: a thread owns 4 units (the 2 target unit bits) x 4 iterations
//   pair    two gates on disjoint unit bits per round trip: a thread owns 16 units, applies gate A to the
//           4 groups along A's bits, then gate B to the 4 groups along B's bits, all in registers
// Matrix elements are kernel parameters (constant bank -> uniform registers), arithmetic is FFMA2.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

static inline uint32_t h_swz(uint32_t u) { return u ^ ((u >> 3) & 7u) ^ ((u >> 6) & 7u) ^ ((u >> 9) & 7u); }
__device__ __forceinline__ uint32_t d_swz(uint32_t u) { return u ^ ((u >> 3) & 7u) ^ ((u >> 6) & 7u) ^ ((u >> 9) & 7u); }

struct F2 { float lo, hi; };
__device__ __forceinline__ void ffma2_bcast(F2& acc, float xlo, float xhi, float s) {
  unsigned long long a, b, c;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(xlo), "f"(xhi));
  asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(s));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc.lo), "f"(acc.hi));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.lo), "=f"(acc.hi) : "l"(c));
}

#define NG 16
struct Params {
  float u[NG][32];            // row-major 4x4 complex
  unsigned short xo[NG][4];   // swizzled XOR offsets of the gate's 4 units (index m = 2 bits)
  unsigned short iter[NG][4]; // single: offsets of the 4 iterations
};

// out = U in for the 4 units `in` (each unit = even / odd amplitude), P/Q accumulators: no rotated copies
__device__ __forceinline__ void apply4(const float4 (&in)[4], float4 (&out)[4], const float* u) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    F2 pe = {0.f, 0.f}, qe = {0.f, 0.f}, po = {0.f, 0.f}, qo = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float ur = u[2 * (i * 4 + j)], ui = u[2 * (i * 4 + j) + 1];
      ffma2_bcast(pe, in[j].x, in[j].y, ur);
      ffma2_bcast(qe, in[j].x, in[j].y, ui);
      ffma2_bcast(po, in[j].z, in[j].w, ur);
      ffma2_bcast(qo, in[j].z, in[j].w, ui);
    }
    out[i] = make_float4(pe.lo - qe.hi, pe.hi + qe.lo, po.lo - qo.hi, po.hi + qo.lo);
  }
}

template <int OCC>
__global__ void __launch_bounds__(256, OCC) k_single(float4* gtile, const unsigned short* __restrict__ tbl_thread,
                                                     const __grid_constant__ Params p, int reps, int io) {
  extern __shared__ __align__(16) unsigned char smem[];
  float4* tile = reinterpret_cast<float4*>(smem);
  const int tid = threadIdx.x;
  if (io) { for (int u = tid; u < 4096; u += 256) tile[d_swz(u)] = gtile[u]; }
  else { for (int u = tid; u < 4096; u += 256) tile[u] = make_float4(u * 1e-4f, tid * 1e-4f, 1e-3f, 2e-3f); }
  __syncthreads();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const uint32_t st = __ldg(&tbl_thread[g * 256 + tid]);
#pragma unroll 1
      for (int it = 0; it < 4; ++it) {
        const uint32_t sb = st ^ p.iter[g][it];
        float4 in[4], out[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) in[m] = tile[sb ^ p.xo[g][m]];
        apply4(in, out, p.u[g]);
#pragma unroll
        for (int m = 0; m < 4; ++m) tile[sb ^ p.xo[g][m]] = out[m];
      }
      __syncthreads();
    }
  }
  if (io) { for (int u = tid; u < 4096; u += 256) gtile[u] = tile[d_swz(u)]; }
  else if (tile[tid].x == 123.456f) gtile[tid] = tile[tid];
}

// pair (2 pr, 2 pr + 1): tbl_thread holds the slot of the thread's only work item
template <int OCC>
__global__ void __launch_bounds__(256, OCC) k_pair(float4* gtile, const unsigned short* __restrict__ tbl_thread,
                                                   const __grid_constant__ Params p, int reps, int io) {
  extern __shared__ __align__(16) unsigned char smem[];
  float4* tile = reinterpret_cast<float4*>(smem);
  const int tid = threadIdx.x;
  if (io) { for (int u = tid; u < 4096; u += 256) tile[d_swz(u)] = gtile[u]; }
  else { for (int u = tid; u < 4096; u += 256) tile[u] = make_float4(u * 1e-4f, tid * 1e-4f, 1e-3f, 2e-3f); }
  __syncthreads();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int pr = 0; pr < NG / 2; ++pr) {
      const uint32_t sb = __ldg(&tbl_thread[pr * 256 + tid]);
      float4 v[4][4];   // [mb][ma]
#pragma unroll
      for (int mb = 0; mb < 4; ++mb) {
        float4 in[4];
#pragma unroll
        for (int ma = 0; ma < 4; ++ma) in[ma] = tile[sb ^ p.xo[2 * pr][ma] ^ p.xo[2 * pr + 1][mb]];
        apply4(in, v[mb], p.u[2 * pr]);
      }
#pragma unroll
      for (int ma = 0; ma < 4; ++ma) {
        float4 in[4], out[4];
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) in[mb] = v[mb][ma];
        apply4(in, out, p.u[2 * pr + 1]);
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) tile[sb ^ p.xo[2 * pr][ma] ^ p.xo[2 * pr + 1][mb]] = out[mb];
      }
      __syncthreads();
    }
  }
  if (io) { for (int u = tid; u < 4096; u += 256) gtile[u] = tile[d_swz(u)]; }
  else if (tile[tid].x == 123.456f) gtile[tid] = tile[tid];
}

typedef std::complex<double> cd;
static uint32_t scatter(uint32_t w, const std::vector<int>& pos) {
  uint32_t u = 0;
  for (size_t i = 0; i < pos.size(); ++i) u |= ((w >> i) & 1u) << pos[i];
  return u;
}
// free unit bits ordered so that the three lowest have distinct residues mod 3
static std::vector<int> lane_order(const std::vector<int>& freeb) {
  std::vector<int> first, rest;
  bool used[3] = {false, false, false};
  for (int b : freeb) { if (first.size() < 3 && !used[b % 3]) { used[b % 3] = true; first.push_back(b); } else rest.push_back(b); }
  first.insert(first.end(), rest.begin(), rest.end());
  return first;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const double ghz = clk_khz * 1e-6;
  std::mt19937 rng(5);
  std::normal_distribution<double> nd(0.0, 1.0);
  // NG gates; gates 2p and 2p+1 act on disjoint unit bits
  Params p;
  memset(&p, 0, sizeof(p));
  std::vector<std::vector<int>> tb(NG);
  std::vector<std::vector<cd>> U(NG, std::vector<cd>(16));
  for (int pr = 0; pr < NG / 2; ++pr) {
    std::vector<int> bits(12);
    for (int i = 0; i < 12; ++i) bits[i] = i;
    std::shuffle(bits.begin(), bits.end(), rng);
    tb[2 * pr] = {std::min(bits[0], bits[1]), std::max(bits[0], bits[1])};
    tb[2 * pr + 1] = {std::min(bits[2], bits[3]), std::max(bits[2], bits[3])};
  }
  std::vector<unsigned short> thr_single(NG * 256), thr_pair(NG / 2 * 256);
  for (int g = 0; g < NG; ++g) {
    for (auto& u : U[g]) u = cd(nd(rng), nd(rng)) / std::sqrt(8.0);
    for (int e = 0; e < 16; ++e) { p.u[g][2 * e] = float(U[g][e].real()); p.u[g][2 * e + 1] = float(U[g][e].imag()); }
    for (int m = 0; m < 4; ++m) p.xo[g][m] = (unsigned short)h_swz(scatter(m, tb[g]));
    std::vector<int> freeb;
    for (int b = 0; b < 12; ++b) if (b != tb[g][0] && b != tb[g][1]) freeb.push_back(b);
    freeb = lane_order(freeb);                       // 10 free bits: 8 thread bits + 2 iteration bits
    for (int tid = 0; tid < 256; ++tid) thr_single[g * 256 + tid] = (unsigned short)h_swz(scatter(tid, freeb));
    for (int it = 0; it < 4; ++it) p.iter[g][it] = (unsigned short)h_swz(scatter(uint32_t(it) << 8, freeb));
  }
  for (int pr = 0; pr < NG / 2; ++pr) {
    std::vector<int> freeb;
    for (int b = 0; b < 12; ++b)
      if (b != tb[2 * pr][0] && b != tb[2 * pr][1] && b != tb[2 * pr + 1][0] && b != tb[2 * pr + 1][1]) freeb.push_back(b);
    freeb = lane_order(freeb);                       // 8 free bits = one work item per thread
    for (int tid = 0; tid < 256; ++tid) thr_pair[pr * 256 + tid] = (unsigned short)h_swz(scatter(tid, freeb));
  }
  // reference
  std::vector<cd> amp(8192);
  std::vector<float> h(16384);
  for (int i = 0; i < 8192; ++i) { h[2 * i] = float(nd(rng) / 64); h[2 * i + 1] = float(nd(rng) / 64); amp[i] = cd(h[2 * i], h[2 * i + 1]); }
  for (int g = 0; g < NG; ++g) {
    const int a0 = tb[g][0] + 1, a1 = tb[g][1] + 1;
    for (uint32_t base = 0; base < 8192; ++base) {
      if (base & ((1u << a0) | (1u << a1))) continue;
      cd in[4], out[4];
      for (int m = 0; m < 4; ++m) in[m] = amp[base | ((m & 1) << a0) | ((m >> 1) << a1)];
      for (int i = 0; i < 4; ++i) { out[i] = 0; for (int j = 0; j < 4; ++j) out[i] += U[g][i * 4 + j] * in[j]; }
      for (int m = 0; m < 4; ++m) amp[base | ((m & 1) << a0) | ((m >> 1) << a1)] = out[m];
    }
  }
  float4* d_tile;
  unsigned short *d_ts, *d_tp;
  CK(cudaMalloc(&d_tile, 65536));
  CK(cudaMalloc(&d_ts, thr_single.size() * 2));
  CK(cudaMalloc(&d_tp, thr_pair.size() * 2));
  CK(cudaMemcpy(d_ts, thr_single.data(), thr_single.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_tp, thr_pair.data(), thr_pair.size() * 2, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto run = [&](const char* name, auto kern, const unsigned short* tbl, int reps) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaMemcpy(d_tile, h.data(), 65536, cudaMemcpyHostToDevice));
    kern<<<1, 256, 65536>>>(d_tile, tbl, p, 1, 1);
    CK(cudaDeviceSynchronize());
    std::vector<float> o(16384);
    CK(cudaMemcpy(o.data(), d_tile, 65536, cudaMemcpyDeviceToHost));
    double err = 0;
    for (int i = 0; i < 8192; ++i) err = fmax(err, std::abs(cd(o[2 * i], o[2 * i + 1]) - amp[i]));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 65536));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    kern<<<sms * occ, 256, 65536>>>(d_tile, tbl, p, 2, 0);
    CK(cudaEventRecord(e0));
    kern<<<sms * occ, 256, 65536>>>(d_tile, tbl, p, reps, 0);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / (double(occ) * reps * NG);
    printf("{\"test\": \"%s\", \"occ\": %d, \"regs\": %d, \"local_bytes\": %zu, \"max_abs_err\": %.3e, \"us_per_tile_gate_per_sm\": %.4f, "
           "\"clk_per_tile_gate_per_sm\": %.0f, \"ms_per_matrix_n30\": %.4f}\n",
           name, occ, fa.numRegs, (size_t)fa.localSizeBytes, err, us, us * 1e3 * ghz, us * 1e-3 * (131072.0 / sms));
    fflush(stdout);
  };
  run("single occ3", k_single<3>, d_ts, 40);
  run("single occ2", k_single<2>, d_ts, 40);
  run("pair occ3", k_pair<3>, d_tp, 40);
  run("pair occ2", k_pair<2>, d_tp, 40);
  return 0;
}
