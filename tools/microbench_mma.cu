// microbench_mma.cu -- can warp-level tensor-core MMA carry the gate arithmetic of the tile kernel?
// (round-1 evidence for hq_mma.cuh; diagnostics only, prints JSON lines)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hybridq_b200/csrc \
//        -o tools/bin/microbench_mma tools/microbench_mma.cu
//
// 1. fragment-layout checks of mma.m16n8k8.tf32 and mma.m8n8k4.f64 against a host product
// 2. issue rate of both instructions (8 independent accumulators per warp, 8/16/24 warps per SM)
// 3. the gate loops of hq_mma.cuh on a 64 KiB shared-memory tile: verified against a host
//    reference, then timed without any global traffic (cycles per tile-gate per SM), k = 2..6,
//    complex64 (3xTF32, unit and amplitude granularity) and complex128 (DMMA)
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "hq_mma.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

using hq::dmma;
using hq::mma_tf32;

// ---------------------------------------------------------------------------------------------
// 1. layout checks
// ---------------------------------------------------------------------------------------------
__global__ void k_layout_tf32(const float* A, const float* B, float* D) {   // A 16x8 row-major, B 8x8 [kk][n], D 16x8
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  uint32_t a[4] = {__float_as_uint(A[g * 8 + t]), __float_as_uint(A[(g + 8) * 8 + t]),
                   __float_as_uint(A[g * 8 + t + 4]), __float_as_uint(A[(g + 8) * 8 + t + 4])};
  float d[4] = {0, 0, 0, 0};
  mma_tf32(d, a, __float_as_uint(B[t * 8 + g]), __float_as_uint(B[(t + 4) * 8 + g]));
  D[g * 8 + 2 * t] = d[0]; D[g * 8 + 2 * t + 1] = d[1];
  D[(g + 8) * 8 + 2 * t] = d[2]; D[(g + 8) * 8 + 2 * t + 1] = d[3];
}
__global__ void k_layout_f64(const double* A, const double* B, double* D) {  // A 8x4, B 4x8 [kk][n], D 8x8
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double d0 = 0, d1 = 0;
  dmma(d0, d1, A[g * 4 + t], B[t * 8 + g]);
  D[g * 8 + 2 * t] = d0; D[g * 8 + 2 * t + 1] = d1;
}

// ---------------------------------------------------------------------------------------------
// 2. issue rates
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rate_tf32(float* out, int iters) {
  float d[8][4];
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u};
#pragma unroll
  for (int c = 0; c < 8; ++c) for (int e = 0; e < 4; ++e) d[c][e] = c + e;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[0]), "r"(a[1]));
  }
  float s = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) for (int e = 0; e < 4; ++e) s += d[c][e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_rate_f64(double* out, int iters) {
  double d[8][2];
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
#pragma unroll
  for (int c = 0; c < 8; ++c) { d[c][0] = c; d[c][1] = c + 1; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(d[c][0]), "+d"(d[c][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) s += d[c][0] + d[c][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#ifdef HQ_TRY_DMMA_16816
__global__ void __launch_bounds__(256) k_rate_f64_big(double* out, int iters) {   // m16n8k16 f64 (sm_90+)
  double d[4][4];
  double a[8], b[4];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = threadIdx.x * 1e-3 + e;
#pragma unroll
  for (int e = 0; e < 4; ++e) b[e] = 1.0 + e;
#pragma unroll
  for (int c = 0; c < 4; ++c) for (int e = 0; e < 4; ++e) d[c][e] = c + e;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(d[c][0]), "+d"(d[c][1]), "+d"(d[c][2]), "+d"(d[c][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) for (int e = 0; e < 4; ++e) s += d[c][e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif

// ---------------------------------------------------------------------------------------------
// 3. gate loops on a shared-memory tile
// ---------------------------------------------------------------------------------------------
static inline uint32_t h_swz(uint32_t u) { return u ^ ((u >> 3) & 7u) ^ ((u >> 6) & 7u) ^ ((u >> 9) & 7u); }
__device__ __forceinline__ uint32_t d_swz(uint32_t u) { return u ^ ((u >> 3) & 7u) ^ ((u >> 6) & 7u) ^ ((u >> 9) & 7u); }
static inline uint32_t h_ampslot(uint32_t a) { return (h_swz(a >> 1) << 1) | (a & 1u); }

struct GateTab {
  uint16_t tbl_thread[256];
  uint16_t tbl_iter[16];
  uint16_t tbl_x[64];
  uint32_t row8;       // amplitude path: XOR offset of row g+8
  uint32_t n_iter;
  uint32_t bf_off;     // offset (in fragments of 16 bytes) of this gate's B fragments
  uint32_t pad;
};

// MODE 0: complex64 unit path, 1: complex64 amplitude path, 2: complex128
template <int KS, int MODE, int OCC, int SPLIT, int UNR>
__global__ void __launch_bounds__(256, OCC) k_gates(void* gtile, const GateTab* __restrict__ tabs, const void* __restrict__ bfr,
                                                    int n_gates, int reps, int io) {
  extern __shared__ __align__(16) unsigned char smem[];
  float4* tile4 = reinterpret_cast<float4*>(smem);
  const int tid = threadIdx.x, lane = tid & 31, t = lane & 3;
  if (io) {
    const float4* src = reinterpret_cast<const float4*>(gtile);
    for (int u = tid; u < 4096; u += 256) tile4[d_swz(u)] = src[u];
  } else {
    for (int u = tid; u < 4096; u += 256) tile4[u] = make_float4(u * 1e-4f, tid * 1e-4f, 1e-3f, 2e-3f);
  }
  __syncthreads();
  for (int r = 0; r < reps; ++r) {
    for (int gi = 0; gi < n_gates; ++gi) {
      const GateTab* g = tabs + gi;
      const uint32_t st = __ldg(&g->tbl_thread[tid]);
      const uint32_t n_iter = __ldg(&g->n_iter);
      uint32_t xo[KS];
#pragma unroll
      for (int s = 0; s < KS; ++s) xo[s] = __ldg(&g->tbl_x[t + 4 * s]);
      if (MODE == 2) {
        const double2* bf = reinterpret_cast<const double2*>(bfr) + __ldg(&g->bf_off) + lane;
        double2 breg[KS <= 2 ? KS * KS : 1];
        if (KS <= 2) {
#pragma unroll
          for (int e = 0; e < KS * KS; ++e) breg[e] = __ldg(&bf[e * 32]);
        }
#pragma unroll 1
        for (uint32_t it = 0; it < n_iter; it += UNR) {
          uint32_t sb[UNR];
#pragma unroll
          for (int u = 0; u < UNR; ++u) sb[u] = st ^ __ldg(&g->tbl_iter[it + u]);
          hq::dmma_iter_f64<KS, UNR, (KS <= 2)>(reinterpret_cast<double2*>(smem), sb, xo, bf, breg, &g->tbl_x[t]);
        }
      } else {
        const float4* bf = reinterpret_cast<const float4*>(bfr) + __ldg(&g->bf_off) + lane;
        float4 breg[KS <= 2 ? KS * KS : 1];
        if (KS <= 2) {
#pragma unroll
          for (int e = 0; e < KS * KS; ++e) breg[e] = __ldg(&bf[e * 32]);
        }
        const uint32_t row8 = __ldg(&g->row8);
#pragma unroll 1
        for (uint32_t it = 0; it < n_iter; it += UNR) {
          uint32_t sb[UNR];
#pragma unroll
          for (int u = 0; u < UNR; ++u) sb[u] = st ^ __ldg(&g->tbl_iter[it + u]);
          if (MODE == 0)
            hq::mma_iter_f32_unit<KS, UNR, (KS <= 2), (KS >= 16), SPLIT>(tile4, sb, xo, bf, breg, &g->tbl_x[t]);
          else
            hq::mma_iter_f32_amp<KS, UNR, (KS <= 2), (KS >= 16), SPLIT>(reinterpret_cast<float2*>(smem), sb, row8, xo, bf, breg, &g->tbl_x[t]);
        }
      }
      __syncthreads();
    }
  }
  if (io) {
    float4* dst = reinterpret_cast<float4*>(gtile);
    for (int u = tid; u < 4096; u += 256) dst[u] = tile4[d_swz(u)];
  } else if (tile4[tid].x == 123.456f) {
    reinterpret_cast<float4*>(gtile)[tid] = tile4[tid];
  }
}

// smem-only floor: the same loop with the arithmetic removed (load + store of every unit)
template <int OCC>
__global__ void __launch_bounds__(256, OCC) k_smem_floor(void* gtile, const GateTab* __restrict__ tabs, int n_gates, int reps) {
  extern __shared__ __align__(16) unsigned char smem[];
  float4* tile4 = reinterpret_cast<float4*>(smem);
  const int tid = threadIdx.x, t = tid & 3;
  for (int u = tid; u < 4096; u += 256) tile4[u] = make_float4(u * 1e-4f, tid * 1e-4f, 1e-3f, 2e-3f);
  __syncthreads();
  for (int r = 0; r < reps; ++r)
    for (int gi = 0; gi < n_gates; ++gi) {
      const GateTab* g = tabs + gi;
      const uint32_t st = __ldg(&g->tbl_thread[tid]);
      const uint32_t n_iter = __ldg(&g->n_iter);
      const uint32_t xo = __ldg(&g->tbl_x[t]);
#pragma unroll 1
      for (uint32_t it = 0; it < n_iter; ++it) {
        const uint32_t sb = st ^ __ldg(&g->tbl_iter[it]) ^ xo;
        float4 v = tile4[sb];
        v.x += 1.f; v.y = v.y * 1.0001f; v.z += v.w;
        tile4[sb] = v;
      }
      __syncthreads();
    }
  if (tile4[tid].x == 123.456f) reinterpret_cast<float4*>(gtile)[tid] = tile4[tid];
}

// ---- host side -------------------------------------------------------------------------------
static uint32_t scatter(uint32_t w, const std::vector<int>& pos, int from) {   // bit (i) of w -> pos[from + i]
  uint32_t u = 0;
  for (size_t i = 0; from + i < pos.size(); ++i) u |= ((w >> i) & 1u) << pos[from + i];
  return u;
}
static float h_rna(float x) {
  uint32_t b; memcpy(&b, &x, 4);
  b = (b + 0x1000u) & 0xffffe000u;
  float r; memcpy(&r, &b, 4);
  return r;
}
typedef std::complex<double> cd;

struct Gate {
  int k;
  std::vector<int> tpos;   // ascending amplitude-level target bits inside the tile
  std::vector<cd> U;       // row-major 2^k x 2^k, matrix bit i <-> tpos[i]
};

// mode 0 / 2: unit granularity over 12 unit bits (targets tpos - V); mode 1: amplitude granularity, 13 bits
static GateTab build_tab(const Gate& g, int mode) {
  GateTab tab; memset(&tab, 0, sizeof(tab));
  const int V = mode == 2 ? 0 : 1;
  const int bits = mode == 1 ? 13 : 12;
  std::vector<int> tb;
  for (int p : g.tpos) tb.push_back(mode == 1 ? p : p - V);
  std::vector<int> nb;
  for (int b = 0; b < bits; ++b) { bool is_t = false; for (int p : tb) is_t |= (p == b); if (!is_t) nb.push_back(b); }
  {  // lane bits: the lowest row bits must land on unit bits whose residues mod 3 differ from those of the
     // lane-t target bits and from each other (conflict-free LDS.128 quarter-warps / LDS.64 half-warps)
    std::vector<int> avoid;
    int want = 1;
    if (mode == 1) { avoid.push_back((tb[1] - 1) % 3); want = 2; }
    else { avoid.push_back(tb[0] % 3); avoid.push_back(tb[1] % 3); }
    for (int c = 0; c < want; ++c)
      for (size_t i = c; i < nb.size(); ++i) {
        const int r = (mode == 1 ? nb[i] - 1 : nb[i]) % 3;
        if (mode == 1 && nb[i] == 0) continue;
        bool bad = false;
        for (int a : avoid) bad |= (a == r);
        if (!bad) { std::swap(nb[c], nb[i]); avoid.push_back(r); break; }
      }
  }
  auto slot = [&](uint32_t x) { return mode == 1 ? h_ampslot(x) : h_swz(x); };
  const int rowbits = mode == 1 ? 4 : 3;          // rows per warp-iteration: 16 (amp path) or 8
  for (int tid = 0; tid < 256; ++tid) {
    const int lane = tid & 31, warp = tid >> 5, gg = lane >> 2;
    tab.tbl_thread[tid] = (uint16_t)slot(scatter(uint32_t(gg) | (uint32_t(warp) << rowbits), nb, 0));
  }
  const int n_iter_log2 = int(nb.size()) - rowbits - 3;
  tab.n_iter = 1u << n_iter_log2;
  for (uint32_t it = 0; it < tab.n_iter; ++it) tab.tbl_iter[it] = (uint16_t)slot(scatter(it, nb, rowbits + 3));
  for (uint32_t m = 0; m < (1u << g.k); ++m) tab.tbl_x[m] = (uint16_t)slot(scatter(m, tb, 0));
  tab.row8 = mode == 1 ? slot(1u << nb[3]) : 0;
  return tab;
}

static void build_bfrag_f32(const Gate& g, std::vector<float>& out) {
  const int dim = 1 << g.k, KS = dim / 4;
  for (int s = 0; s < KS; ++s)
    for (int j = 0; j < KS; ++j)
      for (int lane = 0; lane < 32; ++lane) {
        const int gg = lane >> 2, t = lane & 3, m = 4 * s + t, mo = 4 * j + (gg >> 1), odd = gg & 1;
        const cd u = g.U[mo * dim + m];
        const float b0 = float(odd ? u.imag() : u.real()), b1 = float(odd ? u.real() : -u.imag());
        const float h0 = h_rna(b0), h1 = h_rna(b1);
        out.push_back(h0); out.push_back(h1); out.push_back(h_rna(b0 - h0)); out.push_back(h_rna(b1 - h1));
      }
}
static void build_bfrag_f64(const Gate& g, std::vector<double>& out) {
  const int dim = 1 << g.k, KS = dim / 4;
  for (int s = 0; s < KS; ++s)
    for (int j = 0; j < KS; ++j)
      for (int lane = 0; lane < 32; ++lane) {
        const int gg = lane >> 2, t = lane & 3, m = 4 * s + t, mo = 4 * j + (gg >> 1), odd = gg & 1;
        const cd u = g.U[mo * dim + m];
        out.push_back(odd ? u.imag() : u.real());
        out.push_back(odd ? u.real() : -u.imag());
      }
}

static void host_apply(std::vector<cd>& amp, const Gate& g) {
  const int dim = 1 << g.k;
  const size_t n = amp.size();
  uint32_t mask = 0;
  for (int p : g.tpos) mask |= 1u << p;
  std::vector<cd> in(dim);
  for (size_t base = 0; base < n; ++base) {
    if (base & mask) continue;
    for (int m = 0; m < dim; ++m) in[m] = amp[base | scatter(m, g.tpos, 0)];
    for (int i = 0; i < dim; ++i) {
      cd a = 0;
      for (int m = 0; m < dim; ++m) a += g.U[i * dim + m] * in[m];
      amp[base | scatter(i, g.tpos, 0)] = a;
    }
  }
}

static Gate random_gate(int k, const std::vector<int>& tpos, std::mt19937& rng) {
  Gate g; g.k = k; g.tpos = tpos;
  std::normal_distribution<double> nd(0.0, 1.0);
  const int dim = 1 << k;
  g.U.resize(dim * dim);
  for (auto& u : g.U) u = cd(nd(rng), nd(rng)) / std::sqrt(2.0 * dim);    // rows of norm ~1
  return g;
}

template <int KS, int MODE, int OCC, int SPLIT = 0, int UNR = 1>
static void run_gates(const char* name, const std::vector<Gate>& gates, int reps, double clock_ghz, int sms) {
  const int n_amp = MODE == 2 ? 4096 : 8192;
  std::mt19937 rng(7);
  std::normal_distribution<double> nd(0.0, 1.0);
  std::vector<cd> amp(n_amp);
  for (auto& a : amp) a = cd(nd(rng), nd(rng)) * (1.0 / 64);
  std::vector<GateTab> tabs;
  std::vector<float> bf32;
  std::vector<double> bf64;
  for (const Gate& g : gates) {
    GateTab tab = build_tab(g, MODE);
    tab.bf_off = uint32_t(MODE == 2 ? bf64.size() / 2 : bf32.size() / 4);
    if (MODE == 2) build_bfrag_f64(g, bf64); else build_bfrag_f32(g, bf32);
    tabs.push_back(tab);
  }
  void *d_tile, *d_tabs, *d_bf;
  CK(cudaMalloc(&d_tile, 65536));
  CK(cudaMalloc(&d_tabs, tabs.size() * sizeof(GateTab)));
  const size_t bf_bytes = MODE == 2 ? bf64.size() * 8 : bf32.size() * 4;
  CK(cudaMalloc(&d_bf, bf_bytes));
  CK(cudaMemcpy(d_tabs, tabs.data(), tabs.size() * sizeof(GateTab), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_bf, MODE == 2 ? (void*)bf64.data() : (void*)bf32.data(), bf_bytes, cudaMemcpyHostToDevice));
  std::vector<float> h32(2 * n_amp);
  std::vector<double> h64(2 * n_amp);
  for (int i = 0; i < n_amp; ++i) { h32[2 * i] = float(amp[i].real()); h32[2 * i + 1] = float(amp[i].imag());
                                    h64[2 * i] = amp[i].real(); h64[2 * i + 1] = amp[i].imag(); }
  if (MODE != 2) for (int i = 0; i < n_amp; ++i) amp[i] = cd(h32[2 * i], h32[2 * i + 1]);
  CK(cudaMemcpy(d_tile, MODE == 2 ? (void*)h64.data() : (void*)h32.data(), 65536, cudaMemcpyHostToDevice));
  auto kern = k_gates<KS, MODE, OCC, SPLIT, UNR>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  // correctness: one CTA, one repetition
  kern<<<1, 256, 65536>>>(d_tile, (const GateTab*)d_tabs, d_bf, int(gates.size()), 1, 1);
  CK(cudaDeviceSynchronize());
  for (const Gate& g : gates) host_apply(amp, g);
  double err = 0, mag = 0;
  if (MODE == 2) {
    CK(cudaMemcpy(h64.data(), d_tile, 65536, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n_amp; ++i) { err = fmax(err, std::abs(cd(h64[2 * i], h64[2 * i + 1]) - amp[i])); mag = fmax(mag, std::abs(amp[i])); }
  } else {
    CK(cudaMemcpy(h32.data(), d_tile, 65536, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n_amp; ++i) { err = fmax(err, std::abs(cd(h32[2 * i], h32[2 * i + 1]) - amp[i])); mag = fmax(mag, std::abs(amp[i])); }
  }
  // timing: all SMs, no global traffic
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 65536));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  kern<<<sms * occ, 256, 65536>>>(d_tile, (const GateTab*)d_tabs, d_bf, int(gates.size()), 2, 0);
  CK(cudaEventRecord(e0));
  kern<<<sms * occ, 256, 65536>>>(d_tile, (const GateTab*)d_tabs, d_bf, int(gates.size()), reps, 0);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double tile_gates_per_sm = double(occ) * reps * gates.size();
  const double us_per_tg = ms * 1e3 / tile_gates_per_sm;
  printf("{\"test\": \"gate_loop\", \"name\": \"%s\", \"k\": %d, \"mode\": %d, \"occ\": %d, \"n_gates\": %zu, \"max_abs_err\": %.3e, \"max_abs\": %.3e, "
         "\"us_per_tile_gate_per_sm\": %.4f, \"clk_per_tile_gate_per_sm\": %.0f, \"ms_per_matrix_n30_c64_or_n29_c128\": %.4f}\n",
         name, gates[0].k, MODE, occ, gates.size(), err, mag, us_per_tg, us_per_tg * 1e3 * clock_ghz, us_per_tg * 1e-3 * (131072.0 / sms));
  fflush(stdout);
  cudaFree(d_tile); cudaFree(d_tabs); cudaFree(d_bf);
}

template <int OCC>
static void run_floor(const std::vector<Gate>& gates, int reps, double clock_ghz, int sms) {
  std::vector<GateTab> tabs;
  for (const Gate& g : gates) tabs.push_back(build_tab(g, 0));
  void *d_tile, *d_tabs;
  CK(cudaMalloc(&d_tile, 65536));
  CK(cudaMalloc(&d_tabs, tabs.size() * sizeof(GateTab)));
  CK(cudaMemcpy(d_tabs, tabs.data(), tabs.size() * sizeof(GateTab), cudaMemcpyHostToDevice));
  auto kern = k_smem_floor<OCC>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 65536));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  kern<<<sms * occ, 256, 65536>>>(d_tile, (const GateTab*)d_tabs, int(gates.size()), 2);
  CK(cudaEventRecord(e0));
  kern<<<sms * occ, 256, 65536>>>(d_tile, (const GateTab*)d_tabs, int(gates.size()), reps);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double us_per_tg = ms * 1e3 / (double(occ) * reps * gates.size());
  printf("{\"test\": \"smem_floor\", \"occ\": %d, \"us_per_tile_gate_per_sm\": %.4f, \"clk_per_tile_gate_per_sm\": %.0f}\n", occ, us_per_tg,
         us_per_tg * 1e3 * clock_ghz);
  cudaFree(d_tile); cudaFree(d_tabs);
}

static std::vector<Gate> gate_list(int k, int mode, int n, bool conflict, std::mt19937& rng) {
  // amplitude-level targets inside a 13-bit (c64) / 12-bit (c128) tile
  const int bits = mode == 2 ? 12 : 13;
  const int lo = mode == 0 ? 1 : 0;            // unit path: bit 0 is not a target
  const int V = mode == 2 ? 0 : 1;
  std::vector<Gate> v;
  while (int(v.size()) < n) {
    std::vector<int> tp;
    if (mode == 1) tp.push_back(0);
    while (int(tp.size()) < k) {
      const int p = lo + int(rng() % (bits - lo));
      bool dup = false;
      for (int q : tp) dup |= (q == p);
      if (!dup) tp.push_back(p);
    }
    std::sort(tp.begin(), tp.end());
    if (mode != 1) {
      // residues (mod 3) of the unit bits of the two lowest targets decide the LDS.128 conflicts
      const bool same = ((tp[0] - V) % 3) == ((tp[1] - V) % 3);
      if (same != conflict) continue;
    }
    v.push_back(random_gate(k, tp, rng));
  }
  return v;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const double ghz = clk_khz * 1e-6;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz\": %.3f}\n", prop.name, sms, ghz);

  {  // 1. layouts
    float hA[128], hB[64], hD[128];
    for (int i = 0; i < 128; ++i) hA[i] = float((i * 7) % 13 - 6);
    for (int i = 0; i < 64; ++i) hB[i] = float((i * 5) % 11 - 5);
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, sizeof(hA))); CK(cudaMalloc(&dB, sizeof(hB))); CK(cudaMalloc(&dD, sizeof(hD)));
    CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
    k_layout_tf32<<<1, 32>>>(dA, dB, dD);
    CK(cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost));
    double err = 0;
    for (int r = 0; r < 16; ++r) for (int n = 0; n < 8; ++n) {
      double s = 0; for (int kk = 0; kk < 8; ++kk) s += double(hA[r * 8 + kk]) * hB[kk * 8 + n];
      err = fmax(err, fabs(s - hD[r * 8 + n]));
    }
    printf("{\"test\": \"layout_tf32_m16n8k8\", \"max_abs_err\": %.3e, \"ok\": %s}\n", err, err == 0 ? "true" : "false");
    double gA[32], gB[32], gD[64];
    for (int i = 0; i < 32; ++i) { gA[i] = (i * 7) % 13 - 6 + 0.25; gB[i] = (i * 5) % 11 - 5 + 0.5; }
    double *eA, *eB, *eD;
    CK(cudaMalloc(&eA, sizeof(gA))); CK(cudaMalloc(&eB, sizeof(gB))); CK(cudaMalloc(&eD, sizeof(gD)));
    CK(cudaMemcpy(eA, gA, sizeof(gA), cudaMemcpyHostToDevice)); CK(cudaMemcpy(eB, gB, sizeof(gB), cudaMemcpyHostToDevice));
    k_layout_f64<<<1, 32>>>(eA, eB, eD);
    CK(cudaMemcpy(gD, eD, sizeof(gD), cudaMemcpyDeviceToHost));
    err = 0;
    for (int r = 0; r < 8; ++r) for (int n = 0; n < 8; ++n) {
      double s = 0; for (int kk = 0; kk < 4; ++kk) s += gA[r * 4 + kk] * gB[kk * 8 + n];
      err = fmax(err, fabs(s - gD[r * 8 + n]));
    }
    printf("{\"test\": \"layout_f64_m8n8k4\", \"max_abs_err\": %.3e, \"ok\": %s}\n", err, err < 1e-12 ? "true" : "false");
  }

  {  // 2. issue rates
    void* out;
    CK(cudaMalloc(&out, size_t(sms) * 4 * 256 * 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int ctas = 1; ctas <= 4; ++ctas) {
      const int iters = 20000;
      float ms;
      k_rate_tf32<<<sms * ctas, 256>>>((float*)out, 100);
      CK(cudaEventRecord(e0));
      k_rate_tf32<<<sms * ctas, 256>>>((float*)out, iters);
      CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaEventElapsedTime(&ms, e0, e1));
      double mmas = double(sms) * ctas * 8 * iters * 8;
      printf("{\"test\": \"rate_tf32_m16n8k8\", \"warps_per_sm\": %d, \"mma_per_clk_per_sm\": %.4f, \"tflops\": %.1f}\n", ctas * 8,
             mmas / (ms * 1e-3) / (ghz * 1e9) / sms, mmas * 2048 / (ms * 1e-3) / 1e12);
      k_rate_f64<<<sms * ctas, 256>>>((double*)out, 100);
      CK(cudaEventRecord(e0));
      k_rate_f64<<<sms * ctas, 256>>>((double*)out, iters / 4);
      CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaEventElapsedTime(&ms, e0, e1));
      mmas = double(sms) * ctas * 8 * (iters / 4) * 8;
      printf("{\"test\": \"rate_f64_m8n8k4\", \"warps_per_sm\": %d, \"mma_per_clk_per_sm\": %.4f, \"tflops\": %.2f}\n", ctas * 8,
             mmas / (ms * 1e-3) / (ghz * 1e9) / sms, mmas * 512 / (ms * 1e-3) / 1e12);
#ifdef HQ_TRY_DMMA_16816
      k_rate_f64_big<<<sms * ctas, 256>>>((double*)out, 100);
      CK(cudaEventRecord(e0));
      k_rate_f64_big<<<sms * ctas, 256>>>((double*)out, iters / 16);
      CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaEventElapsedTime(&ms, e0, e1));
      mmas = double(sms) * ctas * 8 * (iters / 16) * 4;
      printf("{\"test\": \"rate_f64_m16n8k16\", \"warps_per_sm\": %d, \"mma_per_clk_per_sm\": %.4f, \"tflops\": %.2f}\n", ctas * 8,
             mmas / (ms * 1e-3) / (ghz * 1e9) / sms, mmas * 4096 / (ms * 1e-3) / 1e12);
#endif
      fflush(stdout);
    }
  }

  std::mt19937 rng(11);
  {  // 3. gate loops
    auto g2 = gate_list(2, 0, 16, false, rng);
    auto g2c = gate_list(2, 0, 16, true, rng);
    auto g3 = gate_list(3, 0, 16, false, rng);
    auto g4 = gate_list(4, 0, 16, false, rng);
    run_floor<3>(g2, 40, ghz, sms);
    run_gates<1, 0, 3, 0, 1>("c64 unit k=2 unr1", g2, 40, ghz, sms);
    run_gates<1, 0, 3, 0, 2>("c64 unit k=2 unr2", g2, 40, ghz, sms);
    run_gates<1, 0, 3, 0, 4>("c64 unit k=2 unr4", g2, 40, ghz, sms);
    run_gates<1, 0, 3, 1, 4>("c64 unit k=2 unr4 split1", g2, 40, ghz, sms);
    run_gates<1, 0, 3, 0, 4>("c64 unit k=2 unr4 same-residue targets", g2c, 40, ghz, sms);
    run_gates<2, 0, 3, 0, 1>("c64 unit k=3 unr1", g3, 20, ghz, sms);
    run_gates<2, 0, 3, 0, 2>("c64 unit k=3 unr2", g3, 20, ghz, sms);
    run_gates<2, 0, 2, 0, 2>("c64 unit k=3 unr2 occ2", g3, 20, ghz, sms);
    run_gates<4, 0, 2, 0, 1>("c64 unit k=4 unr1 occ2", g4, 10, ghz, sms);
    run_gates<4, 0, 2, 0, 2>("c64 unit k=4 unr2 occ2", g4, 10, ghz, sms);
    run_gates<8, 0, 2, 0, 1>("c64 unit k=5", gate_list(5, 0, 8, false, rng), 6, ghz, sms);
    run_gates<16, 0, 2, 0, 1>("c64 unit k=6", gate_list(6, 0, 8, false, rng), 3, ghz, sms);
    run_gates<1, 1, 3, 0, 1>("c64 amp k=2 unr1", gate_list(2, 1, 16, false, rng), 40, ghz, sms);
    run_gates<1, 1, 3, 0, 4>("c64 amp k=2 unr4", gate_list(2, 1, 16, false, rng), 40, ghz, sms);
    run_gates<2, 1, 3, 0, 2>("c64 amp k=3 unr2", gate_list(3, 1, 16, false, rng), 20, ghz, sms);
    run_gates<4, 1, 2, 0, 1>("c64 amp k=4 occ2", gate_list(4, 1, 16, false, rng), 10, ghz, sms);
    run_gates<8, 1, 2, 0, 1>("c64 amp k=5", gate_list(5, 1, 8, false, rng), 6, ghz, sms);
    auto d2 = gate_list(2, 2, 16, false, rng);
    auto d3 = gate_list(3, 2, 16, false, rng);
    auto d4 = gate_list(4, 2, 16, false, rng);
    run_gates<1, 2, 3, 0, 1>("c128 k=2 unr1", d2, 40, ghz, sms);
    run_gates<1, 2, 3, 0, 2>("c128 k=2 unr2", d2, 40, ghz, sms);
    run_gates<1, 2, 3, 0, 4>("c128 k=2 unr4", d2, 40, ghz, sms);
    run_gates<1, 2, 3, 0, 4>("c128 k=2 unr4 same-residue targets", gate_list(2, 2, 16, true, rng), 40, ghz, sms);
    run_gates<2, 2, 3, 0, 1>("c128 k=3 unr1", d3, 20, ghz, sms);
    run_gates<2, 2, 3, 0, 2>("c128 k=3 unr2", d3, 20, ghz, sms);
    run_gates<4, 2, 3, 0, 1>("c128 k=4 unr1 occ3", d4, 10, ghz, sms);
    run_gates<4, 2, 2, 0, 1>("c128 k=4 unr1 occ2", d4, 10, ghz, sms);
    run_gates<4, 2, 2, 0, 2>("c128 k=4 unr2 occ2", d4, 10, ghz, sms);
    run_gates<8, 2, 2, 0, 1>("c128 k=5", gate_list(5, 2, 8, false, rng), 4, ghz, sms);
    run_gates<16, 2, 2, 0, 1>("c128 k=6", gate_list(6, 2, 8, false, rng), 2, ghz, sms);
  }
  return 0;
}
