// hq_kernels.cu -- hand-written sm_100a kernels of the state-vector evolution core.
//
//   hq_tile_kernel<T>      persistent tile kernel: one launch = one pass over the state,
//                          any number of gates with k <= HQ_MAX_K on any target bits
//                          (replaces /root/reference/include/U.h:28-202), optional in-tile
//                          index-bit permutation on write-back (replaces swap.h:47-95 for the
//                          device-resident state).
//   hq_direct_kernel<...>  single-gate kernel without shared memory for k <= 3: each thread
//                          owns whole groups, U sits in the constant bank (kernel parameters).
//   hq_pack / hq_unpack    split planes <-> interleaved complex (python_U.cpp:114-123).
//   hq_bitperm_oop         out-of-place low-bit permutation of a real array (swap.h:28-33),
//                          used by the host-pointer swap_* ABI.
//   hq_init_* / hq_norm2 / hq_vdot / hq_scale   state preparation and reductions.
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

#include "hq_kernels.h"
#include "hq_mma.cuh"
#include "hq_tile.cuh"

#ifndef HQ_K0F_BLOCKS
#define HQ_K0F_BLOCKS 3   // resident CTAs per SM the k <= 2 complex64 tile kernel is compiled for
#endif
#ifndef HQ_MMA_BREG_KS
#define HQ_MMA_BREG_KS 4  // gates with 2^k / 4 <= this keep their B fragments in registers for the whole gate
#endif
#ifndef HQ_K1D_BLOCKS
#define HQ_K1D_BLOCKS 3   // resident CTAs per SM of the k <= 3 complex128 tile kernel (measured: 134.6 vs 142.4 ms/step)
#endif

namespace hq {

// ---------------------------------------------------------------------------------------
// small PTX helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}
// streaming 16-byte store (the tile is not re-read before it has left L2 anyway)
template <typename Unit>
__device__ __forceinline__ void st_stream(Unit* p, const Unit& v) {
  __stcs(p, v);
}

// ---------------------------------------------------------------------------------------
// the tile kernel
//
// Persistent: gridDim.x = SMs x resident CTAs, every CTA walks tiles blockIdx.x,
// blockIdx.x + gridDim.x, ...  With NBUF = 2 the fill of the CTA's next tile is issued
// (cp.async, no registers involved) before the gates of the current tile run, so every CTA
// keeps one whole tile of loads in flight while it computes and drains.
// KCLASS: 0 = passes whose gates all have k <= 2, 1 = k <= 3, 2 = k <= 4, 3 = anything (adds
// the two-phase path); the narrower classes need far fewer registers.
// ---------------------------------------------------------------------------------------
// Fill: local unit c = tid + (i << 8) comes from global unit base + off_t + iter_off[i] and
// goes to slot swz_t ^ iter_swz[i]; iter_* sit in the constant bank (kernel parameter).
template <typename T>
__device__ __forceinline__ void tile_fill(typename Traits<T>::Unit* tile,
                                          const typename Traits<T>::Unit* __restrict__ src,
                                          const HqPassHeader& ph, uint32_t swz_t, int npt) {
#pragma unroll 4
  for (int i = 0; i < npt; ++i) cp_async16(&tile[swz_t ^ ph.iter_swz[i]], src + ph.iter_off[i]);
}

// ---------------------------------------------------------------------------------------
// tensor-core gates (HQ_GATE_MMA): table-driven addressing around the loops of hq_mma.cuh.
// A warp owns row sets it = 0 .. mma_n_iter-1 (UNR of them in flight at a time); warps beyond
// mma_warps have no rows (small tiles only) and skip the gate as a whole, so every mma.sync
// is executed by full warps.
// ---------------------------------------------------------------------------------------
template <int KS, int UNR>
__device__ __forceinline__ void gate_mma_rows_f32(float4* tile, const HqGateDesc* __restrict__ g, uint32_t st,
                                                  uint32_t n_iter, bool amp, uint32_t row8, const uint32_t (&xo)[KS],
                                                  const float4* __restrict__ bf, const float4* breg,
                                                  const uint16_t* __restrict__ xtab) {
#pragma unroll 1
  for (uint32_t it = 0; it < n_iter; it += UNR) {
    uint32_t sb[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) sb[u] = st ^ __ldg(&g->tbl_iter[it + u]);   // measured: the load beats iter_offset() here
    if (amp)
      mma_iter_f32_amp<KS, UNR, (KS <= HQ_MMA_BREG_KS), (KS >= 16)>(reinterpret_cast<float2*>(tile), sb, row8, xo, bf, breg, xtab);
    else
      mma_iter_f32_unit<KS, UNR, (KS <= HQ_MMA_BREG_KS), (KS >= 16)>(tile, sb, xo, bf, breg, xtab);
  }
}

template <int KS, int UNR>
__device__ __forceinline__ void gate_mma(float4* tile, const HqGateDesc* __restrict__ g,
                                         const unsigned char* __restrict__ prog, int tid) {
  const int lane = tid & 31, t = lane & 3;
  if (uint32_t(tid >> 5) >= __ldg(&g->mma_warps)) return;
  const uint32_t st = __ldg(&g->tbl_thread[tid]);
  const uint32_t n_iter = __ldg(&g->mma_n_iter);
  const bool amp = __ldg(&g->mma_amp) != 0;
  const uint32_t row8 = __ldg(&g->mma_row8);
  uint32_t xo[KS];
#pragma unroll
  for (int s = 0; s < KS; ++s) xo[s] = __ldg(&g->tbl_x[t + 4 * s]);
  const float4* bf = reinterpret_cast<const float4*>(prog + __ldg(&g->mat_off)) + lane;
  float4 breg[KS <= HQ_MMA_BREG_KS ? KS * KS : 1];
  if (KS <= HQ_MMA_BREG_KS) {
#pragma unroll
    for (int e = 0; e < KS * KS; ++e) breg[e] = __ldg(&bf[e * 32]);
  }
  const uint16_t* xtab = &g->tbl_x[t];
  if (UNR > 1 && n_iter >= uint32_t(UNR))
    gate_mma_rows_f32<KS, UNR>(tile, g, st, n_iter, amp, row8, xo, bf, breg, xtab);
  else
    gate_mma_rows_f32<KS, 1>(tile, g, st, n_iter, amp, row8, xo, bf, breg, xtab);
}

template <int KS, int UNR>
__device__ __forceinline__ void gate_mma_rows_f64(double2* tile, const HqGateDesc* __restrict__ g, uint32_t st,
                                                  uint32_t n_iter, const uint32_t (&xo)[KS],
                                                  const double2* __restrict__ bf, const double2* breg,
                                                  const uint16_t* __restrict__ xtab) {
#pragma unroll 1
  for (uint32_t it = 0; it < n_iter; it += UNR) {
    uint32_t sb[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) sb[u] = st ^ __ldg(&g->tbl_iter[it + u]);   // measured: the load beats iter_offset() here
    dmma_iter_f64<KS, UNR, (KS <= HQ_MMA_BREG_KS)>(tile, sb, xo, bf, breg, xtab);
  }
}

template <int KS, int UNR>
__device__ __forceinline__ void gate_mma(double2* tile, const HqGateDesc* __restrict__ g,
                                         const unsigned char* __restrict__ prog, int tid) {
  const int lane = tid & 31, t = lane & 3;
  if (uint32_t(tid >> 5) >= __ldg(&g->mma_warps)) return;
  const uint32_t st = __ldg(&g->tbl_thread[tid]);
  const uint32_t n_iter = __ldg(&g->mma_n_iter);
  uint32_t xo[KS];
#pragma unroll
  for (int s = 0; s < KS; ++s) xo[s] = __ldg(&g->tbl_x[t + 4 * s]);
  const double2* bf = reinterpret_cast<const double2*>(prog + __ldg(&g->mat_off)) + lane;
  double2 breg[KS <= HQ_MMA_BREG_KS ? KS * KS : 1];
  if (KS <= HQ_MMA_BREG_KS) {
#pragma unroll
    for (int e = 0; e < KS * KS; ++e) breg[e] = __ldg(&bf[e * 32]);
  }
  const uint16_t* xtab = &g->tbl_x[t];
  if (UNR > 1 && n_iter >= uint32_t(UNR))
    gate_mma_rows_f64<KS, UNR>(tile, g, st, n_iter, xo, bf, breg, xtab);
  else
    gate_mma_rows_f64<KS, 1>(tile, g, st, n_iter, xo, bf, breg, xtab);
}

template <typename Unit> struct IsF64Unit { static const bool value = false; };
template <> struct IsF64Unit<double2> { static const bool value = true; };

// MMAK: largest k the kernel class can meet (2, 3, 4 or HQ_MMA_MAX_K)
template <int MMAK, typename Unit>
__device__ __forceinline__ void gate_mma_dispatch(Unit* tile, const HqGateDesc* g, uint32_t k,
                                                  const unsigned char* prog, int tid) {
  switch (k) {
    case 2: gate_mma<1, 2>(tile, g, prog, tid); break;
    case 3: if (MMAK >= 3) gate_mma<2, IsF64Unit<Unit>::value ? 1 : 2>(tile, g, prog, tid); break;
    case 4: if (MMAK >= 4) gate_mma<4, 1>(tile, g, prog, tid); break;
    case 5: if (MMAK >= 5) gate_mma<8, 1>(tile, g, prog, tid); break;
    case 6: if (MMAK >= 6) gate_mma<16, 1>(tile, g, prog, tid); break;
    default: break;
  }
}

template <int MAXK>
__device__ __forceinline__ void rowpair_dispatch(double2* tile, const HqGateDesc* g, uint32_t k,
                                                 const unsigned char* prog, uint32_t mat_off, int Tu, int tid) {
  const double2* U = reinterpret_cast<const double2*>(prog + mat_off);
  if (k == 2) gate_rowpair_f64<2>(tile, g, U, Tu, tid);
  else if (MAXK >= 3) gate_rowpair_f64<3>(tile, g, U, Tu, tid);
}
template <int MAXK>
__device__ __forceinline__ void rowpair_dispatch(float4*, const HqGateDesc*, uint32_t, const unsigned char*, uint32_t,
                                                 int, int) {}

// One out-of-line copy per kernel of everything but the two-phase path: the register paths, the
// complex128 row-pair scheme and the tensor-core path (it is called from the unrolled fast-slot
// sequence as well as from the gate loop).
template <int MAXK, int MMAK, typename Unit>
__device__ __noinline__ void gate_small_generic(Unit* tile, const HqGateDesc* g, uint32_t k, uint32_t kind,
                                                const unsigned char* prog, uint32_t mat_off, int Tu, int tid) {
  if (kind == HQ_GATE_MMA) {
    gate_mma_dispatch<MMAK>(tile, g, k, prog, tid);
  } else if (IsF64Unit<Unit>::value && kind == HQ_GATE_ROWPAIR) {
    rowpair_dispatch<MAXK>(tile, g, k, prog, mat_off, Tu, tid);
  } else {
    const bool low = !IsF64Unit<Unit>::value && __ldg(&g->tpos[0]) == 0;
    gate_small_dispatch<MAXK>(tile, g, k, low, prog, mat_off, Tu, tid);
  }
}

template <int S, int MAXK>
__device__ __forceinline__ void fast_slot(float4* tile, const HqGateDesc* gates, const HqPassHeader& ph,
                                          const unsigned char* prog, uint32_t n_gates, int Tu, int tid) {
  if (S < int(n_gates)) {
    const HqGateDesc* g = gates + S;
    if ((ph.fast_mask >> S) & 1u) {
      gate_fast_f32_k2<S>(tile, g, ph, Tu, tid);
    } else {
      gate_small_generic<MAXK, MAXK>(tile, g, __ldg(&g->k), __ldg(&g->kind), prog, __ldg(&g->mat_off), Tu, tid);
    }
    __syncthreads();
  }
}

template <typename T, int MAXK>
__device__ __forceinline__ void fast_slots(float4* tile, const HqGateDesc* gates, const HqPassHeader& ph,
                                           const unsigned char* prog, uint32_t n_gates, int Tu, int tid) {
  fast_slot<0, MAXK>(tile, gates, ph, prog, n_gates, Tu, tid);
  fast_slot<1, MAXK>(tile, gates, ph, prog, n_gates, Tu, tid);
  fast_slot<2, MAXK>(tile, gates, ph, prog, n_gates, Tu, tid);
  fast_slot<3, MAXK>(tile, gates, ph, prog, n_gates, Tu, tid);
  fast_slot<4, MAXK>(tile, gates, ph, prog, n_gates, Tu, tid);
  fast_slot<5, MAXK>(tile, gates, ph, prog, n_gates, Tu, tid);
  fast_slot<6, MAXK>(tile, gates, ph, prog, n_gates, Tu, tid);
  fast_slot<7, MAXK>(tile, gates, ph, prog, n_gates, Tu, tid);
}
template <typename T, int MAXK>
__device__ __forceinline__ void fast_slots(double2*, const HqGateDesc*, const HqPassHeader&, const unsigned char*,
                                           uint32_t, int, int) {}

template <typename T, int KCLASS, int NBUF>
__global__ void __launch_bounds__(HQ_THREADS, ((KCLASS == 0 && Traits<T>::V == 1) ? HQ_K0F_BLOCKS
                                : ((KCLASS == 1 && Traits<T>::V == 0) ? HQ_K1D_BLOCKS
                                   : ((KCLASS == 0 || (KCLASS == 1 && Traits<T>::V == 1)) ? 3 : 2))))
hq_tile_kernel(typename Traits<T>::Unit* __restrict__ state, const unsigned char* __restrict__ prog,
               const __grid_constant__ HqPassHeader ph, const unsigned long long n_tiles) {
  typedef typename Traits<T>::Unit Unit;
  typedef typename Traits<T>::Cplx Cplx;
  const int V = Traits<T>::V;
  const int MAXK = KCLASS == 0 ? 2 : (KCLASS == 1 ? 3 : 4);
  const int MMAK = KCLASS == 3 ? HQ_MMA_MAX_K : MAXK;
  extern __shared__ __align__(16) unsigned char smem[];

  const int tid = threadIdx.x;
  const int Tbits = int(ph.tile_bits);
  const int h = int(ph.n_high);
  const int Tu = Tbits - V;
  const int Lu = Tbits - h - V;
  const uint32_t n_units = 1u << Tu;
  const uint32_t n_gates = ph.n_gates;
  const int npt = Tu > HQ_THREADS_LOG2 ? (1 << (Tu - HQ_THREADS_LOG2)) : (uint32_t(tid) < n_units ? 1 : 0);

  Unit* bufs = reinterpret_cast<Unit*>(smem);
  const HqGateDesc* gates = reinterpret_cast<const HqGateDesc*>(prog + ph.gates_off);

  // per-thread constants of the fill/drain addressing
  const uint64_t off_t = unit_offset(uint32_t(tid), Lu, V, ph.high_pos, h);
  const uint32_t swz_t = swz(uint32_t(tid));

  unsigned long long t = blockIdx.x;
  int cur = 0;
  if (NBUF == 2 && t < n_tiles) {
    tile_fill<T>(bufs, state + (tile_base(t, Tbits, h, ph.high_pos) >> V) + off_t, ph, swz_t, npt);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  for (; t < n_tiles; t += gridDim.x) {
    Unit* gptr = state + (tile_base(t, Tbits, h, ph.high_pos) >> V) + off_t;
    Unit* tile = bufs + (size_t(cur) << Tu);
    if (NBUF == 2) {
      const unsigned long long tn = t + gridDim.x;
      if (tn < n_tiles) {
        tile_fill<T>(bufs + (size_t(cur ^ 1) << Tu), state + (tile_base(tn, Tbits, h, ph.high_pos) >> V) + off_t,
                     ph, swz_t, npt);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      }
    } else {
      tile_fill<T>(tile, gptr, ph, swz_t, npt);
      cp_async_wait_all();
    }
    __syncthreads();

    uint32_t gi0 = 0;
    if (V == 1 && KCLASS <= 1 && ph.fast_mask) {
      // unrolled slots: constant-bank matrices for the k = 2 gates, generic code for the others
      gi0 = n_gates < HQ_FAST_SLOTS ? n_gates : HQ_FAST_SLOTS;
      fast_slots<T, MAXK>(tile, gates, ph, prog, n_gates, Tu, tid);
    }
    for (uint32_t gi = gi0; gi < n_gates; ++gi) {
      const HqGateDesc* g = gates + gi;
      const uint32_t k = __ldg(&g->k);
      const uint32_t mat_off = __ldg(&g->mat_off);
      const uint32_t kind = __ldg(&g->kind);
      if (KCLASS < 3 || kind != HQ_GATE_BIG) {
        gate_small_generic<MAXK, MMAK>(tile, g, k, kind, prog, mat_off, Tu, tid);
      } else {
        const Cplx* Ut = reinterpret_cast<const Cplx*>(prog + mat_off);
        const int rounds = big_rounds(Tbits, int(k));
        for (int r = 0; r < rounds; ++r) {
          BigAcc<T> acc;
          gate_big_phaseA<T>(reinterpret_cast<const Cplx*>(tile), *g, Ut, Tbits, tid, r, acc);
          __syncthreads();
          gate_big_phaseB<T>(reinterpret_cast<Cplx*>(tile), *g, acc);
        }
      }
      __syncthreads();
    }

    // drain
    if (!ph.has_perm) {
#pragma unroll 4
      for (int i = 0; i < npt; ++i) st_stream(gptr + ph.iter_off[i], tile[swz_t ^ ph.iter_swz[i]]);
    } else {
      const Cplx* amps = reinterpret_cast<const Cplx*>(tile);
      for (int i = 0; i < npt; ++i) {
        const uint32_t c = uint32_t(tid) + (uint32_t(i) << HQ_THREADS_LOG2);
        Cplx o[1 << V];
#pragma unroll
        for (uint32_t e = 0; e < (1u << V); ++e)
          o[e] = amps[amp_slot<T>(perm_src((c << V) | e, ph.perm, Tbits))];
        st_stream(gptr + ph.iter_off[i], make_unit(o));
      }
    }
    __syncthreads();
    cur ^= (NBUF == 2);
  }
}

static int g_tune_nbuf = 0;          // 0 = auto: double-buffer when it costs no resident CTA
static int g_tune_ctas_per_sm = 0;
void set_tuning(int nbuf, int ctas_per_sm) {
  if (nbuf >= 0 && nbuf <= 2) g_tune_nbuf = nbuf;
  if (ctas_per_sm >= 0) g_tune_ctas_per_sm = ctas_per_sm;
}

size_t tile_pass_smem_bytes(const HqPassHeader& ph, int dtype, int nbuf) {
  const int V = dtype == HQ_DTYPE_C64 ? 1 : 0;
  const int Tu = int(ph.tile_bits) - V;
  return size_t(16 * nbuf) << Tu;
}

static DeviceInfo g_info[64];
static bool g_info_ok[64];

int device_info(DeviceInfo* out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return int(e);
  if (dev < 0 || dev >= 64) return int(cudaErrorInvalidDevice);
  if (!g_info_ok[dev]) {
    e = cudaDeviceGetAttribute(&g_info[dev].sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return int(e);
    e = cudaDeviceGetAttribute(&g_info[dev].max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return int(e);
    g_info_ok[dev] = true;
  }
  *out = g_info[dev];
  return 0;
}

// resident CTAs per SM of one kernel variant at a given dynamic shared-memory size (cached)
template <typename T, int KCLASS, int NBUF>
static int variant_occupancy(size_t smem, const DeviceInfo& di, int dev, int* per_sm_out) {
  static bool attr_set[64];
  static int cache[64][HQ_MAX_UNIT_BITS + 2];
  auto kern = hq_tile_kernel<T, KCLASS, NBUF>;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin);
    if (e != cudaSuccess) return int(e);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return int(e);
    for (int i = 0; i < HQ_MAX_UNIT_BITS + 2; ++i) cache[dev][i] = -1;
    attr_set[dev] = true;
  }
  int slot = 0;
  while ((size_t(16 * NBUF) << slot) < smem && slot < HQ_MAX_UNIT_BITS + 1) ++slot;
  if (cache[dev][slot] < 0) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, HQ_THREADS, smem);
    if (e != cudaSuccess) return int(e);
    cache[dev][slot] = per_sm;
  }
  *per_sm_out = cache[dev][slot];
  return 0;
}

template <typename T, int KCLASS, int NBUF>
static int launch_tile_variant(void* state, unsigned n_qubits, const unsigned char* prog, const HqPassHeader& ph,
                               cudaStream_t stream, int grid_override, size_t smem, const DeviceInfo& di, int per_sm) {
  const unsigned long long n_tiles = 1ull << (n_qubits - ph.tile_bits);
  if (per_sm < 1) return int(cudaErrorLaunchOutOfResources);
  if (g_tune_ctas_per_sm > 0 && g_tune_ctas_per_sm < per_sm) per_sm = g_tune_ctas_per_sm;
  unsigned long long grid = (unsigned long long)di.sm_count * (unsigned long long)per_sm;
  if (grid_override > 0) grid = (unsigned long long)grid_override;
  if (grid > n_tiles) grid = n_tiles;
  hq_tile_kernel<T, KCLASS, NBUF><<<unsigned(grid), HQ_THREADS, smem, stream>>>(
      reinterpret_cast<typename Traits<T>::Unit*>(state), prog, ph, n_tiles);
  return int(cudaGetLastError());
}

template <typename T, int KCLASS>
static int launch_tile_class(void* state, unsigned n_qubits, const unsigned char* prog, const HqPassHeader& ph,
                             cudaStream_t stream, int grid_override, const DeviceInfo& di, int dev) {
  const int dtype = Traits<T>::V == 1 ? HQ_DTYPE_C64 : HQ_DTYPE_C128;
  const size_t smem1 = tile_pass_smem_bytes(ph, dtype, 1), smem2 = tile_pass_smem_bytes(ph, dtype, 2);
  int occ1 = 0, occ2 = 0;
  int rc = variant_occupancy<T, KCLASS, 1>(smem1, di, dev, &occ1);
  if (rc) return rc;
  bool two = false;
  if (g_tune_nbuf != 1 && smem2 <= size_t(di.max_smem_optin)) {
    rc = variant_occupancy<T, KCLASS, 2>(smem2, di, dev, &occ2);
    if (rc) return rc;
    two = g_tune_nbuf == 2 ? occ2 >= 1 : occ2 >= occ1;
  }
  if (two) return launch_tile_variant<T, KCLASS, 2>(state, n_qubits, prog, ph, stream, grid_override, smem2, di, occ2);
  return launch_tile_variant<T, KCLASS, 1>(state, n_qubits, prog, ph, stream, grid_override, smem1, di, occ1);
}

template <typename T>
static int launch_tile_pass_t(void* state, unsigned n_qubits, const unsigned char* prog, const HqPassHeader& ph,
                              cudaStream_t stream, int grid_override) {
  const int dtype = Traits<T>::V == 1 ? HQ_DTYPE_C64 : HQ_DTYPE_C128;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  int dev = 0;
  cudaGetDevice(&dev);
  if (ph.tile_bits > n_qubits) return int(cudaErrorInvalidValue);
  if (tile_pass_smem_bytes(ph, dtype, 1) > size_t(di.max_smem_optin)) return int(cudaErrorInvalidValue);
  const int kclass = ph.max_k <= 2 ? 0 : (ph.max_k <= 3 ? 1 : (ph.max_k <= 4 ? 2 : 3));
  switch (kclass) {
    case 0: return launch_tile_class<T, 0>(state, n_qubits, prog, ph, stream, grid_override, di, dev);
    case 1: return launch_tile_class<T, 1>(state, n_qubits, prog, ph, stream, grid_override, di, dev);
    case 2: return launch_tile_class<T, 2>(state, n_qubits, prog, ph, stream, grid_override, di, dev);
    default: return launch_tile_class<T, 3>(state, n_qubits, prog, ph, stream, grid_override, di, dev);
  }
}

int launch_tile_pass(int dtype, void* state, unsigned n_qubits, const unsigned char* prog,
                     const HqPassHeader& ph, void* stream, int grid_override) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return dtype == HQ_DTYPE_C64 ? launch_tile_pass_t<float>(state, n_qubits, prog, ph, s, grid_override)
                               : launch_tile_pass_t<double>(state, n_qubits, prog, ph, s, grid_override);
}

// ---------------------------------------------------------------------------------------
// direct single-gate kernel (no shared memory).  KK = number of UNIT-level target bits;
// LOW (complex64 only) = matrix bit 0 sits on amplitude bit 0, i.e. inside the unit.
// Every thread owns ITEMS whole work items: it issues all its 16-byte loads first, then the
// arithmetic with U read straight from the constant bank, then the stores.
// ---------------------------------------------------------------------------------------
template <typename T, int K>
struct DirectParams {
  T U[2 * (1 << K) * (1 << K)];   // row-major interleaved
  unsigned char upos[4];          // ascending unit-level target bits
};

template <typename T, int K, bool LOW, int ITEMS>
__global__ void __launch_bounds__(256)
hq_direct_kernel(typename Traits<T>::Unit* __restrict__ state, const unsigned long long n_work,
                 const DirectParams<T, K> p) {
  typedef typename Traits<T>::Unit Unit;
  const int KK = LOW ? K - 1 : K;
  const int UD = 1 << KK;
  const int DIM = 1 << K;
  const unsigned long long w0 =
      ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x);
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;

  for (unsigned long long wb = w0; wb < n_work; wb += stride * ITEMS) {
    Unit in[ITEMS][UD];
    uint64_t addr[ITEMS];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
      const unsigned long long w = wb + (unsigned long long)it * stride;
      addr[it] = open_gaps(w, p.upos, KK);
      if (w < n_work) {
#pragma unroll
        for (int m = 0; m < UD; ++m) in[it][m] = state[addr[it] | deposit(m, p.upos, KK)];
      }
    }
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
      const unsigned long long w = wb + (unsigned long long)it * stride;
      if (w >= n_work) continue;
      if (Traits<T>::V == 1 && !LOW) {
        const float4* x = reinterpret_cast<const float4*>(in[it]);
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
          float a0r = 0.f, a0i = 0.f, a1r = 0.f, a1i = 0.f;
#pragma unroll
          for (int j = 0; j < DIM; ++j) {
            const float ur = float(p.U[2 * (i * DIM + j)]), ui = float(p.U[2 * (i * DIM + j) + 1]);
            cmac(a0r, a0i, ur, ui, x[j].x, x[j].y);
            cmac(a1r, a1i, ur, ui, x[j].z, x[j].w);
          }
          float4 o = make_float4(a0r, a0i, a1r, a1i);
          st_stream(reinterpret_cast<float4*>(&state[addr[it] | deposit(i, p.upos, KK)]), o);
        }
      } else if (Traits<T>::V == 1) {
        const float4* x = reinterpret_cast<const float4*>(in[it]);
#pragma unroll
        for (int iu = 0; iu < UD; ++iu) {
          float a0r = 0.f, a0i = 0.f, a1r = 0.f, a1i = 0.f;
#pragma unroll
          for (int ju = 0; ju < UD; ++ju) {
            const int e00 = 2 * ((2 * iu) * DIM + 2 * ju), e10 = 2 * ((2 * iu + 1) * DIM + 2 * ju);
            cmac(a0r, a0i, float(p.U[e00]), float(p.U[e00 + 1]), x[ju].x, x[ju].y);
            cmac(a0r, a0i, float(p.U[e00 + 2]), float(p.U[e00 + 3]), x[ju].z, x[ju].w);
            cmac(a1r, a1i, float(p.U[e10]), float(p.U[e10 + 1]), x[ju].x, x[ju].y);
            cmac(a1r, a1i, float(p.U[e10 + 2]), float(p.U[e10 + 3]), x[ju].z, x[ju].w);
          }
          float4 o = make_float4(a0r, a0i, a1r, a1i);
          st_stream(reinterpret_cast<float4*>(&state[addr[it] | deposit(iu, p.upos, KK)]), o);
        }
      } else {
        const double2* x = reinterpret_cast<const double2*>(in[it]);
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
          double ar = 0., ai = 0.;
#pragma unroll
          for (int j = 0; j < DIM; ++j)
            cmac(ar, ai, double(p.U[2 * (i * DIM + j)]), double(p.U[2 * (i * DIM + j) + 1]), x[j].x, x[j].y);
          double2 o = make_double2(ar, ai);
          st_stream(reinterpret_cast<double2*>(&state[addr[it] | deposit(i, p.upos, KK)]), o);
        }
      }
    }
  }
}

template <typename T, int K, bool LOW, int ITEMS>
static int launch_direct_t(void* state, unsigned n_qubits, const void* U_host,
                           const unsigned* pos_sorted, cudaStream_t stream) {
  const int V = Traits<T>::V;
  const int KK = LOW ? K - 1 : K;
  DirectParams<T, K> p;
  memcpy(p.U, U_host, sizeof(p.U));
  memset(p.upos, 0, sizeof(p.upos));
  for (int i = 0; i < KK; ++i) p.upos[i] = (unsigned char)(pos_sorted[i + (LOW ? 1 : 0)] - V);
  const unsigned long long n_work = 1ull << (n_qubits - V - KK);
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  unsigned long long blocks = (n_work + 256ull * ITEMS - 1) / (256ull * ITEMS);
  const unsigned long long cap = (unsigned long long)di.sm_count * 8ull * 16ull;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  hq_direct_kernel<T, K, LOW, ITEMS><<<unsigned(blocks), 256, 0, stream>>>(
      reinterpret_cast<typename Traits<T>::Unit*>(state), n_work, p);
  return int(cudaGetLastError());
}

int launch_direct_gate(int dtype, void* state, unsigned n_qubits, const void* U_host,
                       const unsigned* pos_sorted, unsigned k, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (k < 1 || k > 3 || n_qubits < k + 1) return int(cudaErrorInvalidValue);
  if (dtype == HQ_DTYPE_C64) {
    const bool low = pos_sorted[0] == 0;
    switch (k) {
      case 1: return low ? launch_direct_t<float, 1, true, 4>(state, n_qubits, U_host, pos_sorted, s)
                         : launch_direct_t<float, 1, false, 4>(state, n_qubits, U_host, pos_sorted, s);
      case 2: return low ? launch_direct_t<float, 2, true, 4>(state, n_qubits, U_host, pos_sorted, s)
                         : launch_direct_t<float, 2, false, 2>(state, n_qubits, U_host, pos_sorted, s);
      default: return low ? launch_direct_t<float, 3, true, 2>(state, n_qubits, U_host, pos_sorted, s)
                          : launch_direct_t<float, 3, false, 1>(state, n_qubits, U_host, pos_sorted, s);
    }
  }
  switch (k) {
    case 1: return launch_direct_t<double, 1, false, 4>(state, n_qubits, U_host, pos_sorted, s);
    case 2: return launch_direct_t<double, 2, false, 2>(state, n_qubits, U_host, pos_sorted, s);
    default: return launch_direct_t<double, 3, false, 1>(state, n_qubits, U_host, pos_sorted, s);
  }
}

// ---------------------------------------------------------------------------------------
// pack / unpack
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void hq_pack_kernel(const T* __restrict__ re, const T* __restrict__ im, T* __restrict__ out,
                               unsigned long long n) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    out[2 * i] = re[i];
    out[2 * i + 1] = im[i];
  }
}
template <typename T>
__global__ void hq_unpack_kernel(const T* __restrict__ in, T* __restrict__ re, T* __restrict__ im,
                                 unsigned long long n) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    re[i] = in[2 * i];
    im[i] = in[2 * i + 1];
  }
}

static unsigned grid_for(unsigned long long n, unsigned threads) {
  DeviceInfo di;
  unsigned cap = 148 * 16;
  if (device_info(&di) == 0) cap = unsigned(di.sm_count) * 16;
  unsigned long long b = (n + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return unsigned(b);
}

int launch_pack(int dtype, const void* re, const void* im, void* out, uint64_t n, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == HQ_DTYPE_C64)
    hq_pack_kernel<float><<<grid_for(n, 256), 256, 0, s>>>((const float*)re, (const float*)im, (float*)out, n);
  else
    hq_pack_kernel<double><<<grid_for(n, 256), 256, 0, s>>>((const double*)re, (const double*)im, (double*)out, n);
  return int(cudaGetLastError());
}
int launch_unpack(int dtype, const void* in, void* re, void* im, uint64_t n, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == HQ_DTYPE_C64)
    hq_unpack_kernel<float><<<grid_for(n, 256), 256, 0, s>>>((const float*)in, (float*)re, (float*)im, n);
  else
    hq_unpack_kernel<double><<<grid_for(n, 256), 256, 0, s>>>((const double*)in, (double*)re, (double*)im, n);
  return int(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------
// out-of-place low-bit permutation of a real array
// ---------------------------------------------------------------------------------------
struct PermParams {
  unsigned char pos[32];
  unsigned m;
};
template <typename E>
__global__ void hq_bitperm_oop_kernel(const E* __restrict__ in, E* __restrict__ out,
                                      unsigned long long n, const PermParams p) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const unsigned long long mask = (1ull << p.m) - 1ull;
  for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    unsigned long long src = 0;
    for (unsigned i = 0; i < p.m; ++i) src ^= ((j >> i) & 1ull) << p.pos[i];
    out[j] = in[(j & ~mask) | src];
  }
}
int launch_bitperm_oop(int elem_bytes, const void* in, void* out, unsigned n_bits,
                       const unsigned* pos, unsigned m, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (m > 32 || m > n_bits) return int(cudaErrorInvalidValue);
  PermParams p;
  memset(&p, 0, sizeof(p));
  p.m = m;
  for (unsigned i = 0; i < m; ++i) p.pos[i] = (unsigned char)pos[i];
  const unsigned long long n = 1ull << n_bits;
  if (elem_bytes == 4)
    hq_bitperm_oop_kernel<uint32_t><<<grid_for(n, 256), 256, 0, s>>>((const uint32_t*)in, (uint32_t*)out, n, p);
  else if (elem_bytes == 8)
    hq_bitperm_oop_kernel<uint64_t><<<grid_for(n, 256), 256, 0, s>>>((const uint64_t*)in, (uint64_t*)out, n, p);
  else
    return int(cudaErrorInvalidValue);
  return int(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------
// state preparation and reductions
// ---------------------------------------------------------------------------------------
struct ProductParams {
  // per index bit b (LSB = 0): amplitude factor for bit value 0 and 1 (real: 0, 1, +-1/sqrt2)
  double f0[48], f1[48];
  unsigned n;
};
template <typename T>
__global__ void hq_init_product_kernel(T* __restrict__ state, unsigned long long n_amps, const ProductParams p) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
    double a = 1.0;
    for (unsigned b = 0; b < p.n; ++b) a *= ((i >> b) & 1ull) ? p.f1[b] : p.f0[b];
    state[2 * i] = T(a);
    state[2 * i + 1] = T(0);
  }
}
int launch_init_product(int dtype, void* state, unsigned n_qubits, const char* spec, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (n_qubits > 48) return int(cudaErrorInvalidValue);
  ProductParams p;
  memset(&p, 0, sizeof(p));
  p.n = n_qubits;
  const double r = 0.70710678118654752440;
  for (unsigned q = 0; q < n_qubits; ++q) {
    const unsigned b = n_qubits - 1 - q;   // spec[0] is the most significant bit
    switch (spec[q]) {
      case '0': p.f0[b] = 1; p.f1[b] = 0; break;
      case '1': p.f0[b] = 0; p.f1[b] = 1; break;
      case '+': p.f0[b] = r; p.f1[b] = r; break;
      case '-': p.f0[b] = r; p.f1[b] = -r; break;
      default: return int(cudaErrorInvalidValue);
    }
  }
  const unsigned long long n = 1ull << n_qubits;
  if (dtype == HQ_DTYPE_C64)
    hq_init_product_kernel<float><<<grid_for(n, 256), 256, 0, s>>>((float*)state, n, p);
  else
    hq_init_product_kernel<double><<<grid_for(n, 256), 256, 0, s>>>((double*)state, n, p);
  return int(cudaGetLastError());
}

// splitmix64-based counter RNG + Box-Muller: amplitude i depends only on (seed, i)
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
template <typename T>
__global__ void hq_init_random_kernel(T* __restrict__ state, unsigned long long n_amps,
                                      unsigned long long seed, unsigned long long index_offset) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
    const unsigned long long g = i + index_offset;
    const unsigned long long a = splitmix64(seed ^ (2 * g + 0x1234567ull));
    const unsigned long long b = splitmix64(a ^ (2 * g + 1));
    const double u1 = (double((a >> 11) + 1ull)) * (1.0 / 9007199254740993.0);
    const double u2 = double(b >> 11) * (1.0 / 9007199254740992.0);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    state[2 * i] = T(rad * cs);
    state[2 * i + 1] = T(rad * sn);
  }
}
int launch_init_random(int dtype, void* state, unsigned n_qubits, uint64_t seed, uint64_t index_offset,
                       void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const unsigned long long n = 1ull << n_qubits;
  if (dtype == HQ_DTYPE_C64)
    hq_init_random_kernel<float><<<grid_for(n, 256), 256, 0, s>>>((float*)state, n, seed, index_offset);
  else
    hq_init_random_kernel<double><<<grid_for(n, 256), 256, 0, s>>>((double*)state, n, seed, index_offset);
  return int(cudaGetLastError());
}

template <typename T>
__global__ void hq_scale_kernel(T* __restrict__ state, unsigned long long n_reals, T f) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_reals; i += stride)
    state[i] *= f;
}
int launch_scale(int dtype, void* state, uint64_t n_amps, double factor, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == HQ_DTYPE_C64)
    hq_scale_kernel<float><<<grid_for(2 * n_amps, 256), 256, 0, s>>>((float*)state, 2 * n_amps, float(factor));
  else
    hq_scale_kernel<double><<<grid_for(2 * n_amps, 256), 256, 0, s>>>((double*)state, 2 * n_amps, factor);
  return int(cudaGetLastError());
}

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double warp_part[32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) warp_part[w] = v;
  __syncthreads();
  double r = 0;
  if (w == 0) {
    r = lane < int(blockDim.x >> 5) ? warp_part[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;   // valid in thread 0
}

template <typename T>
__global__ void hq_norm2_kernel(const T* __restrict__ state, unsigned long long n_reals, double* __restrict__ partial) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  double acc = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_reals; i += stride) {
    const double v = double(state[i]);
    acc += v * v;
  }
  const double r = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}
int launch_norm2(int dtype, const void* state, uint64_t n_amps, double* partial, unsigned n_partial, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (n_partial < 1) return int(cudaErrorInvalidValue);
  if (dtype == HQ_DTYPE_C64)
    hq_norm2_kernel<float><<<n_partial, 256, 0, s>>>((const float*)state, 2 * n_amps, partial);
  else
    hq_norm2_kernel<double><<<n_partial, 256, 0, s>>>((const double*)state, 2 * n_amps, partial);
  return int(cudaGetLastError());
}

template <typename T>
__global__ void hq_vdot_kernel(const T* __restrict__ a, const T* __restrict__ b, unsigned long long n_amps,
                               double* __restrict__ partial) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  double re = 0, im = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
    const double ar = double(a[2 * i]), ai = double(a[2 * i + 1]);
    const double br = double(b[2 * i]), bi = double(b[2 * i + 1]);
    re += ar * br + ai * bi;
    im += ar * bi - ai * br;
  }
  const double r = block_sum(re);
  const double q = block_sum(im);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = r;
    partial[2 * blockIdx.x + 1] = q;
  }
}
int launch_vdot(int dtype, const void* a, const void* b, uint64_t n_amps, double* partial, unsigned n_partial,
                void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (n_partial < 1) return int(cudaErrorInvalidValue);
  if (dtype == HQ_DTYPE_C64)
    hq_vdot_kernel<float><<<n_partial, 256, 0, s>>>((const float*)a, (const float*)b, n_amps, partial);
  else
    hq_vdot_kernel<double><<<n_partial, 256, 0, s>>>((const double*)a, (const double*)b, n_amps, partial);
  return int(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------
// measurement support (device-native counterparts of the reference's FunctionalGates
// /root/reference/hybridq/gate/measure.py:25-75 and gate/projection.py:25-68)
// ---------------------------------------------------------------------------------------
struct BitsParams {
  unsigned char pos[HQ_MAX_K];   // index bit of outcome bit j
  unsigned k;
};

__device__ __forceinline__ unsigned outcome_of(unsigned long long i, const BitsParams& p) {
  unsigned s = 0;
  for (unsigned j = 0; j < p.k; ++j) s |= unsigned((i >> p.pos[j]) & 1ull) << j;
  return s;
}

// out[2 s], out[2 s + 1] += sum over amplitudes with outcome s of re^2, im^2 (global double atomics;
// `out` must be zeroed).  A thread keeps a running sum and flushes it to the CTA's shared-memory
// histogram only when its outcome changes, the CTA flushes once at the end.
template <typename T>
__global__ void __launch_bounds__(256) hq_marginal_kernel(const T* __restrict__ state, unsigned long long n_amps,
                                                          const BitsParams p, double* __restrict__ out) {
  extern __shared__ double hist[];
  const unsigned bins = 2u << p.k;
  for (unsigned b = threadIdx.x; b < bins; b += blockDim.x) hist[b] = 0.0;
  __syncthreads();
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  double are = 0, aim = 0;
  unsigned cur = 0;
  bool have = false;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
    const unsigned s = outcome_of(i, p);
    if (have && s != cur) {
      atomicAdd(&hist[2 * cur], are);
      atomicAdd(&hist[2 * cur + 1], aim);
      are = aim = 0;
    }
    cur = s;
    have = true;
    const double re = double(state[2 * i]), im = double(state[2 * i + 1]);
    are += re * re;
    aim += im * im;
  }
  if (have) {
    atomicAdd(&hist[2 * cur], are);
    atomicAdd(&hist[2 * cur + 1], aim);
  }
  __syncthreads();
  for (unsigned b = threadIdx.x; b < bins; b += blockDim.x)
    if (hist[b] != 0.0) atomicAdd(&out[b], hist[b]);
}

int launch_marginal(int dtype, const void* state, unsigned n_qubits, const unsigned* pos, unsigned k, double* out_dev,
                    void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (k > HQ_MAX_K || k > n_qubits) return int(cudaErrorInvalidValue);
  BitsParams p;
  memset(&p, 0, sizeof(p));
  p.k = k;
  for (unsigned j = 0; j < k; ++j) {
    if (pos[j] >= n_qubits) return int(cudaErrorInvalidValue);
    p.pos[j] = (unsigned char)pos[j];
  }
  const unsigned long long n = 1ull << n_qubits;
  const size_t smem = (size_t(2) << k) * sizeof(double);
  if (dtype == HQ_DTYPE_C64)
    hq_marginal_kernel<float><<<grid_for(n, 256), 256, smem, s>>>((const float*)state, n, p, out_dev);
  else
    hq_marginal_kernel<double><<<grid_for(n, 256), 256, smem, s>>>((const double*)state, n, p, out_dev);
  return int(cudaGetLastError());
}

// amplitudes whose outcome differs from `outcome` become 0, the others are scaled plane-wise
template <typename T>
__global__ void __launch_bounds__(256) hq_project_kernel(T* __restrict__ state, unsigned long long n_amps, const BitsParams p,
                                                         unsigned outcome, T scale_re, T scale_im) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
    const bool keep = outcome_of(i, p) == outcome;
    typename Traits<T>::Cplx v = reinterpret_cast<typename Traits<T>::Cplx*>(state)[i];
    v.x = keep ? v.x * scale_re : T(0);
    v.y = keep ? v.y * scale_im : T(0);
    reinterpret_cast<typename Traits<T>::Cplx*>(state)[i] = v;
  }
}

int launch_project(int dtype, void* state, unsigned n_qubits, const unsigned* pos, unsigned k, unsigned outcome,
                   double scale_re, double scale_im, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (k > HQ_MAX_K || k > n_qubits || outcome >= (1u << k)) return int(cudaErrorInvalidValue);
  BitsParams p;
  memset(&p, 0, sizeof(p));
  p.k = k;
  for (unsigned j = 0; j < k; ++j) {
    if (pos[j] >= n_qubits) return int(cudaErrorInvalidValue);
    p.pos[j] = (unsigned char)pos[j];
  }
  const unsigned long long n = 1ull << n_qubits;
  if (dtype == HQ_DTYPE_C64)
    hq_project_kernel<float><<<grid_for(n, 256), 256, 0, s>>>((float*)state, n, p, outcome, float(scale_re), float(scale_im));
  else
    hq_project_kernel<double><<<grid_for(n, 256), 256, 0, s>>>((double*)state, n, p, outcome, scale_re, scale_im);
  return int(cudaGetLastError());
}

}  // namespace hq
