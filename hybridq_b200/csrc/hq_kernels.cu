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
#include <cstdlib>
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
#ifndef HQ_TILE_QUEUE
#define HQ_TILE_QUEUE 0   // 1: persistent grid + global tile counter instead of one CTA per tile (experiment)
#endif
#ifndef HQ_K1F_BLOCKS
#define HQ_K1F_BLOCKS 2   // resident CTAs per SM of the k <= 3 complex64 tile kernel (128 registers: the k = 3 FFMA2 slots keep 16 accumulator pairs)
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
// (a function of its own: its register needs -- the compact sparse form keeps four units and their constants live --
// must not weigh on the allocation of the tensor-core and register paths below)
template <typename Unit>
__device__ __noinline__ void gate_dr1_outofline(Unit* tile, const HqGateDesc* g, uint32_t k, const unsigned char* prog,
                                                uint32_t mat_off, int Tu, int tid) {
  gate_dr1_dispatch(tile, g, k, prog, mat_off, Tu, tid);
}

template <int MAXK, int MMAK, typename Unit>
__device__ __noinline__ void gate_small_generic(Unit* tile, const HqGateDesc* g, uint32_t k, uint32_t kind,
                                                const unsigned char* prog, uint32_t mat_off, int Tu, int tid) {
  if (kind == HQ_GATE_MMA) {
    gate_mma_dispatch<MMAK>(tile, g, k, prog, tid);
  } else if (kind == HQ_GATE_DR1) {
    if (MAXK >= 3) gate_dr1_outofline(tile, g, k, prog, mat_off, Tu, tid);     // scalar + rank one (k = 3, 4)
  } else if (IsF64Unit<Unit>::value && kind == HQ_GATE_ROWPAIR) {
    rowpair_dispatch<MAXK>(tile, g, k, prog, mat_off, Tu, tid);
  } else {
    const bool low = !IsF64Unit<Unit>::value && __ldg(&g->tpos[0]) == 0;
    gate_small_dispatch<MAXK>(tile, g, k, low, prog, mat_off, Tu, tid);
  }
}

// Who synchronises between two gates of a tile: the whole CTA (hq_tile_kernel: one tile per CTA at a time) or
// one 256-thread consumer group through a named barrier (hq_ring_kernel: two groups per CTA, each on its own tile).
struct CtaSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
struct GroupSync {
  int id;   // named barrier 1 + group
  __device__ __forceinline__ void operator()() const {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(HQ_THREADS) : "memory");
  }
};

// One out-of-line copy per kernel of the constant-bank FFMA2 gates (k = 1..3, with / without a target on
// amplitude bit 0): the slot index is a run-time (CTA-uniform) value, so the matrix elements are fetched with
// uniform constant loads at a register offset and there is one code copy, not one per slot.
#ifndef HQ_FAST_TEMPLATED
#define HQ_FAST_TEMPLATED 1   // 1: one code copy per slot, immediate constant-bank offsets; 0: run-time slot index
#endif
// (inlined on purpose: `ph` must stay the kernel's own __grid_constant__ parameter for its elements to be
// constant-bank operands; behind a call it decays to a generic pointer and every element becomes an LD)
template <int MAXK>
__device__ __forceinline__ void gate_fast_slot(float4* tile, const StreamRegs& g, const HqPassHeader& ph, uint32_t slot,
                                               int Tu, int tid) {
#if HQ_FAST_TEMPLATED
  switch (slot) {
    case 0: gate_fast_f32<0, MAXK>(tile, g, ph, 0u, Tu, tid); break;
    case 1: gate_fast_f32<1, MAXK>(tile, g, ph, 1u, Tu, tid); break;
    case 2: gate_fast_f32<2, MAXK>(tile, g, ph, 2u, Tu, tid); break;
    case 3: gate_fast_f32<3, MAXK>(tile, g, ph, 3u, Tu, tid); break;
    case 4: gate_fast_f32<4, MAXK>(tile, g, ph, 4u, Tu, tid); break;
    case 5: gate_fast_f32<5, MAXK>(tile, g, ph, 5u, Tu, tid); break;
    case 6: gate_fast_f32<6, MAXK>(tile, g, ph, 6u, Tu, tid); break;
    default: gate_fast_f32<7, MAXK>(tile, g, ph, 7u, Tu, tid); break;
  }
#else
  gate_fast_f32<-1, MAXK>(tile, g, ph, slot, Tu, tid);
#endif
}
template <int MAXK>
__device__ __forceinline__ void gate_fast_slot(double2*, const StreamRegs&, const HqPassHeader&, uint32_t, int, int) {}

// Every gate of the pass on one tile held in shared memory; `sync` separates consecutive gates and is also
// called after the last one (the drain reads what other threads wrote).
// FAST: the kernel variant carries the constant-bank FFMA2 slots (complex64, passes whose gates all have k <= 3);
// other variants send flagged gates down the generic path (the planner stores their matrices in the program too).
template <typename T, int KCLASS, bool FAST, class Sync>
__device__ __forceinline__ void apply_pass_gates(typename Traits<T>::Unit* tile, const HqGateDesc* gates,
                                                 const HqPassHeader& ph, const unsigned char* prog, int Tbits, int Tu,
                                                 int tid, const Sync& sync) {
  typedef typename Traits<T>::Cplx Cplx;
  const int V = Traits<T>::V;
  const int MAXK = KCLASS == 0 ? 2 : (KCLASS == 1 ? 3 : 4);
  const int MMAK = KCLASS == 3 ? HQ_MMA_MAX_K : MAXK;
  const uint32_t n_gates = ph.n_gates;
  const bool any_fast = FAST && V == 1 && KCLASS <= 1 && ph.fast_mask != 0;
  StreamRegs sr;
  if (any_fast && (ph.fast_mask & 1u)) sr = load_stream_regs(gates, tid);
  for (uint32_t gi = 0; gi < n_gates; ++gi) {
    const HqGateDesc* g = gates + gi;
    if (any_fast && gi < 32u && ((ph.fast_mask >> gi) & 1u)) {
      // complex64 k <= 3: constant-bank FFMA2; slot = how many slot matrices came before (other kinds in between --
      // e.g. the channels of a density-matrix pass -- do not use up slots)
      gate_fast_slot<MAXK>(tile, sr, ph, uint32_t(__popc(ph.fast_mask & ((1u << gi) - 1u))), Tu, tid);
      // the next slot gate's addressing constants travel while this thread waits at the barrier
      if (gi + 1 < n_gates && gi + 1 < 32u && ((ph.fast_mask >> (gi + 1)) & 1u)) sr = load_stream_regs(g + 1, tid);
      if ((ph.chain_mask >> gi) & 1u) __syncwarp();      // the next gate's warps work on the very same units
      else sync();
      continue;
    }
    if (any_fast && gi + 1 < n_gates && gi + 1 < 32u && ((ph.fast_mask >> (gi + 1)) & 1u)) sr = load_stream_regs(g + 1, tid);
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
        sync();
        gate_big_phaseB<T>(reinterpret_cast<Cplx*>(tile), *g, acc);
      }
    }
    sync();
  }
}

struct HqXchg {            // exchange redirect of the drain (all zero = plain in-place pass)
  uint32_t s;              // number of local index bits swapped with rank bits (0 .. 3)
  uint32_t mine;           // this rank's digit (value of the rank bits being swapped), deposited at pos[]
  uint8_t pos[4];          // local AMPLITUDE-bit positions leaving the shard (>= V)
  uint32_t reserved;
  void* dst[8];            // destination buffer of digit D (dst[mine] is this rank's own second buffer)
  const void* src;         // tile kernel only: read the tiles from here instead of `state` (null = state)
  unsigned long long* queue;   // tile kernel, HQ_TILE_QUEUE builds: tile counter of a persistent grid (null = static)
};

// XCHG (NBUF = 1 only): the drain writes to the buffers of HqXchg instead of back in place (see hq_ring_kernel)
template <typename T, int KCLASS, int NBUF, bool XCHG = false>
__global__ void __launch_bounds__(HQ_THREADS, ((KCLASS == 0 && Traits<T>::V == 1) ? HQ_K0F_BLOCKS
                                : ((KCLASS == 1 && Traits<T>::V == 0) ? HQ_K1D_BLOCKS
                                   : ((KCLASS == 1 && Traits<T>::V == 1) ? HQ_K1F_BLOCKS : (KCLASS == 0 ? 3 : 2)))))
hq_tile_kernel(typename Traits<T>::Unit* __restrict__ state, const unsigned char* __restrict__ prog,
               const __grid_constant__ HqPassHeader ph, const unsigned long long n_tiles,
               const __grid_constant__ HqXchg xg) {
  typedef typename Traits<T>::Unit Unit;
  typedef typename Traits<T>::Cplx Cplx;
  const int V = Traits<T>::V;
  extern __shared__ __align__(16) unsigned char smem[];

  const int tid = threadIdx.x;
  const int Tbits = int(ph.tile_bits);
  const int h = int(ph.n_high);
  const int Tu = Tbits - V;
  const int Lu = Tbits - h - V;
  const uint32_t n_units = 1u << Tu;
  const int npt = Tu > HQ_THREADS_LOG2 ? (1 << (Tu - HQ_THREADS_LOG2)) : (uint32_t(tid) < n_units ? 1 : 0);

  Unit* bufs = reinterpret_cast<Unit*>(smem);
  const HqGateDesc* gates = reinterpret_cast<const HqGateDesc*>(prog + ph.gates_off);

  // per-thread constants of the fill/drain addressing
  const uint64_t off_t = unit_offset(uint32_t(tid), Lu, V, ph.high_pos, h);
  const uint32_t swz_t = swz(uint32_t(tid));

  unsigned long long t = blockIdx.x;
#if HQ_TILE_QUEUE
  // persistent grid fed from a global tile counter: the next tile index is fetched while the current tile is being
  // processed (thread 0, right after the fill barrier) and read after the barrier that ends the tile
  // (the slot lives behind the tile in the dynamic shared memory: tile_pass_smem_bytes adds 16 bytes)
  volatile unsigned long long& s_next = *reinterpret_cast<volatile unsigned long long*>(smem + (size_t(16 * NBUF) << Tu));
  unsigned long long* const queue = NBUF == 1 ? xg.queue : nullptr;
  if (queue) {
    if (tid == 0) s_next = atomicAdd(queue, 1ull);
    __syncthreads();
    t = s_next;
  }
#endif
  int cur = 0;
  if (NBUF == 2 && t < n_tiles) {
    tile_fill<T>(bufs, state + (tile_base(t, Tbits, h, ph.high_pos) >> V) + off_t, ph, swz_t, npt);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  const Unit* const fill_src = (XCHG && xg.src) ? reinterpret_cast<const Unit*>(xg.src) : state;
#if HQ_TILE_QUEUE
  for (; t < n_tiles; t = queue ? s_next : t + gridDim.x) {
#else
  for (; t < n_tiles; t += gridDim.x) {
#endif
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
      tile_fill<T>(tile, fill_src + (gptr - state), ph, swz_t, npt);
      cp_async_wait_all();
    }
    __syncthreads();
#if HQ_TILE_QUEUE
    if (queue && tid == 0) s_next = atomicAdd(queue, 1ull);     // everybody has read the previous value by now
#endif

    apply_pass_gates<T, KCLASS, true>(tile, gates, ph, prog, Tbits, Tu, tid, CtaSync());

    // drain
    if (XCHG) {
      // exchange redirect: unit u of the shard goes to buffer dst[D(u)] at u with the swapped bits set to `mine`
      uint64_t xmask = 0, xmine = 0;
      for (uint32_t j = 0; j < xg.s; ++j) {
        xmask |= uint64_t(1) << (xg.pos[j] - V);
        xmine |= uint64_t((xg.mine >> j) & 1u) << (xg.pos[j] - V);
      }
      const uint64_t ubase = uint64_t(gptr - state);
      const Cplx* amps = reinterpret_cast<const Cplx*>(tile);
#pragma unroll 4
      for (int i = 0; i < npt; ++i) {
        const uint64_t u = ubase + ph.iter_off[i];
        Unit v;
        if (!ph.has_perm) {
          v = tile[swz_t ^ ph.iter_swz[i]];
        } else {
          const uint32_t c = uint32_t(tid) + (uint32_t(i) << HQ_THREADS_LOG2);
          Cplx o[1 << V];
#pragma unroll
          for (uint32_t e = 0; e < (1u << V); ++e) o[e] = amps[amp_slot<T>(perm_src((c << V) | e, ph.perm, Tbits))];
          v = make_unit(o);
        }
        uint32_t D = 0;
        for (uint32_t j = 0; j < xg.s; ++j) D |= uint32_t((u >> (xg.pos[j] - V)) & 1ull) << j;
        st_stream(reinterpret_cast<Unit*>(xg.dst[D]) + ((u & ~xmask) | xmine), v);
      }
    } else if (!ph.has_perm) {
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

static bool persistent_grid() {
  static int persistent = -1;
  if (persistent < 0) {
    const char* e = getenv("HQ_PERSISTENT");
    persistent = (e && atoi(e) > 0) ? 1 : 0;
  }
  return persistent == 1;
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
  return (size_t(16 * nbuf) << Tu) + (HQ_TILE_QUEUE ? 16 : 0);
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
template <typename T, int KCLASS, int NBUF, bool XCHG = false>
static int variant_occupancy(size_t smem, const DeviceInfo& di, int dev, int* per_sm_out) {
  static bool attr_set[64];
  static int cache[64][HQ_MAX_UNIT_BITS + 2];
  auto kern = hq_tile_kernel<T, KCLASS, NBUF, XCHG>;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin);
    if (e != cudaSuccess) return int(e);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return int(e);
    for (int i = 0; i < HQ_MAX_UNIT_BITS + 2; ++i) cache[dev][i] = -1;
    attr_set[dev] = true;
  }
  int slot = 0;
  while ((size_t(16 * NBUF) << slot) + (HQ_TILE_QUEUE ? 16 : 0) < smem && slot < HQ_MAX_UNIT_BITS + 1) ++slot;
  if (cache[dev][slot] < 0) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, HQ_THREADS, smem);
    if (e != cudaSuccess) return int(e);
    cache[dev][slot] = per_sm;
  }
  *per_sm_out = cache[dev][slot];
  return 0;
}

template <typename T, int KCLASS, int NBUF, bool XCHG = false>
static int launch_tile_variant(void* state, unsigned n_qubits, const unsigned char* prog, const HqPassHeader& ph,
                               cudaStream_t stream, int grid_override, size_t smem, const DeviceInfo& di, int per_sm,
                               const HqXchg* xgp = nullptr) {
  HqXchg xg;
  if (xgp) xg = *xgp;
  else memset(&xg, 0, sizeof(xg));
#if HQ_TILE_QUEUE
  bool use_queue = false;
  if (NBUF == 1 && grid_override <= 0) {
    // a ring of 64 tile counters per device; each launch zeroes the one it uses in stream order
    static unsigned long long* q[64];
    static unsigned qi[64];
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!q[dev] && cudaMalloc(reinterpret_cast<void**>(&q[dev]), 64 * sizeof(unsigned long long)) != cudaSuccess) q[dev] = nullptr;
    if (q[dev]) {
      xg.queue = q[dev] + (qi[dev]++ & 63u);
      if (cudaMemsetAsync(xg.queue, 0, sizeof(unsigned long long), stream) == cudaSuccess) use_queue = true;
      else xg.queue = nullptr;
    }
  }
#endif
  const unsigned long long n_tiles = 1ull << (n_qubits - ph.tile_bits);
  if (per_sm < 1) return int(cudaErrorLaunchOutOfResources);
  if (g_tune_ctas_per_sm > 0 && g_tune_ctas_per_sm < per_sm) per_sm = g_tune_ctas_per_sm;
  unsigned long long grid = (unsigned long long)di.sm_count * (unsigned long long)per_sm;
  {
    // One CTA per tile by default: MEASURED on B200 (profiles/r02/bench_o_*.log, n = 30 bench circuit) a persistent
    // grid of SMs x resident CTAs with a static round-robin of tiles takes 129.0 ms/step, 4 / 32 / 128 times as
    // many CTAs 125.0 / 123.5 / 119.7 ms and one CTA per tile 117.4 ms: the SMs do not run at one speed (two dies,
    // near / far L2 slices), and the hardware CTA scheduler refills whichever SM frees up.  HQ_PERSISTENT=1 (or a
    // grid_override) restores the persistent form; the kernel's tile loop handles either.
#if HQ_TILE_QUEUE
    if (!persistent_grid() && !use_queue) grid = n_tiles;
#else
    if (!persistent_grid()) grid = n_tiles;
#endif
    if (grid > (1ull << 30)) grid = 1ull << 30;
  }
  if (grid_override > 0) grid = (unsigned long long)grid_override;
  if (grid > n_tiles) grid = n_tiles;
  hq_tile_kernel<T, KCLASS, NBUF, XCHG><<<unsigned(grid), HQ_THREADS, smem, stream>>>(
      reinterpret_cast<typename Traits<T>::Unit*>(state), prog, ph, n_tiles, xg);
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
  // (with one CTA per tile there is no next tile to prefetch: the double-buffered variant only makes sense on a
  // persistent grid, or when asked for explicitly)
  if (g_tune_nbuf != 1 && (persistent_grid() || g_tune_nbuf == 2) && smem2 <= size_t(di.max_smem_optin)) {
    rc = variant_occupancy<T, KCLASS, 2>(smem2, di, dev, &occ2);
    if (rc) return rc;
    two = g_tune_nbuf == 2 ? occ2 >= 1 : occ2 >= occ1;
  }
  if (two) return launch_tile_variant<T, KCLASS, 2>(state, n_qubits, prog, ph, stream, grid_override, smem2, di, occ2);
  return launch_tile_variant<T, KCLASS, 1>(state, n_qubits, prog, ph, stream, grid_override, smem1, di, occ1);
}

// exchange-redirect pass on the tile kernel (single-buffered, one CTA per tile)
template <typename T, int KCLASS>
static int launch_tile_xchg_class(void* state, unsigned n_qubits, const unsigned char* prog, const HqPassHeader& ph,
                                  const HqXchg& xg, cudaStream_t stream, int grid_override, const DeviceInfo& di, int dev) {
  const int dtype = Traits<T>::V == 1 ? HQ_DTYPE_C64 : HQ_DTYPE_C128;
  const size_t smem1 = tile_pass_smem_bytes(ph, dtype, 1);
  int occ = 0;
  const int rc = variant_occupancy<T, KCLASS, 1, true>(smem1, di, dev, &occ);
  if (rc) return rc;
  return launch_tile_variant<T, KCLASS, 1, true>(state, n_qubits, prog, ph, stream, grid_override, smem1, di, occ, &xg);
}
template <typename T>
static int launch_tile_xchg_t(void* state, unsigned n_qubits, const unsigned char* prog, const HqPassHeader& ph,
                              const HqXchg& xg, cudaStream_t stream, int grid_override, const DeviceInfo& di, int dev) {
  const int kclass = ph.max_k <= 2 ? 0 : (ph.max_k <= 3 ? 1 : (ph.max_k <= 4 ? 2 : 3));
  switch (kclass) {
    case 0: return launch_tile_xchg_class<T, 0>(state, n_qubits, prog, ph, xg, stream, grid_override, di, dev);
    case 1: return launch_tile_xchg_class<T, 1>(state, n_qubits, prog, ph, xg, stream, grid_override, di, dev);
    case 2: return launch_tile_xchg_class<T, 2>(state, n_qubits, prog, ph, xg, stream, grid_override, di, dev);
    default: return launch_tile_xchg_class<T, 3>(state, n_qubits, prog, ph, xg, stream, grid_override, di, dev);
  }
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
// the ring kernel: the same pass as hq_tile_kernel, software-pipelined inside ONE persistent CTA per SM
//
//   warps  0 ..  7   consumer group 0   (256 threads each: exactly the thread count the gate code and the lane
//   warps  8 .. 15   consumer group 1    tables of HqGateDesc are written for)
//
// Shared memory is a ring of HQ_RING_STAGES = 3 tile buffers.  Tile i of the CTA lives in stage i % 3 and is
// processed by group i % 2: the group waits on the stage's `full` mbarrier, applies every gate of the pass with
// named barriers between gates, streams the result back to HBM straight from shared memory, and then -- the
// stage being free again -- issues the 16-byte cp.async copies (LDGSTS: no registers, arbitrary swizzled slots)
// of tile i + 3 into it; their completion is signalled on `full` by cp.async.mbarrier.arrive.noinc, one arrival
// per thread of the group.  So at any time two tiles are being computed on while a third is in flight from HBM
// and the drain stores of the previous ones are still retiring: a group never waits for its own fill, the fill
// it issues is consumed by the OTHER group half a tile time later.  (A first version had a 17th warp as a
// dedicated producer and `empty` mbarriers; 17 warps cap the kernel at 96 registers -- five warps share one
// sub-partition's register file -- which the k = 3 FFMA2 slots do not fit in.)
//
// XCHG: the drain writes to OTHER buffers -- dst[D] is the buffer (local or a peer GPU's, mapped over NVLink)
// that receives the amplitudes whose local index bits xg.pos[] spell D, stored at the same local index with
// those bits replaced by this rank's own digit: a rank-bit <-> local-bit exchange fused into the last pass
// before it (hybridq_b200/dist.py), no separate permutation pass, no staging copy.
// ---------------------------------------------------------------------------------------
#define HQ_RING_STAGES 3
#define HQ_RING_GROUPS 2
#define HQ_RING_THREADS (HQ_RING_GROUPS * HQ_THREADS)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
// every cp.async this thread has issued so far arrives on `bar` when it has landed (no pending-count increment:
// the arrival was budgeted at mbar_init)
__device__ __forceinline__ void mbar_arrive_cp_async(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// Wait for the phase with the given parity to complete.  A protocol bug would spin forever and hang the GPU;
// after ~2^26 failed probes (seconds) the kernel traps instead, which surfaces as a launch failure.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_addr(bar);
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (!done && spins > (1u << 26)) __trap();
  }
}


template <typename T, int KCLASS, bool XCHG>
__global__ void __launch_bounds__(HQ_RING_THREADS, 1)
hq_ring_kernel(typename Traits<T>::Unit* __restrict__ state, const unsigned char* __restrict__ prog,
               const __grid_constant__ HqPassHeader ph, const unsigned long long n_tiles,
               const __grid_constant__ HqXchg xg) {
  typedef typename Traits<T>::Unit Unit;
  typedef typename Traits<T>::Cplx Cplx;
  const int V = Traits<T>::V;
  extern __shared__ __align__(16) unsigned char smem[];

  const int Tbits = int(ph.tile_bits);
  const int h = int(ph.n_high);
  const int Tu = Tbits - V;
  const int Lu = Tbits - h - V;
  const int npt = 1 << (Tu - HQ_THREADS_LOG2);               // the launcher guarantees Tu >= HQ_THREADS_LOG2
  Unit* const bufs = reinterpret_cast<Unit*>(smem);
  uint64_t* const full = reinterpret_cast<uint64_t*>(smem + (size_t(16 * HQ_RING_STAGES) << Tu));

  if (threadIdx.x == 0) {
    for (int st = 0; st < HQ_RING_STAGES; ++st) mbar_init(&full[st], HQ_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const unsigned long long first = blockIdx.x;
  const unsigned long long count = first < n_tiles ? (n_tiles - first + gridDim.x - 1) / gridDim.x : 0ull;
  // broadcast from lane 0 so that the compiler knows the group is warp-uniform (uniform-register operands)
  const int group = __shfl_sync(0xffffffffu, int(threadIdx.x >> HQ_THREADS_LOG2), 0);
  const int tid = threadIdx.x & (HQ_THREADS - 1);
  const GroupSync sync{1 + group};
  const HqGateDesc* gates = reinterpret_cast<const HqGateDesc*>(prog + ph.gates_off);
  const uint64_t off_t = unit_offset(uint32_t(tid), Lu, V, ph.high_pos, h);
  const uint32_t swz_t = swz(uint32_t(tid));
  uint64_t xmask = 0, xmine = 0;
  if (XCHG) {
    for (uint32_t j = 0; j < xg.s; ++j) {
      xmask |= uint64_t(1) << (xg.pos[j] - V);
      xmine |= uint64_t((xg.mine >> j) & 1u) << (xg.pos[j] - V);
    }
  }

  // this thread's share of the fill of the CTA's tile number i into its stage
  auto issue_fill = [&](unsigned long long i) {
    const int st = int(i % HQ_RING_STAGES);
    const unsigned long long t = first + i * gridDim.x;
    tile_fill<T>(bufs + (size_t(st) << Tu), state + (tile_base(t, Tbits, h, ph.high_pos) >> V) + off_t, ph, swz_t, npt);
    mbar_arrive_cp_async(&full[st]);
  };
  // prologue: tiles 0 and 2 are fetched by group 0, tile 1 by group 1
  for (unsigned long long i = group; i < HQ_RING_STAGES && i < count; i += HQ_RING_GROUPS) issue_fill(i);

  for (unsigned long long i = group; i < count; i += HQ_RING_GROUPS) {
    const int st = int(i % HQ_RING_STAGES);
    mbar_wait(&full[st], uint32_t((i / HQ_RING_STAGES) & 1ull));
    Unit* tile = bufs + (size_t(st) << Tu);
    const unsigned long long t = first + i * gridDim.x;
    const uint64_t ubase = (tile_base(t, Tbits, h, ph.high_pos) >> V) + off_t;

    if (ph.n_gates) apply_pass_gates<T, KCLASS, true>(tile, gates, ph, prog, Tbits, Tu, tid, sync);

    // drain straight from shared memory
    if (!ph.has_perm) {
#pragma unroll 4
      for (int it = 0; it < npt; ++it) {
        const uint64_t u = ubase + ph.iter_off[it];
        const Unit v = tile[swz_t ^ ph.iter_swz[it]];
        if (XCHG) {
          uint32_t D = 0;
          for (uint32_t j = 0; j < xg.s; ++j) D |= uint32_t((u >> (xg.pos[j] - V)) & 1ull) << j;
          st_stream(reinterpret_cast<Unit*>(xg.dst[D]) + ((u & ~xmask) | xmine), v);
        } else {
          st_stream(state + u, v);
        }
      }
    } else {
      const Cplx* amps = reinterpret_cast<const Cplx*>(tile);
      for (int it = 0; it < npt; ++it) {
        const uint32_t c = uint32_t(tid) + (uint32_t(it) << HQ_THREADS_LOG2);
        Cplx o[1 << V];
#pragma unroll
        for (uint32_t e = 0; e < (1u << V); ++e) o[e] = amps[amp_slot<T>(perm_src((c << V) | e, ph.perm, Tbits))];
        const uint64_t u = ubase + ph.iter_off[it];
        if (XCHG) {
          uint32_t D = 0;
          for (uint32_t j = 0; j < xg.s; ++j) D |= uint32_t((u >> (xg.pos[j] - V)) & 1ull) << j;
          st_stream(reinterpret_cast<Unit*>(xg.dst[D]) + ((u & ~xmask) | xmine), make_unit(o));
        } else {
          st_stream(state + u, make_unit(o));
        }
      }
    }
    if (i + HQ_RING_STAGES < count) {
      sync();                       // every thread of the group has read its part of the stage
      issue_fill(i + HQ_RING_STAGES);
    }
  }
  // no cp.async is left in flight: every fill that was issued has been waited for by one of the groups
}

static int g_tune_ring = -1;         // -1 = auto (large states), 0 = never, 1 = whenever the tile allows it
void set_ring(int mode) {
  if (mode >= -1 && mode <= 1) g_tune_ring = mode;
}

size_t ring_smem_bytes(const HqPassHeader& ph, int dtype) {
  const int V = dtype == HQ_DTYPE_C64 ? 1 : 0;
  const int Tu = int(ph.tile_bits) - V;
  return (size_t(16 * HQ_RING_STAGES) << Tu) + HQ_RING_STAGES * sizeof(uint64_t);
}

// can this pass run on the ring kernel?  (tile of at least one unit per consumer thread)
static bool ring_eligible(const HqPassHeader& ph, int dtype, unsigned n_qubits, const DeviceInfo& di) {
  const int V = dtype == HQ_DTYPE_C64 ? 1 : 0;
  const int Tu = int(ph.tile_bits) - V;
  if (Tu < HQ_THREADS_LOG2 || ph.tile_bits > n_qubits) return false;
  return ring_smem_bytes(ph, dtype) <= size_t(di.max_smem_optin);
}

template <typename T, int KCLASS, bool XCHG>
static int launch_ring_variant(void* state, unsigned n_qubits, const unsigned char* prog, const HqPassHeader& ph,
                               const HqXchg& xg, cudaStream_t stream, int grid_override, const DeviceInfo& di, int dev) {
  static bool attr_set[64];
  auto kern = hq_ring_kernel<T, KCLASS, XCHG>;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin);
    if (e != cudaSuccess) return int(e);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return int(e);
    attr_set[dev] = true;
  }
  const int dtype = Traits<T>::V == 1 ? HQ_DTYPE_C64 : HQ_DTYPE_C128;
  const unsigned long long n_tiles = 1ull << (n_qubits - ph.tile_bits);
  unsigned long long grid = (unsigned long long)di.sm_count;
  if (grid_override > 0) grid = (unsigned long long)grid_override;
  if (grid > n_tiles) grid = n_tiles;
  kern<<<unsigned(grid), HQ_RING_THREADS, ring_smem_bytes(ph, dtype), stream>>>(
      reinterpret_cast<typename Traits<T>::Unit*>(state), prog, ph, n_tiles, xg);
  return int(cudaGetLastError());
}

template <typename T, bool XCHG>
static int launch_ring_t(void* state, unsigned n_qubits, const unsigned char* prog, const HqPassHeader& ph,
                         const HqXchg& xg, cudaStream_t stream, int grid_override, const DeviceInfo& di, int dev) {
  const int kclass = ph.max_k <= 2 ? 0 : (ph.max_k <= 3 ? 1 : (ph.max_k <= 4 ? 2 : 3));
  switch (kclass) {
    case 0: return launch_ring_variant<T, 0, XCHG>(state, n_qubits, prog, ph, xg, stream, grid_override, di, dev);
    case 1: return launch_ring_variant<T, 1, XCHG>(state, n_qubits, prog, ph, xg, stream, grid_override, di, dev);
    case 2: return launch_ring_variant<T, 2, XCHG>(state, n_qubits, prog, ph, xg, stream, grid_override, di, dev);
    default: return launch_ring_variant<T, 3, XCHG>(state, n_qubits, prog, ph, xg, stream, grid_override, di, dev);
  }
}

// The pass entry point: ring kernel for large states (or when forced), hq_tile_kernel otherwise.
// xchg != nullptr (an exchange redirect, see HqXchg) always takes the ring kernel.
int launch_pass(int dtype, void* state, unsigned n_qubits, const unsigned char* prog, const HqPassHeader& ph,
                const HqXchgDesc* xchg, void* stream, int grid_override) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  int dev = 0;
  cudaGetDevice(&dev);
  if (ph.tile_bits > n_qubits) return int(cudaErrorInvalidValue);
  const bool can_ring = ring_eligible(ph, dtype, n_qubits, di);
  HqXchg xg;
  memset(&xg, 0, sizeof(xg));
  if (xchg && (xchg->s || xchg->src || xchg->dst[0])) {
    if (!can_ring || xchg->s > 3) return int(cudaErrorInvalidValue);
    xg.src = xchg->src;
    const unsigned V = dtype == HQ_DTYPE_C64 ? 1u : 0u;
    xg.s = xchg->s;
    xg.mine = xchg->mine;
    for (unsigned j = 0; j < xchg->s; ++j) {
      if (xchg->pos[j] < V || xchg->pos[j] >= n_qubits) return int(cudaErrorInvalidValue);
      xg.pos[j] = (uint8_t)xchg->pos[j];
    }
    for (unsigned d = 0; d < (1u << xchg->s); ++d) {
      if (!xchg->dst[d] && xchg->s) return int(cudaErrorInvalidValue);
      xg.dst[d] = xchg->dst[d] ? xchg->dst[d] : state;      // s = 0: in place unless a destination is given
    }
    // which kernel carries the redirect: the tile kernel (three CTAs per SM, one per tile) unless HQ_XCHG_RING=1
    static int xchg_ring = -1;
    if (xchg_ring < 0) {
      const char* e = getenv("HQ_XCHG_RING");
      xchg_ring = (e && atoi(e) > 0) ? 1 : 0;
    }
    if ((!xchg_ring || xg.src) && tile_pass_smem_bytes(ph, dtype, 1) <= size_t(di.max_smem_optin))
      return dtype == HQ_DTYPE_C64 ? launch_tile_xchg_t<float>(state, n_qubits, prog, ph, xg, s, grid_override, di, dev)
                                   : launch_tile_xchg_t<double>(state, n_qubits, prog, ph, xg, s, grid_override, di, dev);
    return dtype == HQ_DTYPE_C64 ? launch_ring_t<float, true>(state, n_qubits, prog, ph, xg, s, grid_override, di, dev)
                                 : launch_ring_t<double, true>(state, n_qubits, prog, ph, xg, s, grid_override, di, dev);
  }
  // auto: measured on B200 (profiles/r02/sweep_ring_b.jsonl) the ring kernel only wins on passes that are purely
  // HBM-bound (one small gate: 2.61 vs 2.82 ms at n = 30) -- with two or more gates per pass the gate arithmetic
  // saturates the shared-memory and FMA pipes and three independent CTAs per SM hide the fills just as well --
  // so it is used for the exchange-redirect passes (above) and on request.
  const bool use_ring = can_ring && g_tune_ring == 1;
  if (use_ring)
    return dtype == HQ_DTYPE_C64 ? launch_ring_t<float, false>(state, n_qubits, prog, ph, xg, s, grid_override, di, dev)
                                 : launch_ring_t<double, false>(state, n_qubits, prog, ph, xg, s, grid_override, di, dev);
  return launch_tile_pass(dtype, state, n_qubits, prog, ph, stream, grid_override);
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
#define HQ_MARGINAL_SMEM_K 10   // outcomes up to 2^10 are accumulated in a shared-memory histogram first
struct BitsParams {
  unsigned char pos[32];               // index bit of outcome bit j
  unsigned k;
  unsigned long long cond_mask;        // only amplitudes with (i & cond_mask) == cond_value take part
  unsigned long long cond_value;
};

__device__ __forceinline__ unsigned outcome_of(unsigned long long i, const BitsParams& p) {
  unsigned s = 0;
  for (unsigned j = 0; j < p.k; ++j) s |= unsigned((i >> p.pos[j]) & 1ull) << j;
  return s;
}

// out[2 s], out[2 s + 1] += sum over amplitudes with outcome s of re^2, im^2 (global double atomics;
// `out` must be zeroed).  A thread keeps a running sum and flushes it only when its outcome changes:
// SMEM = true (k <= HQ_MARGINAL_SMEM_K) to the CTA's shared-memory histogram, which the CTA flushes once at
// the end; SMEM = false (any k <= 24) straight to the global histogram.
template <typename T, bool SMEM>
__global__ void __launch_bounds__(256) hq_marginal_kernel(const T* __restrict__ state, unsigned long long n_amps,
                                                          const BitsParams p, double* __restrict__ out) {
  extern __shared__ double hist[];
  const unsigned bins = 2u << p.k;
  if (SMEM) {
    for (unsigned b = threadIdx.x; b < bins; b += blockDim.x) hist[b] = 0.0;
    __syncthreads();
  }
  double* acc = SMEM ? hist : out;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  double are = 0, aim = 0;
  unsigned cur = 0;
  bool have = false;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
    if ((i & p.cond_mask) != p.cond_value) continue;
    const unsigned s = outcome_of(i, p);
    if (have && s != cur) {
      atomicAdd(&acc[2 * cur], are);
      atomicAdd(&acc[2 * cur + 1], aim);
      are = aim = 0;
    }
    cur = s;
    have = true;
    const double re = double(state[2 * i]), im = double(state[2 * i + 1]);
    are += re * re;
    aim += im * im;
  }
  if (have) {
    atomicAdd(&acc[2 * cur], are);
    atomicAdd(&acc[2 * cur + 1], aim);
  }
  if (SMEM) {
    __syncthreads();
    for (unsigned b = threadIdx.x; b < bins; b += blockDim.x)
      if (hist[b] != 0.0) atomicAdd(&out[b], hist[b]);
  }
}

int launch_marginal(int dtype, const void* state, unsigned n_qubits, const unsigned* pos, unsigned k,
                    uint64_t cond_mask, uint64_t cond_value, double* out_dev, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (k > HQ_MARGINAL_MAX_K || k > n_qubits) return int(cudaErrorInvalidValue);
  BitsParams p;
  memset(&p, 0, sizeof(p));
  p.k = k;
  p.cond_mask = cond_mask;
  p.cond_value = cond_value;
  for (unsigned j = 0; j < k; ++j) {
    if (pos[j] >= n_qubits) return int(cudaErrorInvalidValue);
    p.pos[j] = (unsigned char)pos[j];
  }
  const unsigned long long n = 1ull << n_qubits;
  if (k <= HQ_MARGINAL_SMEM_K) {
    const size_t smem = (size_t(2) << k) * sizeof(double);
    if (dtype == HQ_DTYPE_C64)
      hq_marginal_kernel<float, true><<<grid_for(n, 256), 256, smem, s>>>((const float*)state, n, p, out_dev);
    else
      hq_marginal_kernel<double, true><<<grid_for(n, 256), 256, smem, s>>>((const double*)state, n, p, out_dev);
  } else {
    if (dtype == HQ_DTYPE_C64)
      hq_marginal_kernel<float, false><<<grid_for(n, 256), 256, 0, s>>>((const float*)state, n, p, out_dev);
    else
      hq_marginal_kernel<double, false><<<grid_for(n, 256), 256, 0, s>>>((const double*)state, n, p, out_dev);
  }
  return int(cudaGetLastError());
}

// amplitudes with (i & mask) != value become 0, the others are scaled plane-wise (any number of bits)
template <typename T>
__global__ void __launch_bounds__(256) hq_project_kernel(T* __restrict__ state, unsigned long long n_amps,
                                                         unsigned long long mask, unsigned long long value,
                                                         T scale_re, T scale_im) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
    const bool keep = (i & mask) == value;
    typename Traits<T>::Cplx v = reinterpret_cast<typename Traits<T>::Cplx*>(state)[i];
    v.x = keep ? v.x * scale_re : T(0);
    v.y = keep ? v.y * scale_im : T(0);
    reinterpret_cast<typename Traits<T>::Cplx*>(state)[i] = v;
  }
}

int launch_project(int dtype, void* state, unsigned n_qubits, uint64_t mask, uint64_t value,
                   double scale_re, double scale_im, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (n_qubits < 64 && ((mask >> n_qubits) || (value & ~mask))) return int(cudaErrorInvalidValue);
  const unsigned long long n = 1ull << n_qubits;
  if (dtype == HQ_DTYPE_C64)
    hq_project_kernel<float><<<grid_for(n, 256), 256, 0, s>>>((float*)state, n, mask, value, float(scale_re), float(scale_im));
  else
    hq_project_kernel<double><<<grid_for(n, 256), 256, 0, s>>>((double*)state, n, mask, value, scale_re, scale_im);
  return int(cudaGetLastError());
}

}  // namespace hq
