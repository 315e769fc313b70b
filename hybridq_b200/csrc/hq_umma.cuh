// hq_umma.cuh -- one dense k = 4, 5 or 6 gate on a complex64 state with the 5th-generation tensor cores:
// `tcgen05.mma kind::tf32` issued by one thread per CTA, accumulators in TMEM, `tcgen05.ld` epilogue.
// (The kernel template also instantiates for k = 3; the library uses it from k = 4, below that the FMA paths are exact
// and as fast.)
//
// This is the "genuine dense contraction" case of the north star (replaces the runtime-k loop of
// /root/reference/include/U.h:123-202): per group of 2^k amplitudes the gate is the real product (k = 5 shown)
//     D[1 x 64] = A[1 x 64] * Bs^T,   A = the group's reals (re, im interleaved), Bs = real form of U,
// and a CTA multiplies 128 groups at a time: M = 128, N = K = 64.  Accuracy: 3xTF32 -- hi * hi + lo * hi + hi * lo with
// hi = the operand rounded to TF32 (nearest) and lo = the rounded remainder; the hi * hi products and the corrections
// go to separate fp32 TMEM accumulators that the epilogue adds (see the kernel for why).
//
// Operand layout (K-major, no swizzle, validated bit-exact by tools/microbench_tcgen05.cu): 16-byte units,
//     A unit (row r, K-chunk c)  at shared slot  c * LBO + r      (LBO = 128 units, or skewed, see below; SBO = 8)
//     B unit (row n, K-chunk c)  at shared slot  c * N   + n      (LBO = N units, SBO = 8 units)
// An A unit holds 4 consecutive reals = the two amplitudes 2c, 2c + 1 of the group, i.e. the pair that differs in
// the gate's LOWEST target bit.  In the plain mapping thread r gathers the units of row r from global memory (for a
// fixed amplitude number the 32 lanes of a warp read consecutive groups: contiguous when the low bits are not
// targets), splits hi / lo in registers and stores both copies with conflict-free 16-byte shared stores; the epilogue
// reads row r of D back from TMEM lane r and scatters its output amplitudes.  With targets among the low bits the
// lanes follow the memory order of the tile instead (MODEB, further down).
//
// Bring-up harness: tools/umma_gate_test.cu; host-side check of the lane map: tools/umma_lane_map_check.cu;
// measurements: profiles/r02/umma_*, sweep_umma_n30.jsonl, ncu_umma_raw.csv; design notes: DESIGN.md section 3.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

// hi * hi accumulators per tile (see the kernel); 3 is what fits next to the correction accumulator for k = 5
#ifndef HQ_UMMA_NACC
#define HQ_UMMA_NACC 3
#endif
// Largest K-chunk stride (LBO of the A descriptor, 16-byte units) the shared-memory buffers are sized for.  128 = dense;
// 129 / 130 / 132 (validated on B200: the no-swizzle layout only needs 16-byte aligned core matrices) skew consecutive
// chunks by 1 / 2 / 4 bank groups so that a quarter-warp covering 8 / 4 / 2 chunks x 1 / 2 / 4 rows stays conflict free.
#define HQ_UMMA_LBO_MAX 132
// timing experiments only (wrong results): 1 = every MMA into accumulator 0, the epilogue unchanged
#ifndef HQ_UMMA_DEBUG_ONE_CHAIN
#define HQ_UMMA_DEBUG_ONE_CHAIN 0
#endif

namespace hq {

struct UmmaPos {
  unsigned char tpos[8];     // ascending amplitude-bit positions of matrix bits 0 .. k-1
  // lane map of the MODEB kernels (umma_lane_map): bit i of cmask set = lane bit i counts K-chunks (else rows);
  // nchunk = popcount(cmask); lbo = K-chunk stride of the A operand in shared memory, in 16-byte units
  unsigned char cmask, nchunk;
  unsigned short lbo;
};

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// shared-memory matrix descriptor: start address, LBO, SBO in 16-byte units; version 1 (Blackwell); no swizzle
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_units, uint32_t sbo_units) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr >> 4) & 0x3fffu);
  d |= uint64_t(lbo_units & 0x3fffu) << 16;
  d |= uint64_t(sbo_units & 0x3fffu) << 32;
  d |= uint64_t(1) << 46;
  return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, dense, M = 128, N
__host__ __device__ inline uint32_t instr_desc(uint32_t n) {
  uint32_t d = 0;
  d |= 1u << 4;             // c_format = F32
  d |= 2u << 7;             // a_format = TF32
  d |= 2u << 10;            // b_format = TF32
  d |= (n >> 3) << 17;      // n_dim
  d |= (128u >> 4) << 24;   // m_dim
  return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u));
}
// spin on an mbarrier phase; traps instead of hanging the GPU if the protocol is broken
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spins > (1u << 26)) __trap();
  }
}
// 32 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

// fp32 -> TF32, round to nearest (ties away): unlike truncation it leaves no bias for the norm to drift on.
// Integer add + mask = what cvt.rna.tf32.f32 does for finite values, in 2 instructions instead of the 4-5 ptxas
// emits for the cvt (inf / nan guard); a state vector holds no inf / nan worth preserving.
__device__ __forceinline__ float tf32_rna(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
// 16 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

template <int CB>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[32]) {
  if (CB == 32) tmem_ld32(taddr, v);
  else tmem_ld16(taddr, v);
}

}  // namespace umma

// TMEM columns of one CTA: up to three hi * hi accumulators and one for the corrections, R columns each (power of two)
template <int KQ>
__host__ __device__ constexpr int umma_tmem_cols() {
  return KQ == 6 ? 512 : (KQ == 5 ? 256 : (KQ == 4 ? 128 : 64));
}
// how many parts ahead the global loads run (register buffers): measured on B200 (profiles/r02/umma_gate_test_*.log),
// two help k = 6 (+12 %, one CTA per SM) and k = 4 (+5 %), not k = 5 (registers: 255 instead of 190 per thread)
template <int KQ, bool MODEB>
__host__ __device__ constexpr int umma_prefetch() {
#ifdef HQ_UMMA_PREFETCH
  return HQ_UMMA_PREFETCH;
#else
  return (KQ == 5 || (KQ == 6 && MODEB)) ? 1 : 2;      // k = 6 with the memory-order lane map would spill with two buffers
#endif
}
// parts the K dimension is processed in (one A buffer in shared memory per part)
template <int KQ>
__host__ __device__ constexpr int umma_ksplit() {
  return KQ == 6 ? 2 : 1;
}

// ---- the MODEB lane map, shared by the kernel and by the host-side exhaustive check (tools/umma_lane_map_check.cu) ----
// lane -> its share of the K-chunk number and of the row number: lane bit b counts chunks if bit b of cmask is set
__host__ __device__ inline void umma_lane_split(unsigned cmask, int lane, int& c_lane, int& r_lane) {
  int ci = 0, ri = 0;
  c_lane = 0;
  r_lane = 0;
  for (int b = 0; b < 5; ++b) {
    if ((cmask >> b) & 1u) c_lane |= ((lane >> b) & 1) << ci++;
    else r_lane |= ((lane >> b) & 1) << ri++;
  }
}
// logical slot (chunk * 128 + row) of the unit a thread handles in iteration i of a sweep over nch chunks x 128 rows;
// A = number of chunk bits among the lane bits.  q = warp * nch + i supplies the remaining 2 + A row bits, then the
// remaining chunk bits.
__host__ __device__ inline int umma_unit_slot(int warp, int i, int nch, int A, int c_lane, int r_lane) {
  const int q = warp * nch + i;
  const int r = ((q & ((4 << A) - 1)) << (5 - A)) | r_lane, c = ((q >> (2 + A)) << A) | c_lane;
  return c * 128 + r;
}

// How a warp's lanes are spread over a tile's 16-byte units (row r, K-chunk c) while it is loaded and stored:
//   MODEB = false: lane = 5 low row bits; thread r owns row r, iteration i handles K-chunk i.  Used when the five lowest
//                  amplitude bits outside the lowest target are all non-targets (consecutive groups are contiguous).
//   MODEB = true:  the lane bits follow the memory order of the tile: lane bit i is the i-th lowest amplitude bit among
//                  row bits and K-chunk bits (targets 1 ..), so that the 8 lanes of a quarter-warp -- the unit in which
//                  128-bit accesses are coalesced -- read one contiguous run whatever the targets are.  The chunk stride
//                  of the A operand is skewed (UmmaPos::lbo) to keep the 16-byte shared accesses of such a quarter-warp
//                  conflict free, and the epilogue is staged through the (free) A buffers to use the same mapping.
// PAIR16: the lowest target is amplitude bit 0, so a unit is 16 contiguous bytes in memory (one 128-bit access).
//
// k = 6 (K = N = 128) does not fit with the whole A tile resident: the K dimension is processed in KSPLIT = 2 halves
// through one A buffer (fill half, MMAs of that half over the matching K-chunks of the resident B, next half), the
// accumulators take all 512 TMEM columns, and one CTA per SM runs.
//
// state: 2^n interleaved complex64 amplitudes; n_tiles = 2^(n - KQ - 7) tiles of 128 groups
template <int KQ, bool MODEB, bool PAIR16>
__global__ void __launch_bounds__(128) hq_umma_gate_kernel(float2* __restrict__ state, const unsigned long long n_tiles,
                                                           const UmmaPos p, const float4* __restrict__ Bhi,
                                                           const float4* __restrict__ Blo) {
  constexpr int DIM = 1 << KQ;       // amplitudes per group
  constexpr int R = 2 * DIM;         // reals per group = N = K
  constexpr int CH = R / 4;          // 16-byte K-chunks per row
  constexpr int KSPLIT = umma_ksplit<KQ>();
  constexpr int CHH = CH / KSPLIT;   // K-chunks of the A buffer
  constexpr int PF = umma_prefetch<KQ, MODEB>();   // parts the loads run ahead
  constexpr int LBO_MAX = HQ_UMMA_LBO_MAX;
  const int LBO = MODEB ? int(p.lbo) : 128;
  auto phys = [&](int slot) { return (slot >> 7) * LBO + (slot & 127); };   // shared index of logical slot chunk * 128 + row
  static_assert(KQ >= 3 && KQ <= 6, "tile and TMEM budget are sized for k = 3 .. 6");
  // The tensor core adds into its fp32 accumulator with truncation (round toward zero); chaining all 3 * R / 8 MMAs
  // through one accumulator shrinks every amplitude by ~7e-7 per gate (measured: norm - 1 = -2.0e-4 after 300 k = 5
  // gates).  So the big hi * hi products go to NACC separate accumulators (short chains, each starting from zero),
  // the small lo * hi and hi * lo corrections to one more, and the epilogue adds them with round-to-nearest FADDs
  // (the scheme of hq_mma.cuh, within the 512 TMEM columns of an SM).
  constexpr int KSTEPS = R / 8;
  constexpr int KSTEPS_H = KSTEPS / KSPLIT;
  constexpr int NACC = KSTEPS < HQ_UMMA_NACC ? KSTEPS : HQ_UMMA_NACC;
  constexpr int COLS = umma_tmem_cols<KQ>();
  static_assert((NACC + 1) * R <= COLS, "TMEM columns");
  extern __shared__ __align__(128) unsigned char smem[];
  float4* const sAhi = reinterpret_cast<float4*>(smem);
  float4* const sAlo = sAhi + CHH * LBO_MAX;       // contiguous with sAhi: together they stage the MODEB epilogue
  float4* const sBhi = sAlo + CHH * LBO_MAX;
  float4* const sBlo = sBhi + CH * R;
  unsigned long long* const dep = reinterpret_cast<unsigned long long*>(sBlo + CH * R);   // dep[j]: offset of amplitude j
  unsigned long long* const rowoff = dep + DIM;                                            // rowoff[r]: offset of row r
  unsigned long long* const bar = rowoff + 128;
  uint32_t* const tmem_holder = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // group index -> amplitude index with zeros at the target bits (a bit permutation: OR-separable)
  auto spread = [&](unsigned long long g) {
#pragma unroll
    for (int b = 0; b < KQ; ++b) {
      const unsigned long long low = (1ull << p.tpos[b]) - 1ull;
      g = ((g & ~low) << 1) | (g & low);
    }
    return g;
  };
  for (int i = tid; i < CH * R; i += 128) {
    sBhi[i] = Bhi[i];
    sBlo[i] = Blo[i];
  }
  if (tid < DIM) {
    unsigned long long d = 0;
    for (int b = 0; b < KQ; ++b) d |= (unsigned long long)((tid >> b) & 1) << p.tpos[b];
    dep[tid] = d;
  }
  rowoff[tid] = spread((unsigned long long)tid);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(umma::smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(umma::smem_u32(tmem_holder)),
                 "r"(uint32_t(COLS)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tmem_holder;
  const uint32_t idesc = umma::instr_desc(R);
  const uint32_t a_hi = umma::smem_u32(sAhi), a_lo = umma::smem_u32(sAlo), b_hi = umma::smem_u32(sBhi), b_lo = umma::smem_u32(sBlo);
  const unsigned long long pair_bit = 1ull << p.tpos[0];

  // unit handled by this thread in iteration i of a sweep over NCH K-chunks x 128 rows: shared slot = chunk * 128 + row
  // MODEB: this lane's share of the chunk and row numbers (lane bits compacted by cmask), the rest comes from
  // q = warp * nch + i: row bits first (2 + A of them), then chunk bits
  const int A = MODEB ? int(p.nchunk) : 0;
  int c_lane = 0, r_lane = 0;
  if (MODEB) umma_lane_split(p.cmask, lane, c_lane, r_lane);
  auto unit_slot = [&](int i, int nch) {
    if (!MODEB) return i * 128 + tid;
    return umma_unit_slot(warp, i, nch, A, c_lane, r_lane);
  };
  // amplitude offset inside the tile of the unit in `slot`, K-chunks counted from chunk0
  auto unit_off = [&](int slot, int chunk0) { return rowoff[slot & 127] | dep[2 * (chunk0 + (slot >> 7))]; };
  // The units of one K-part of a tile live in registers: all loads of a part are in flight at once.  With PF = 2 two
  // register buffers alternate, so that while part s is converted, multiplied and written back, the loads of parts
  // s + 1 and s + 2 are outstanding (those of s + 2 are issued into the buffer of s as soon as it has been stored to
  // shared memory): up to 2 x 32 KiB per CTA in flight.  With PF = 1 there is one buffer and one part in flight.  Part s of this CTA = K-part s % KSPLIT of tile blockIdx.x + (s / KSPLIT) * gridDim.x.
  auto part_tile = [&](unsigned long long sidx) { return (unsigned long long)blockIdx.x + (sidx / KSPLIT) * gridDim.x; };
  auto load_part = [&](float4 (&buf)[CHH], unsigned long long sidx) {
    const unsigned long long t = part_tile(sidx);
    if (t >= n_tiles) return;
    const unsigned long long tb = spread(t * 128ull);
    const int h = int(sidx % KSPLIT);
#pragma unroll
    for (int i = 0; i < CHH; ++i) {
      const unsigned long long a = tb | unit_off(unit_slot(i, CHH), h * CHH);
      if (PAIR16) {
        buf[i] = *reinterpret_cast<const float4*>(&state[a]);
      } else {
        const float2 x0 = state[a], x1 = state[a | pair_bit];
        buf[i] = make_float4(x0.x, x0.y, x1.x, x1.y);
      }
    }
  };
  auto store_unit = [&](unsigned long long a, const float4 v) {
    if (PAIR16) {
      *reinterpret_cast<float4*>(&state[a]) = v;
    } else {
      state[a] = make_float2(v.x, v.y);
      state[a | pair_bit] = make_float2(v.z, v.w);
    }
  };
  uint32_t phase = 0;
  auto step = [&](float4 (&x)[CHH], unsigned long long sidx) {
    const int h = int(sidx % KSPLIT);
    // split hi / lo, store both copies in the canonical layout
#pragma unroll
    for (int i = 0; i < CHH; ++i) {
      float4 hi, lo;
      hi.x = umma::tf32_rna(x[i].x);
      hi.y = umma::tf32_rna(x[i].y);
      hi.z = umma::tf32_rna(x[i].z);
      hi.w = umma::tf32_rna(x[i].w);
      lo.x = umma::tf32_rna(x[i].x - hi.x);
      lo.y = umma::tf32_rna(x[i].y - hi.y);
      lo.z = umma::tf32_rna(x[i].z - hi.z);
      lo.w = umma::tf32_rna(x[i].w - hi.w);
      const int slot = unit_slot(i, CHH);
      sAhi[phys(slot)] = hi;
      sAlo[phys(slot)] = lo;
    }
    // generic-proxy writes of the operands must be visible to the async proxy the tensor core reads through
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      // hi * hi: K-step kg goes to accumulator kg * NACC / KSTEPS (a chain restarts from zero when the index changes);
      // lo * hi and hi * lo: one chain in accumulator NACC, interleaved with the hi * hi MMAs so that the tensor pipe
      // always has an independent MMA to run while a chained one waits for its accumulator
#pragma unroll 1
      for (int ks = 0; ks < KSTEPS_H; ++ks) {
        const int kg = h * KSTEPS_H + ks;
        const uint32_t a_off = uint32_t(ks) * 2u * uint32_t(LBO) * 16u, b_off = uint32_t(kg) * 2u * uint32_t(R) * 16u;
        const uint64_t da_hi = umma::smem_desc(a_hi + a_off, LBO, 8), da_lo = umma::smem_desc(a_lo + a_off, LBO, 8);
        const uint64_t db_hi = umma::smem_desc(b_hi + b_off, R, 8), db_lo = umma::smem_desc(b_lo + b_off, R, 8);
        const int acc = kg * NACC / KSTEPS;
        const bool first = kg == 0 || (kg - 1) * NACC / KSTEPS != acc;
        if (HQ_UMMA_DEBUG_ONE_CHAIN) {
          umma::mma_tf32(tmem, da_hi, db_hi, idesc, kg ? 1u : 0u);
          umma::mma_tf32(tmem, da_lo, db_hi, idesc, 1u);
          umma::mma_tf32(tmem, da_hi, db_lo, idesc, 1u);
        } else {
          umma::mma_tf32(tmem + uint32_t(acc * R), da_hi, db_hi, idesc, first ? 0u : 1u);
          umma::mma_tf32(tmem + uint32_t(NACC * R), da_lo, db_hi, idesc, kg ? 1u : 0u);
          umma::mma_tf32(tmem + uint32_t(NACC * R), da_hi, db_lo, idesc, 1u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(umma::smem_u32(bar))
                   : "memory");
    }
    load_part(x, sidx + PF);                           // the registers of this part are free: fetch a later part into them
    umma::mbar_wait(umma::smem_u32(bar), phase);      // this part's MMAs are done: the A buffer is free again
    phase ^= 1u;
    if (h + 1 < KSPLIT) return;
    const unsigned long long tbase = spread(part_tile(sidx) * 128ull);
    asm volatile("tcgen05.fence::after_thread_sync;");
    // epilogue: row `tid` of D = TMEM lane tid (warp w owns lanes 32 w .. 32 w + 31), up to 32 columns per load
    constexpr int CB = R < 32 ? R : 32;
#pragma unroll 1
    for (int c0 = 0; c0 < R; c0 += CB) {
      uint32_t v[32], w[32];
      const uint32_t lane_base = tmem + (uint32_t(warp * 32) << 16) + uint32_t(c0);
      umma::tmem_ld<CB>(lane_base, v);
#pragma unroll
      for (int a = 1; a <= NACC; ++a) {
        umma::tmem_ld<CB>(lane_base + uint32_t(a * R), w);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int e = 0; e < CB; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(w[e]));
      }
#pragma unroll
      for (int u = 0; u < CB / 4; ++u) {
        const float4 d = make_float4(__uint_as_float(v[4 * u]), __uint_as_float(v[4 * u + 1]), __uint_as_float(v[4 * u + 2]),
                                     __uint_as_float(v[4 * u + 3]));
        const int c = c0 / 4 + u;
        if (MODEB) {
          sAhi[c * LBO + tid] = d;          // CH * LBO units = the A_hi and A_lo buffers together
        } else {
          store_unit(tbase | rowoff[tid] | dep[2 * c], d);
        }
      }
    }
    if (MODEB) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int slot = unit_slot(i, CH);
        store_unit(tbase | unit_off(slot, 0), sAhi[phys(slot)]);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();       // the operands and the accumulator may be overwritten
  };
  if (PF == 2) {
    float4 xa[CHH], xb[CHH];
    load_part(xa, 0);
    load_part(xb, 1);
#pragma unroll 1
    for (unsigned long long sidx = 0;; sidx += 2) {
      if (part_tile(sidx) >= n_tiles) break;
      step(xa, sidx);
      if (part_tile(sidx + 1) >= n_tiles) break;
      step(xb, sidx + 1);
    }
  } else {
    float4 xa[CHH];
    load_part(xa, 0);
#pragma unroll 1
    for (unsigned long long sidx = 0; part_tile(sidx) < n_tiles; ++sidx) step(xa, sidx);
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(uint32_t(COLS)));
}

template <int KQ>
inline size_t umma_smem_bytes() {
  constexpr int DIM = 1 << KQ, R = 2 * DIM, CH = R / 4, CHH = CH / umma_ksplit<KQ>();
  return size_t(2 * CHH * HQ_UMMA_LBO_MAX + 2 * CH * R) * 16 + size_t(DIM + 128) * 8 + 8 + 16;
}

// grid of the most recent launch (diagnostics)
inline unsigned& umma_last_grid() {
  static unsigned g = 0;
  return g;
}

// Lane map in memory order (see the kernel): scan the amplitude bits upwards, skip the lowest target (it lives inside a
// unit), and give the next five bits to the lanes -- a target is the next K-chunk bit (at most max_chunk_bits of them),
// anything else the next row bit.  The quarter-warp (lane bits 0 .. 2) then spans 2^a chunks x 2^(3-a) rows, which are
// conflict free in shared memory when consecutive chunks are 2^(3-a) bank groups apart: lbo = 128 + 2^(3-a) (128 if a = 0).
inline void umma_lane_map(UmmaPos& p, int k, int max_chunk_bits) {
  int lane_i = 0, chunks = 0, quarter_chunks = 0, t = 1;
  p.cmask = 0;
  for (int b = 0; lane_i < 5; ++b) {
    if (b == int(p.tpos[0])) continue;
    if (t < k && b == int(p.tpos[t])) {
      ++t;
      if (chunks >= max_chunk_bits) continue;      // a higher chunk bit: left to the iterations
      p.cmask |= (unsigned char)(1u << lane_i);
      ++chunks;
      if (lane_i < 3) ++quarter_chunks;
    }
    ++lane_i;
  }
  p.nchunk = (unsigned char)chunks;
  p.lbo = (unsigned short)(quarter_chunks ? 128 + (8 >> quarter_chunks) : 128);
}

// One dense k = KQ gate on the whole state (n >= KQ + 7).  Bhi / Blo: device pointers to the real form of U split
// into TF32 hi / lo parts, in canonical units (unit (n, c) at index c * R + n holds Bs[n][4c .. 4c + 3]).
// mode: -1 = lane mapping from the target positions, 0 = rows only, 1 = memory-order map even when it equals rows.
template <int KQ>
inline int launch_umma_gate(float2* state, unsigned n_qubits, const UmmaPos& p_in, const float4* Bhi, const float4* Blo,
                            cudaStream_t stream, int mode = -1, int ctas_per_sm = 0) {
  if (n_qubits < unsigned(KQ) + 7u) return int(cudaErrorInvalidValue);
  const size_t smem = umma_smem_bytes<KQ>();
  const bool pair16 = p_in.tpos[0] == 0;
  constexpr int CHH_BITS = KQ - 1 - (umma_ksplit<KQ>() == 2 ? 1 : 0);      // log2 of the K-chunks in the A buffer
  UmmaPos p = p_in;
  umma_lane_map(p, KQ, CHH_BITS);
  if (mode < 0) {
    // 128-byte lines one warp-wide access touches (16 amplitudes a line) with the lanes on the five lowest row bits
    // versus on the five lowest bits in memory order; the memory-order kernels cost ~10 % (staged epilogue, registers;
    // k = 6: one register buffer instead of two), so they must save at least half (k = 6: three quarters) of the lines
    int lines_rows = 1, lines_mem = 1, seen = 0, rows_seen = 0;
    for (int b = 0, t = 0; rows_seen < 5; ++b) {
      const bool is_target = t < KQ && b == int(p.tpos[t]);
      if (is_target) ++t;
      if (!is_target) {
        ++rows_seen;
        if (b >= 4) lines_rows *= 2;
      }
      if (b != int(p.tpos[0]) && seen < 5) {
        ++seen;
        if (b >= 4) lines_mem *= 2;
      }
    }
    mode = (p.nchunk && lines_mem * (KQ == 6 ? 4 : 2) <= lines_rows) ? 1 : 0;
  }
  if (mode == 0) {          // rows only (forced, or nothing to gain)
    p.cmask = 0;
    p.nchunk = 0;
    p.lbo = 128;
  }
  void (*kern)(float2*, unsigned long long, UmmaPos, const float4*, const float4*) =
      mode ? (pair16 ? hq_umma_gate_kernel<KQ, true, true> : hq_umma_gate_kernel<KQ, true, false>)
           : (pair16 ? hq_umma_gate_kernel<KQ, false, true> : hq_umma_gate_kernel<KQ, false, false>);
  // per variant: opt in to the shared-memory size once, and ask the runtime how many CTAs really fit on an SM
  // (registers included) so that the persistent grid is exactly one wave; TMEM columns bound it as well
  static int resident_all[64][4] = {};      // per device: the shared-memory opt-in is a per-device attribute
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return int(cudaErrorInvalidDevice);
  int* const resident = resident_all[dev];
  const int which = (mode ? 2 : 0) + (pair16 ? 1 : 0);
  if (!resident[which]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return int(e);
    // CTAs that fit on an SM: registers (allocated per warp in units of 8 per thread), shared memory (227 KiB usable,
    // 1 KiB reserved per CTA), TMEM columns
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return int(e);
    int occ = 65536 / (((fa.numRegs + 7) & ~7) * 128);
    const int by_smem = int((227u * 1024u) / (smem + 1024u));
    if (occ > by_smem) occ = by_smem;
    if (occ > 512 / umma_tmem_cols<KQ>()) occ = 512 / umma_tmem_cols<KQ>();
    if (occ < 1) return int(cudaErrorLaunchOutOfResources);
    resident[which] = occ;
  }
  const unsigned long long n_tiles = 1ull << (n_qubits - unsigned(KQ) - 7u);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (ctas_per_sm <= 0 || ctas_per_sm > resident[which]) ctas_per_sm = resident[which];
  unsigned long long grid = (unsigned long long)sms * (unsigned long long)ctas_per_sm;
  if (grid > n_tiles) grid = n_tiles;
  umma_last_grid() = unsigned(grid);
  kern<<<unsigned(grid), 128, smem, stream>>>(state, n_tiles, p, Bhi, Blo);
  return int(cudaGetLastError());
}

}  // namespace hq
