// hq_common.h -- structures shared by the host planner (hq_plan.cpp), the CUDA kernels
// (hq_kernels.cu) and the CPU emulation of the kernel phases used by the unit tests
// (hq_emu.cpp).  Plain C++14, no CUDA types.
//
// Vocabulary (used everywhere in csrc/):
//   amplitude      one complex number of the state vector (complex64 = 8 B, complex128 = 16 B)
//   unit           16 bytes of state: 2 amplitudes (complex64, V = 1) or 1 (complex128, V = 0)
//   tile           the 2^T amplitudes one CTA holds in shared memory: the low L index bits
//                  (one contiguous "run" of 2^L amplitudes) plus h = T - L arbitrary "high"
//                  index bits.  A tile is closed under every gate whose targets are tile bits.
//   pass           one sweep over the whole state (one kernel launch): every tile is loaded
//                  once, all gates of the pass are applied in shared memory, and the tile is
//                  written back once.  Algorithmic HBM bytes per pass = 2 * 2^n * sizeof(amp).
//   gate-apply     one dense 2^k x 2^k matrix applied to the state (the reference's
//                  apply_U call, /root/reference/include/python_U.cpp:131-143).
#pragma once
#include <cstdint>

#define HQ_DTYPE_C64 0
#define HQ_DTYPE_C128 1

#define HQ_MAX_UNIT_BITS 12   // 2^12 units * 16 B = 64 KiB of shared memory per tile
#define HQ_MAX_HIGH 10        // at most 2^10 runs per tile
#define HQ_MAX_K 10           // largest gate handled by the tile kernels
#define HQ_SMALL_K 4          // k <= 4: register path; k >= 5: two-phase path
#define HQ_THREADS_LOG2 8
#define HQ_THREADS (1 << HQ_THREADS_LOG2)
#define HQ_BIG_ROWS 8         // output rows per thread in the two-phase (k >= 5) path

#define HQ_GATE_SMALL 0
#define HQ_GATE_BIG 1
#define HQ_GATE_ROWPAIR 2     // complex128 k = 2, 3: row-pair scheme (see HqGateDesc::tbl_rthread)
#define HQ_GATE_MMA 3         // tensor-core path (hq_mma.cuh): mma.sync 3xTF32 / FP64, 2 <= k <= HQ_MMA_MAX_K
#define HQ_MMA_MAX_K 6
#define HQ_GATE_DR1 4         // "scalar + rank one": U = lambda * 1 + u v^T, 3 <= k <= HQ_DR1_MAX_K (hq_tile.cuh gate_dr1):
                              // what a depolarizing channel is as a super-operator (hybridq/noise/channel/channel.py:413-529
                              // build the dense 4^k x 4^k matrix; here it costs 2 * 2^k MACs per group instead of 4^k)
#define HQ_DR1_MAX_K 4

#define HQ_MAX_PER_THREAD 16   // units per thread per tile = 2^(unit bits - 8) <= 16
#define HQ_MAX_PASS_GATES 24   // kernel matrices per pass after merging
#define HQ_FAST_SLOTS 8        // constant-bank gate slots per pass (complex64, k <= 3); 12 measured 6 % slower on the
                               // bench circuit (code size / register allocation of the slot switch), profiles/r02
#define HQ_FAST_MAX_K 3

struct HqGateDesc {        // 1280 bytes, lives in the device program buffer (read through L1)
  uint32_t k;              // number of target bits
  uint32_t kind;           // HQ_GATE_SMALL / HQ_GATE_BIG / HQ_GATE_ROWPAIR
  uint32_t mat_off;        // byte offset of the matrix from the program base
                           //   small: row-major 2^k x 2^k, interleaved (re, im)
                           //   big  : column-major (transposed), interleaved
                           //   mma  : per-lane B fragments, 16 bytes each: [(s * KS + j) * 32 + lane]
                           //   dr1  : lambda, u[2^k], v[2^k], interleaved (re, im)
  uint32_t n_free;         // number of entries in q[]
  uint8_t tpos[16];        // ascending LOCAL amplitude-bit positions of matrix bits 0..k-1
  uint8_t q[16];           // small: ordering of the non-target local UNIT bits (work-item bit b
                           //        -> unit bit q[b]); big: non-target local AMPLITUDE bits
  // Register-path lane tables, precomputed by the planner so that the kernel does no bit
  // scattering at all.  Work item w = tid + (it << 8) of a gate touches the units
  //   slot(w, m) = tbl_thread[tid] ^ tbl_iter[it] ^ tbl_x[m],   m = 0 .. 2^KK - 1
  // (already swizzled shared-memory slots; swz is GF(2)-linear so the XOR composes).
  // HQ_GATE_MMA reuses the three tables with its own lane mapping (lane = 4 g + t, see hq_mma.cuh):
  //   slot(row set it, row g, m) = tbl_thread[tid] ^ tbl_iter[it] ^ tbl_x[m],  m = t + 4 s
  // in units (unit path) or in float2 amplitudes (complex64 amplitude path, mma_amp = 1).
  uint16_t tbl_thread[HQ_THREADS];
  uint16_t tbl_iter[16];
  uint16_t tbl_x[64];
  // complex128 "row-pair" scheme (k = 2, 3): thread tid = (group slot gs, row pair rp) with
  // rp = tid & (2^(k-1) - 1), gs = tid >> (k-1); the units of its group are
  //   slot(gs, it, m) = tbl_rthread[gs] ^ tbl_riter[it] ^ tbl_x[m]
  // and it produces rows 2rp, 2rp+1 of the group with those two matrix rows held in registers.
  uint16_t tbl_rthread[HQ_THREADS];
  uint16_t tbl_riter[16];
  // HQ_GATE_MMA
  uint32_t mma_n_iter;     // row sets per warp
  uint32_t mma_warps;      // warps that have work (all 8 unless the tile is small)
  uint32_t mma_amp;        // complex64: 1 = amplitude granularity (amplitude bit 0 is a target)
  uint32_t mma_row8;       // amplitude path: XOR offset of row g + 8
};

struct HqPassHeader {      // passed to the kernel by value (constant bank)
  uint32_t n_gates;
  uint32_t tile_bits;      // T, in amplitudes
  uint32_t n_high;         // h
  uint32_t gates_off;      // byte offset of HqGateDesc[0] of this pass from the program base
  uint8_t high_pos[16];    // ascending GLOBAL amplitude-bit positions of tile bits L..T-1
  // optional local bit permutation applied when the tile is written back (in-place index-bit
  // swap kernel, replaces /root/reference/include/swap.h): out_local[j] = tile[sigma(j)],
  // sigma(j) = XOR_i bit_i(j) << perm[i].  has_perm = 0 -> identity.
  uint32_t has_perm;
  uint8_t perm[16];
  uint32_t max_k;          // largest k among the gates of the pass (selects the kernel variant)
  uint32_t reserved;
  // Fill/drain addressing: local unit c = tid + (i << 8) lives at global unit
  //   tile_base + off(tid) + iter_off[i]   and at shared-memory slot  swz(tid) ^ iter_swz[i]
  // (off() deposits the bits of c at their global positions, so it splits over disjoint bits).
  uint64_t iter_off[HQ_MAX_PER_THREAD];
  uint32_t iter_swz[HQ_MAX_PER_THREAD];
  // Fast slots (complex64): kernel matrix s < HQ_FAST_SLOTS of the pass with k <= 3 on the FMA path has bit s of
  // fast_mask set, fast_k[s] = k | (4 if matrix bit 0 is amplitude bit 0) and its 2^k x 2^k matrix (row-major,
  // re/im) in fast_u[s].  The header is a kernel parameter, so every matrix element is a constant-bank /
  // uniform-register operand of an FFMA2: on B200 a 3-register FFMA issues at 21 TFMA/s, one with a
  // constant-bank operand at 35 TFMA/s (profiles/r01/microbench_fma.jsonl).
  uint32_t fast_mask;
  // bit s set: slot gates s and s + 1 belong to one warp-closed chain (hq_plan.cpp) -- every warp touches the same
  // set of units in both, so only __syncwarp() separates them, not a CTA barrier
  uint32_t chain_mask;
  uint8_t fast_k[HQ_FAST_SLOTS];
  float fast_u[HQ_FAST_SLOTS][2 << (2 * HQ_FAST_MAX_K)];
};

static_assert(sizeof(HqGateDesc) == 624 + 96 + 512 + 32 + 16, "HqGateDesc layout");
static_assert(sizeof(HqPassHeader) == 256 + 8 + ((HQ_FAST_SLOTS + 7) & ~7) + 4 * 128 * HQ_FAST_SLOTS, "HqPassHeader layout");
static_assert(sizeof(HqPassHeader) + 256 <= 32764, "the pass header travels as a kernel parameter (large-parameter limit)");
