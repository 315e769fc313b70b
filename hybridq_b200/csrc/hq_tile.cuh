// hq_tile.cuh -- the phases of the tile kernel, written once and compiled twice:
//   * by nvcc for sm_100a inside hq_kernels.cu (the product), and
//   * by g++ inside hq_emu.cpp, where a host loop plays the role of the threads of one CTA
//     (test infrastructure only; lets the CPU-only test-suite check the index math).
//
// What the kernel computes is the reference's U::apply (/root/reference/include/U.h:28-102,
// :123-202): for every group of 2^k amplitudes that differ only in the k target bits,
// psi' = U psi, in place.  HOW it is computed is B200-first and shares nothing with the
// reference: the state is interleaved complex in HBM, a CTA stages a tile of 2^T amplitudes
// in shared memory with 16-byte cp.async copies of whole contiguous runs (so global traffic
// is fully coalesced whatever the target bits are), applies every gate of the pass to the
// tile from registers, and writes the tile back once.
//
// Shared-memory layout: the tile is an array of 16-byte units indexed by the local unit
// index u; unit u lives at physical slot swz(u) = u ^ f(u >> 3), where f is a GF(2)-linear map of
// unit bits 3..11 onto slot bits 0..2 (so swz(a ^ b) = swz(a) ^ swz(b) and table entries compose
// by XOR).  Every unit bit b therefore has a 3-bit "bank vector" swz_vec(b) -- 1, 2, 4 for bits
// 0..2 -- and a quarter-warp of 16-byte accesses whose lanes differ in three unit bits is
// conflict-free exactly when the three vectors are linearly independent.  The vectors are
//   bit      0 1 2 3 4 5 6 7 8 9 10 11
//   vector   1 2 4 3 6 5 7 1 2 4  3  6
// i.e. all seven non-zero vectors are used and only 5 of the 66 bit pairs share one (the earlier
// u ^ (u>>3 & 7) ^ (u>>6 & 7) ^ (u>>9 & 7) used three vectors, 18 colliding pairs, which cost the
// tensor-core k = 2 gates a 2-way conflict on 27 % of the target pairs: ncu, profiles/r01).
// Fills and drains touch 8 consecutive units per quarter-warp (vectors 1, 2, 4): conflict-free;
// gate work items are mapped to lanes by the host (HqGateDesc::q, mma_layout) so that the
// lane bits of a quarter-warp land on independent vectors whenever the gate's bits allow it.
#pragma once
#include "hq_common.h"

#ifdef __CUDACC__
#define HQ_DEV __device__ __forceinline__
#define HQ_HD __host__ __device__ __forceinline__
#define HQ_LDG(p) __ldg(p)
#define HQ_UNROLL _Pragma("unroll")
#define HQ_NOUNROLL _Pragma("unroll 1")
#define HQ_ROWUNROLL _Pragma("unroll 1")
#else
#define HQ_DEV inline
#define HQ_HD inline
#define HQ_LDG(p) (*(p))
#define HQ_UNROLL
#define HQ_NOUNROLL
#define HQ_ROWUNROLL
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline double2 make_double2(double a, double b) { return double2{a, b}; }
#endif

namespace hq {

template <typename T> struct Traits;
template <> struct Traits<float> {
  typedef float4 Unit;      // 2 amplitudes: (x, y) = even amplitude re/im, (z, w) = odd
  typedef float2 Cplx;
  static const int V = 1;   // log2(amplitudes per unit)
};
template <> struct Traits<double> {
  typedef double2 Unit;     // 1 amplitude
  typedef double2 Cplx;
  static const int V = 0;
};

// bank vector of unit bit b (b < 12)
HQ_HD uint32_t swz_vec(int b) {
  // 3 bits per unit bit, bit 0 first: 1 2 4 3 6 5 7 1 2 4 3 6
  const uint64_t packed = 01ull | (02ull << 3) | (04ull << 6) | (03ull << 9) | (06ull << 12) | (05ull << 15) |
                          (07ull << 18) | (01ull << 21) | (02ull << 24) | (04ull << 27) | (03ull << 30) | (06ull << 33);
  return uint32_t(packed >> (3 * b)) & 7u;
}

// f restricted to one 3-bit group of unit bits (first = 3, 6 or 9), as an 8-entry table packed in 24 bits
HQ_HD uint32_t swz_group_table(int first) {
  uint32_t t = 0;
  for (uint32_t g = 0; g < 8; ++g) {
    uint32_t v = 0;
    for (int i = 0; i < 3; ++i)
      if ((g >> i) & 1u) v ^= swz_vec(first + i);
    t |= v << (3 * g);
  }
  return t;
}

HQ_HD uint32_t swz(uint32_t u) {
  const uint32_t t1 = swz_group_table(3), t2 = swz_group_table(6), t3 = swz_group_table(9);   // folded to constants
  return u ^ ((t1 >> (3 * ((u >> 3) & 7u))) & 7u) ^ ((t2 >> (3 * ((u >> 6) & 7u))) & 7u) ^
         ((t3 >> (3 * ((u >> 9) & 7u))) & 7u);
}

// Open a zero gap at every position of `pos` (ascending, final coordinates).
HQ_HD uint64_t open_gaps(uint64_t x, const uint8_t* pos, int n) {
  for (int i = 0; i < n; ++i) {
    const uint64_t low = (uint64_t(1) << pos[i]) - 1;
    x = ((x & ~low) << 1) | (x & low);
  }
  return x;
}

HQ_HD uint64_t deposit(uint32_t m, const uint8_t* pos, int n) {
  uint64_t y = 0;
  for (int i = 0; i < n; ++i) y |= uint64_t((m >> i) & 1u) << pos[i];
  return y;
}

// First amplitude index of tile `t`.
HQ_HD uint64_t tile_base(uint64_t t, int tile_bits, int n_high, const uint8_t* high_pos) {
  return open_gaps(t << (tile_bits - n_high), high_pos, n_high);
}

// ---------------------------------------------------------------------------------------
// complex helpers
// ---------------------------------------------------------------------------------------
HQ_DEV float hq_fma(float a, float b, float c) {
#ifdef __CUDACC__
  return __fmaf_rn(a, b, c);
#else
  return a * b + c;
#endif
}
HQ_DEV double hq_fma(double a, double b, double c) {
#ifdef __CUDACC__
  return __fma_rn(a, b, c);
#else
  return a * b + c;
#endif
}

// (ar, ai) += (ur, ui) * (xr, xi) as exactly four FMAs (the arithmetic of U.h:93-94)
template <typename R>
HQ_DEV void cmac(R& ar, R& ai, R ur, R ui, R xr, R xi) {
  ar = hq_fma(ur, xr, ar);
  ar = hq_fma(-ui, xi, ar);
  ai = hq_fma(ur, xi, ai);
  ai = hq_fma(ui, xr, ai);
}

// Work item w (bits [0, n_free)) -> local unit/amp index with zeros at the target bits
// (used by the planner for the lane tables and by the two-phase path).
HQ_HD uint32_t scatter_bits(uint32_t w, const uint8_t* q, int from, int to) {
  uint32_t u = 0;
  for (int b = from; b < to; ++b) u |= ((w >> (b - from)) & 1u) << q[b];
  return u;
}

// Iteration offsets without loads: every lane table is GF(2)-linear in its index, so
//   tbl_iter[it] = XOR over the set bits b of it of tbl_iter[1 << b].
// The four basis entries are read once per gate; inside the loops the offset is a handful of
// (warp-uniform) logic operations instead of a dependent L1 load in front of every shared load.
struct IterBasis { uint32_t b[4]; };
HQ_DEV IterBasis load_iter_basis(const uint16_t* __restrict__ tbl) {
  IterBasis r;
  HQ_UNROLL
  for (int i = 0; i < 4; ++i) r.b[i] = HQ_LDG(&tbl[1 << i]);
  return r;
}
HQ_DEV uint32_t iter_offset(const IterBasis& r, uint32_t it) {
  return ((it & 1u) ? r.b[0] : 0u) ^ ((it & 2u) ? r.b[1] : 0u) ^ ((it & 4u) ? r.b[2] : 0u) ^ ((it & 8u) ? r.b[3] : 0u);
}

// ---------------------------------------------------------------------------------------
// register path: the slot of unit m of work item w = tid + (it << 8) is
//   tbl_thread[tid] ^ tbl_iter[it] ^ tbl_x[m]      (see HqGateDesc)
// GateIO loads / stores the units of one work item.
// ---------------------------------------------------------------------------------------

// complex64, no target on amplitude bit 0: KK unit-level target bits, every thread handles two
// groups at once (the even and the odd amplitude of its units).
template <int KK>
HQ_DEV void gate_small_f32(float4* tile, const HqGateDesc* __restrict__ g, const float2* __restrict__ U,
                           int Tu, int tid) {
  const int DIM = 1 << KK;
  const int nq = Tu - KK;
  const uint32_t nwork = 1u << nq;
  if (uint32_t(tid) >= nwork) return;
  const uint32_t niter = nq > HQ_THREADS_LOG2 ? (nwork >> HQ_THREADS_LOG2) : 1u;
  const uint32_t st = HQ_LDG(&g->tbl_thread[tid]);
  const IterBasis ib = load_iter_basis(g->tbl_iter);
  uint32_t xo[DIM];
  HQ_UNROLL
  for (int m = 0; m < DIM; ++m) xo[m] = HQ_LDG(&g->tbl_x[m]);

  float2 Ur[KK <= 2 ? DIM * DIM : 1];
  if (KK <= 2) {
    HQ_UNROLL
    for (int e = 0; e < DIM * DIM; ++e) Ur[e] = HQ_LDG(&U[e]);
  }

  HQ_NOUNROLL
  for (uint32_t it = 0; it < niter; ++it) {
    const uint32_t sb = st ^ iter_offset(ib, it);
    float4 in[DIM];
    HQ_UNROLL
    for (int m = 0; m < DIM; ++m) in[m] = tile[sb ^ xo[m]];
    if (KK <= 2) {
      HQ_UNROLL
      for (int i = 0; i < DIM; ++i) {
        float a0r = 0.f, a0i = 0.f, a1r = 0.f, a1i = 0.f;
        HQ_UNROLL
        for (int j = 0; j < DIM; ++j) {
          const float2 u = Ur[i * DIM + j];
          cmac(a0r, a0i, u.x, u.y, in[j].x, in[j].y);
          cmac(a1r, a1i, u.x, u.y, in[j].z, in[j].w);
        }
        tile[sb ^ xo[i]] = make_float4(a0r, a0i, a1r, a1i);
      }
    } else {
      HQ_ROWUNROLL
      for (int i = 0; i < DIM; ++i) {
        float a0r = 0.f, a0i = 0.f, a1r = 0.f, a1i = 0.f;
        const float4* row = reinterpret_cast<const float4*>(U + i * DIM);
        HQ_UNROLL
        for (int j = 0; j < DIM; j += 2) {
          const float4 u = HQ_LDG(&row[j >> 1]);
          cmac(a0r, a0i, u.x, u.y, in[j].x, in[j].y);
          cmac(a1r, a1i, u.x, u.y, in[j].z, in[j].w);
          cmac(a0r, a0i, u.z, u.w, in[j + 1].x, in[j + 1].y);
          cmac(a1r, a1i, u.z, u.w, in[j + 1].z, in[j + 1].w);
        }
        tile[sb ^ uint32_t(HQ_LDG(&g->tbl_x[i]))] = make_float4(a0r, a0i, a1r, a1i);
      }
    }
  }
}

// Fast slot: the same arithmetic as gate_small_f32<2>, with the 4x4 matrix taken from the pass
// header (kernel parameter = constant bank) at compile-time offsets, so that every FFMA has
// only two register operands and issues at full rate.
// acc(lo, hi) += x(lo, hi) * s  -- one FFMA2 with the scalar broadcast from a uniform register
// (SASS: FFMA2 R, R.F32x2, UR.F32, R.F32x2).
struct F2 { float lo, hi; };
HQ_DEV void ffma2_bcast(F2& acc, const F2& x, float s) {
#ifdef __CUDACC__
  unsigned long long a, b, c;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(x.lo), "f"(x.hi));
  asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(s));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc.lo), "f"(acc.hi));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.lo), "=f"(acc.hi) : "l"(c));
#else
  acc.lo = x.lo * s + acc.lo;
  acc.hi = x.hi * s + acc.hi;
#endif
}

// Streaming FFMA2 gate (complex64 fast slots): the 2^K x 2^K matrix sits in the pass header, i.e. in the constant
// bank, so every matrix element reaches the FMA pipe as a uniform-register operand (SASS: FFMA2 R, R.F32x2,
// UR.F32, R.F32x2) and costs no register and no shared-memory traffic.  The thread walks the COLUMNS of the
// matrix: one 16-byte shared load brings the two amplitudes of a unit, which are multiplied into all 2^K row
// accumulators at once (32 independent FFMA2 per column for K = 3) -- so only the accumulators (2^K pairs per
// group) and one column of inputs are live, whatever K is: 20 registers of state for K = 2, 36 for K = 3.
// The accumulation order per output amplitude is j ascending with re*re, -im*im / re*im, im*re interleaved,
// the order of the reference's scalar loop (U.h:93-94) up to fp32 FMA contraction.
//   KK   number of UNIT-level target bits (tbl_x has 2^KK entries)
//   LOW  false: no target on amplitude bit 0: K = KK, the even and the odd amplitude of a unit belong to two
//               different groups, the thread updates both;
//        true : matrix bit 0 is amplitude bit 0: K = KK + 1, a unit holds columns 2j, 2j + 1 of ONE group.
//   S    slot index when it is known at compile time (every matrix element is then a constant-bank operand at
//        an immediate offset), -1 = use the run-time `slot`
// The per-thread addressing constants of a slot gate (13 table entries).  The kernel loads them for gate g + 1
// right before the barrier that ends gate g, so their L1 latency is spent waiting at the barrier.
struct StreamRegs {
  uint32_t st;
  IterBasis ib;
  uint32_t xo[8];
};
HQ_DEV StreamRegs load_stream_regs(const HqGateDesc* __restrict__ g, int tid) {
  StreamRegs r;
  r.st = HQ_LDG(&g->tbl_thread[tid]);
  r.ib = load_iter_basis(g->tbl_iter);
  HQ_UNROLL
  for (int m = 0; m < 8; ++m) r.xo[m] = HQ_LDG(&g->tbl_x[m]);
  return r;
}

template <int KK, bool LOW, int S>
HQ_DEV void gate_stream_f32(float4* tile, const StreamRegs& sr, const HqPassHeader& ph, uint32_t slot,
                            int Tu, int tid) {
  const int UD = 1 << KK;                    // units per work item
  const int DIM = LOW ? 2 * UD : UD;         // matrix dimension
  const int nq = Tu - KK;
  const uint32_t nwork = 1u << nq;
  if (uint32_t(tid) >= nwork) return;
  const uint32_t niter = nq > HQ_THREADS_LOG2 ? (nwork >> HQ_THREADS_LOG2) : 1u;
  const uint32_t st = sr.st;
  const IterBasis ib = sr.ib;
  uint32_t xo[UD];
  HQ_UNROLL
  for (int m = 0; m < UD; ++m) xo[m] = sr.xo[m];
  // matrix element e of the slot: straight out of the constant bank (row-major, (re, im) interleaved)
#define HQ_FAST_U(e) (S >= 0 ? ph.fast_u[S >= 0 ? S : 0][e] : ph.fast_u[slot][e])
  // software pipeline over the work items of this thread: the shared loads of item it + 1 are issued before the
  // FFMA2 block of item it, so their latency hides behind 16 * 4^KK independent FFMA2 (different items never
  // share a slot, so loading ahead of the previous item's stores is safe)
  // MEASURED (profiles/r02/sweep_ring_c.jsonl): this costs the k = 2 gates 8 % (4.36 vs 4.02 ms for a 4-gate pass at
  // n = 30: the extra 16 live registers and moves outweigh the hidden latency when 24 warps per SM already
  // interleave), so it is compiled out by default.
#ifndef HQ_STREAM_PREFETCH
#define HQ_STREAM_PREFETCH 0
#endif
  const bool PREFETCH = HQ_STREAM_PREFETCH && UD <= 4;
  float4 v[UD];
  if (PREFETCH) {
    const uint32_t sb0 = st ^ iter_offset(ib, 0);
    HQ_UNROLL
    for (int m = 0; m < UD; ++m) v[m] = tile[sb0 ^ xo[m]];
  }
  HQ_NOUNROLL
  for (uint32_t it = 0; it < niter; ++it) {
    const uint32_t sb = st ^ iter_offset(ib, it);
    float4 nv[PREFETCH ? UD : 1];
    if (PREFETCH) {
      // (the last item re-reads itself instead of branching: the loads are simply not used)
      const uint32_t sbn = st ^ iter_offset(ib, it + 1 < niter ? it + 1 : it);
      HQ_UNROLL
      for (int m = 0; m < UD; ++m) nv[m] = tile[sbn ^ xo[m]];
    }
    if (!LOW) {
      F2 ae[DIM], ao[DIM];
      HQ_UNROLL
      for (int i = 0; i < DIM; ++i) { ae[i].lo = ae[i].hi = 0.f; ao[i].lo = ao[i].hi = 0.f; }
      HQ_UNROLL
      for (int j = 0; j < DIM; ++j) {
        const float4 w = PREFETCH ? v[j] : tile[sb ^ xo[j]];      // (no prefetch: loaded column by column)
        const F2 e = {w.x, w.y}, ie = {-w.y, w.x}, o = {w.z, w.w}, io = {-w.w, w.z};
        HQ_UNROLL
        for (int i = 0; i < DIM; ++i) {
          const float ur = HQ_FAST_U(2 * (i * DIM + j)), ui = HQ_FAST_U(2 * (i * DIM + j) + 1);
          ffma2_bcast(ae[i], e, ur);       // (ar, ai) += ur * (xr, xi)
          ffma2_bcast(ae[i], ie, ui);      // (ar, ai) += ui * (-xi, xr)
          ffma2_bcast(ao[i], o, ur);
          ffma2_bcast(ao[i], io, ui);
        }
      }
      HQ_UNROLL
      for (int i = 0; i < DIM; ++i) tile[sb ^ xo[i]] = make_float4(ae[i].lo, ae[i].hi, ao[i].lo, ao[i].hi);
    } else {
      F2 a[DIM];
      HQ_UNROLL
      for (int i = 0; i < DIM; ++i) a[i].lo = a[i].hi = 0.f;
      HQ_UNROLL
      for (int ju = 0; ju < UD; ++ju) {
        const float4 w = PREFETCH ? v[ju] : tile[sb ^ xo[ju]];
        const F2 e = {w.x, w.y}, ie = {-w.y, w.x}, o = {w.z, w.w}, io = {-w.w, w.z};
        HQ_UNROLL
        for (int i = 0; i < DIM; ++i) {
          const float ur0 = HQ_FAST_U(2 * (i * DIM + 2 * ju)), ui0 = HQ_FAST_U(2 * (i * DIM + 2 * ju) + 1);
          const float ur1 = HQ_FAST_U(2 * (i * DIM + 2 * ju + 1)), ui1 = HQ_FAST_U(2 * (i * DIM + 2 * ju + 1) + 1);
          ffma2_bcast(a[i], e, ur0);
          ffma2_bcast(a[i], ie, ui0);
          ffma2_bcast(a[i], o, ur1);
          ffma2_bcast(a[i], io, ui1);
        }
      }
      HQ_UNROLL
      for (int iu = 0; iu < UD; ++iu)
        tile[sb ^ xo[iu]] = make_float4(a[2 * iu].lo, a[2 * iu].hi, a[2 * iu + 1].lo, a[2 * iu + 1].hi);
    }
    if (PREFETCH) {
      HQ_UNROLL
      for (int m = 0; m < UD; ++m) v[m] = nv[m];
    }
  }
}

#undef HQ_FAST_U

// one fast slot: k = ph.fast_k[slot] in 1..3, low = matrix bit 0 on amplitude bit 0
template <int S, int MAXK>
HQ_DEV void gate_fast_f32(float4* tile, const StreamRegs& g, const HqPassHeader& ph, uint32_t slot,
                          int Tu, int tid) {
  const uint32_t k = ph.fast_k[slot] & 3u;
  const bool low = (ph.fast_k[slot] & 4u) != 0;
  if (!low) {
    if (k == 2) gate_stream_f32<2, false, S>(tile, g, ph, slot, Tu, tid);
    else if (k == 3) { if (MAXK >= 3) gate_stream_f32<3, false, S>(tile, g, ph, slot, Tu, tid); }
    else gate_stream_f32<1, false, S>(tile, g, ph, slot, Tu, tid);
  } else {
    if (k == 2) gate_stream_f32<1, true, S>(tile, g, ph, slot, Tu, tid);
    else if (k == 3) { if (MAXK >= 3) gate_stream_f32<2, true, S>(tile, g, ph, slot, Tu, tid); }
    else gate_stream_f32<0, true, S>(tile, g, ph, slot, Tu, tid);
  }
}

// complex64, matrix bit 0 sits on amplitude bit 0 (inside the unit): K = KK + 1 matrix bits,
// one group per thread, amplitude m = unit (m >> 1) half (m & 1).
template <int KK>
HQ_DEV void gate_small_f32_low(float4* tile, const HqGateDesc* __restrict__ g, const float2* __restrict__ U,
                               int Tu, int tid) {
  const int UD = 1 << KK;        // units per group
  const int DIM = 2 << KK;       // amplitudes per group
  const int nq = Tu - KK;
  const uint32_t nwork = 1u << nq;
  if (uint32_t(tid) >= nwork) return;
  const uint32_t niter = nq > HQ_THREADS_LOG2 ? (nwork >> HQ_THREADS_LOG2) : 1u;
  const uint32_t st = HQ_LDG(&g->tbl_thread[tid]);
  const IterBasis ib = load_iter_basis(g->tbl_iter);
  uint32_t xo[UD];
  HQ_UNROLL
  for (int m = 0; m < UD; ++m) xo[m] = HQ_LDG(&g->tbl_x[m]);

  float2 Ur[KK <= 1 ? DIM * DIM : 1];
  if (KK <= 1) {
    HQ_UNROLL
    for (int e = 0; e < DIM * DIM; ++e) Ur[e] = HQ_LDG(&U[e]);
  }

  HQ_NOUNROLL
  for (uint32_t it = 0; it < niter; ++it) {
    const uint32_t sb = st ^ iter_offset(ib, it);
    float4 in[UD];
    HQ_UNROLL
    for (int m = 0; m < UD; ++m) in[m] = tile[sb ^ xo[m]];
    if (KK <= 1) {
      HQ_UNROLL
      for (int iu = 0; iu < UD; ++iu) {
        float a0r = 0.f, a0i = 0.f, a1r = 0.f, a1i = 0.f;
        HQ_UNROLL
        for (int ju = 0; ju < UD; ++ju) {
          const float2 u00 = Ur[(2 * iu) * DIM + 2 * ju], u01 = Ur[(2 * iu) * DIM + 2 * ju + 1];
          const float2 u10 = Ur[(2 * iu + 1) * DIM + 2 * ju], u11 = Ur[(2 * iu + 1) * DIM + 2 * ju + 1];
          cmac(a0r, a0i, u00.x, u00.y, in[ju].x, in[ju].y);
          cmac(a0r, a0i, u01.x, u01.y, in[ju].z, in[ju].w);
          cmac(a1r, a1i, u10.x, u10.y, in[ju].x, in[ju].y);
          cmac(a1r, a1i, u11.x, u11.y, in[ju].z, in[ju].w);
        }
        tile[sb ^ xo[iu]] = make_float4(a0r, a0i, a1r, a1i);
      }
    } else {
      HQ_ROWUNROLL
      for (int iu = 0; iu < UD; ++iu) {
        float a0r = 0.f, a0i = 0.f, a1r = 0.f, a1i = 0.f;
        const float4* r0 = reinterpret_cast<const float4*>(U + (2 * iu) * DIM);
        const float4* r1 = reinterpret_cast<const float4*>(U + (2 * iu + 1) * DIM);
        HQ_UNROLL
        for (int ju = 0; ju < UD; ++ju) {
          const float4 ua = HQ_LDG(&r0[ju]);
          const float4 ub = HQ_LDG(&r1[ju]);
          cmac(a0r, a0i, ua.x, ua.y, in[ju].x, in[ju].y);
          cmac(a0r, a0i, ua.z, ua.w, in[ju].z, in[ju].w);
          cmac(a1r, a1i, ub.x, ub.y, in[ju].x, in[ju].y);
          cmac(a1r, a1i, ub.z, ub.w, in[ju].z, in[ju].w);
        }
        tile[sb ^ uint32_t(HQ_LDG(&g->tbl_x[iu]))] = make_float4(a0r, a0i, a1r, a1i);
      }
    }
  }
}

// complex128: unit = amplitude, one group per thread.
template <int KK>
HQ_DEV void gate_small_f64(double2* tile, const HqGateDesc* __restrict__ g, const double2* __restrict__ U,
                           int Tu, int tid) {
  const int DIM = 1 << KK;
  const int nq = Tu - KK;
  const uint32_t nwork = 1u << nq;
  if (uint32_t(tid) >= nwork) return;
  const uint32_t niter = nq > HQ_THREADS_LOG2 ? (nwork >> HQ_THREADS_LOG2) : 1u;
  const uint32_t st = HQ_LDG(&g->tbl_thread[tid]);
  const IterBasis ib = load_iter_basis(g->tbl_iter);
  uint32_t xo[DIM];
  HQ_UNROLL
  for (int m = 0; m < DIM; ++m) xo[m] = HQ_LDG(&g->tbl_x[m]);

  double2 Ur[KK <= 1 ? DIM * DIM : 1];
  if (KK <= 1) {
    HQ_UNROLL
    for (int e = 0; e < DIM * DIM; ++e) Ur[e] = HQ_LDG(&U[e]);
  }

  HQ_NOUNROLL
  for (uint32_t it = 0; it < niter; ++it) {
    const uint32_t sb = st ^ iter_offset(ib, it);
    double2 in[DIM];
    HQ_UNROLL
    for (int m = 0; m < DIM; ++m) in[m] = tile[sb ^ xo[m]];
    if (KK <= 1) {
      HQ_UNROLL
      for (int i = 0; i < DIM; ++i) {
        double ar = 0., ai = 0.;
        HQ_UNROLL
        for (int j = 0; j < DIM; ++j) cmac(ar, ai, Ur[i * DIM + j].x, Ur[i * DIM + j].y, in[j].x, in[j].y);
        tile[sb ^ xo[i]] = make_double2(ar, ai);
      }
    } else {
      HQ_ROWUNROLL
      for (int i = 0; i < DIM; ++i) {
        double ar = 0., ai = 0.;
        const double2* row = U + i * DIM;
        HQ_UNROLL
        for (int j = 0; j < DIM; ++j) {
          const double2 u = HQ_LDG(&row[j]);
          cmac(ar, ai, u.x, u.y, in[j].x, in[j].y);
        }
        tile[sb ^ uint32_t(HQ_LDG(&g->tbl_x[i]))] = make_double2(ar, ai);
      }
    }
  }
}

// complex128, row-pair scheme for K = 2, 3: a group's 2^K rows are produced by 2^(K-1) adjacent
// lanes, each holding two matrix rows in registers for the whole gate, so one 16-byte shared
// load feeds 8 DFMAs (the plain scheme above feeds 4 and is bound by the matrix loads).
// The lanes of a group sit in one warp: a __syncwarp() between the loads and the stores keeps
// the update in place.  `sync_warp` is a no-op in the CPU emulation (lanes run in sequence there,
// so the emulation splits the function in two phases instead).
template <int KK>
struct RowPairRegs {
  double2 u0[1 << KK], u1[1 << KK];
  uint32_t xo[1 << KK];
};

template <int KK>
HQ_DEV void rowpair_load_rows(RowPairRegs<KK>& r, const double2* __restrict__ U, int tid) {
  const int DIM = 1 << KK;
  const int rp = tid & ((DIM >> 1) - 1);
  HQ_UNROLL
  for (int j = 0; j < DIM; ++j) {
    r.u0[j] = HQ_LDG(&U[(2 * rp) * DIM + j]);
    r.u1[j] = HQ_LDG(&U[(2 * rp + 1) * DIM + j]);
  }
}
template <int KK>
HQ_DEV void rowpair_load_offsets(RowPairRegs<KK>& r, const HqGateDesc* __restrict__ g) {
  HQ_UNROLL
  for (int j = 0; j < (1 << KK); ++j) r.xo[j] = HQ_LDG(&g->tbl_x[j]);
}

// phase A of one iteration: gather the group's inputs and compute the two output rows
template <int KK>
HQ_DEV bool rowpair_compute(const double2* tile, const HqGateDesc* __restrict__ g, const RowPairRegs<KK>& r,
                            int Tu, int tid, uint32_t it, double2& o0, double2& o1, uint32_t& s0, uint32_t& s1) {
  const int DIM = 1 << KK;
  const int lrp = KK - 1;
  const int nq = Tu - KK;
  const uint32_t gs = uint32_t(tid) >> lrp;
  const int gbits = HQ_THREADS_LOG2 - lrp;
  if (nq < gbits && gs >= (1u << nq)) return false;
  const uint32_t rp = uint32_t(tid) & ((DIM >> 1) - 1);
  const uint32_t sb = uint32_t(HQ_LDG(&g->tbl_rthread[gs])) ^ uint32_t(HQ_LDG(&g->tbl_riter[it]));
  double a0r = 0., a0i = 0., a1r = 0., a1i = 0.;
  HQ_UNROLL
  for (int j = 0; j < DIM; ++j) {
    const double2 x = tile[sb ^ r.xo[j]];
    cmac(a0r, a0i, r.u0[j].x, r.u0[j].y, x.x, x.y);
    cmac(a1r, a1i, r.u1[j].x, r.u1[j].y, x.x, x.y);
  }
  o0 = make_double2(a0r, a0i);
  o1 = make_double2(a1r, a1i);
  s0 = sb ^ uint32_t(HQ_LDG(&g->tbl_x[2 * rp]));
  s1 = sb ^ uint32_t(HQ_LDG(&g->tbl_x[2 * rp + 1]));
  return true;
}

HQ_DEV uint32_t rowpair_iters(int Tu, int KK) {
  const int nq = Tu - KK;
  const int gbits = HQ_THREADS_LOG2 - (KK - 1);
  return nq > gbits ? (1u << (nq - gbits)) : 1u;
}

#ifdef __CUDACC__
template <int KK>
__device__ __forceinline__ void gate_rowpair_f64(double2* tile, const HqGateDesc* __restrict__ g,
                                                 const double2* __restrict__ U, int Tu, int tid) {
  RowPairRegs<KK> r;
  rowpair_load_rows<KK>(r, U, tid);
  rowpair_load_offsets<KK>(r, g);
  const uint32_t niter = rowpair_iters(Tu, KK);
  HQ_NOUNROLL
  for (uint32_t it = 0; it < niter; ++it) {
    double2 o0, o1;
    uint32_t s0 = 0, s1 = 0;
    const bool ok = rowpair_compute<KK>(tile, g, r, Tu, tid, it, o0, o1, s0, s1);
    __syncwarp();              // every lane of the group has read its inputs
    if (ok) {
      tile[s0] = o0;
      tile[s1] = o1;
    }
  }
}
#endif

// ---------------------------------------------------------------------------------------
// two-phase path for k >= 5 (amplitude granularity, any precision).  A round handles
// HQ_THREADS * HQ_BIG_ROWS / 2^k whole groups: in phase A every thread accumulates
// HQ_BIG_ROWS output rows of one group from shared memory into registers, the CTA
// synchronises, and in phase B the rows are written back.  Lanes of a warp share the row
// block (so the matrix loads are warp-uniform) and differ in the group.
// The matrix is stored column-major: Ut[j * dim + i] = U[i][j].
// ---------------------------------------------------------------------------------------
template <typename T>
struct BigAcc {
  T re[HQ_BIG_ROWS], im[HQ_BIG_ROWS];
  uint32_t base;    // local amplitude index of the group (zeros at the target bits)
  uint32_t row0;
  bool active;
};

template <typename T>
HQ_DEV uint32_t amp_slot(uint32_t a) {   // local amplitude index -> index into a Cplx view of the tile
  const int V = Traits<T>::V;
  return (swz(a >> V) << V) | (a & ((1u << V) - 1u));
}

HQ_DEV int big_rounds(int T, int k) {
  // groups in the tile / groups per round
  const int groups_log2 = T - k;
  const int per_round_log2 = HQ_THREADS_LOG2 + 3 - k;   // log2(HQ_THREADS * HQ_BIG_ROWS / 2^k)
  const int r = groups_log2 - per_round_log2;
  return r > 0 ? (1 << r) : 1;
}

template <typename T>
HQ_DEV void gate_big_phaseA(const typename Traits<T>::Cplx* tile, const HqGateDesc& g,
                            const typename Traits<T>::Cplx* __restrict__ Ut, int Tbits, int tid,
                            int round, BigAcc<T>& acc) {
  const int k = int(g.k);
  const uint32_t dim = 1u << k;
  const int groups_log2 = Tbits - k;
  int per_round_log2 = HQ_THREADS_LOG2 + 3 - k;
  if (per_round_log2 > groups_log2) per_round_log2 = groups_log2;
  const uint32_t gl = uint32_t(tid) & ((1u << per_round_log2) - 1u);
  const uint32_t rb = uint32_t(tid) >> per_round_log2;     // row block
  acc.active = (rb * HQ_BIG_ROWS) < dim;
  if (!acc.active) return;
  const uint32_t grp = (uint32_t(round) << per_round_log2) | gl;
  acc.base = scatter_bits(grp, g.q, 0, groups_log2);
  acc.row0 = rb * HQ_BIG_ROWS;
  HQ_UNROLL
  for (int r = 0; r < HQ_BIG_ROWS; ++r) { acc.re[r] = T(0); acc.im[r] = T(0); }
  HQ_NOUNROLL
  for (uint32_t j = 0; j < dim; ++j) {
    const uint32_t a = acc.base | uint32_t(deposit(j, g.tpos, k));
    const typename Traits<T>::Cplx x = tile[amp_slot<T>(a)];
    const typename Traits<T>::Cplx* col = Ut + size_t(j) * dim + acc.row0;
    HQ_UNROLL
    for (int r = 0; r < HQ_BIG_ROWS; ++r) {
      const typename Traits<T>::Cplx u = HQ_LDG(&col[r]);
      cmac(acc.re[r], acc.im[r], u.x, u.y, x.x, x.y);
    }
  }
}

template <typename T>
HQ_DEV void gate_big_phaseB(typename Traits<T>::Cplx* tile, const HqGateDesc& g,
                            const BigAcc<T>& acc) {
  if (!acc.active) return;
  const int k = int(g.k);
  HQ_UNROLL
  for (int r = 0; r < HQ_BIG_ROWS; ++r) {
    const uint32_t a = acc.base | uint32_t(deposit(acc.row0 + r, g.tpos, k));
    typename Traits<T>::Cplx o;
    o.x = acc.re[r];
    o.y = acc.im[r];
    tile[amp_slot<T>(a)] = o;
  }
}

// ---------------------------------------------------------------------------------------
// "scalar + rank one" gates (HQ_GATE_DR1): psi' = lambda psi + u (v . psi) on every group of 2^k amplitudes.
// Same lane tables and work-item layout as the register path; `p` = [lambda, u[2^k], v[2^k]] (complex).
// A thread first reduces v . psi over its group(s), then rescales and adds: 2 * 2^k complex MACs per group.
// ---------------------------------------------------------------------------------------
template <typename R>
HQ_DEV void cmul_add(R& ar, R& ai, R ur, R ui, R xr, R xi) { cmac(ar, ai, ur, ui, xr, xi); }

// Register-light on purpose (two sweeps over the group's units, nothing but the two dot products live between
// them): the gate shares a kernel with the FFMA2 slots and must not push their accumulators out of registers; the
// second read of the units costs half a shared-memory round trip.
template <int KK, bool LOW>
HQ_DEV void gate_dr1_f32(float4* tile, const HqGateDesc* __restrict__ g, const float2* __restrict__ p, int Tu, int tid) {
  const int UD = 1 << KK;
  const int DIM = LOW ? 2 * UD : UD;
  const int nq = Tu - KK;
  const uint32_t nwork = 1u << nq;
  if (uint32_t(tid) >= nwork) return;
  const uint32_t niter = nq > HQ_THREADS_LOG2 ? (nwork >> HQ_THREADS_LOG2) : 1u;
  const uint32_t st = HQ_LDG(&g->tbl_thread[tid]);
  const IterBasis ib = load_iter_basis(g->tbl_iter);
  const float2 lam = HQ_LDG(&p[0]);
  const float2* __restrict__ u = p + 1;
  const float2* __restrict__ v = p + 1 + DIM;
  // sparse form (planner, "sparse scalar + rank one"): lambda has been moved into another matrix of the plan and at
  // most four units of the group hold a non-zero u or v component; the trailer lists them compactly -- per unit its
  // slot offset and its v and u components -- so that every load below is independent of the others (one latency)
  // and the gate is 4 LDS + 4 STS per work item.  Nothing else of the group is touched.
  const float2* __restrict__ sp = p + 1 + 2 * DIM;
  if (HQ_LDG(&sp[0]).y != 0.f) {
    uint32_t xm[4];
    float2 va[4], vb[4], ua[4], ub[4];
    HQ_UNROLL
    for (int s = 0; s < 4; ++s) {
      xm[s] = uint32_t(HQ_LDG(&sp[1 + 5 * s]).x);
      va[s] = HQ_LDG(&sp[2 + 5 * s]);
      ua[s] = HQ_LDG(&sp[4 + 5 * s]);
      if (LOW) {
        vb[s] = HQ_LDG(&sp[3 + 5 * s]);
        ub[s] = HQ_LDG(&sp[5 + 5 * s]);
      }
    }
    HQ_NOUNROLL
    for (uint32_t it = 0; it < niter; ++it) {
      const uint32_t sb = st ^ iter_offset(ib, it);
      float4 x[4];
      HQ_UNROLL
      for (int s = 0; s < 4; ++s) x[s] = tile[sb ^ xm[s]];
      float d0r = 0.f, d0i = 0.f, d1r = 0.f, d1i = 0.f;
      HQ_UNROLL
      for (int s = 0; s < 4; ++s) {
        if (!LOW) {
          cmac(d0r, d0i, va[s].x, va[s].y, x[s].x, x[s].y);
          cmac(d1r, d1i, va[s].x, va[s].y, x[s].z, x[s].w);
        } else {
          cmac(d0r, d0i, va[s].x, va[s].y, x[s].x, x[s].y);
          cmac(d0r, d0i, vb[s].x, vb[s].y, x[s].z, x[s].w);
        }
      }
      HQ_UNROLL
      for (int s = 0; s < 4; ++s) {
        if (!LOW) {
          cmac(x[s].x, x[s].y, ua[s].x, ua[s].y, d0r, d0i);
          cmac(x[s].z, x[s].w, ua[s].x, ua[s].y, d1r, d1i);
        } else {
          cmac(x[s].x, x[s].y, ua[s].x, ua[s].y, d0r, d0i);
          cmac(x[s].z, x[s].w, ub[s].x, ub[s].y, d0r, d0i);
        }
        tile[sb ^ xm[s]] = x[s];
      }
    }
    return;
  }
  HQ_NOUNROLL
  for (uint32_t it = 0; it < niter; ++it) {
    const uint32_t sb = st ^ iter_offset(ib, it);
    float d0r = 0.f, d0i = 0.f, d1r = 0.f, d1i = 0.f;      // v . psi of the even / odd group (or of the one group)
    HQ_UNROLL
    for (int m = 0; m < UD; ++m) {
      const float4 x = tile[sb ^ uint32_t(HQ_LDG(&g->tbl_x[m]))];
      if (!LOW) {
        const float2 vj = HQ_LDG(&v[m]);
        cmac(d0r, d0i, vj.x, vj.y, x.x, x.y);
        cmac(d1r, d1i, vj.x, vj.y, x.z, x.w);
      } else {
        const float2 va = HQ_LDG(&v[2 * m]), vb = HQ_LDG(&v[2 * m + 1]);
        cmac(d0r, d0i, va.x, va.y, x.x, x.y);
        cmac(d0r, d0i, vb.x, vb.y, x.z, x.w);
      }
    }
    HQ_UNROLL
    for (int m = 0; m < UD; ++m) {
      const uint32_t slot = sb ^ uint32_t(HQ_LDG(&g->tbl_x[m]));
      const float4 x = tile[slot];
      float ar = 0.f, ai = 0.f, br = 0.f, bi = 0.f;
      cmac(ar, ai, lam.x, lam.y, x.x, x.y);
      cmac(br, bi, lam.x, lam.y, x.z, x.w);
      if (!LOW) {
        const float2 ui = HQ_LDG(&u[m]);
        cmac(ar, ai, ui.x, ui.y, d0r, d0i);
        cmac(br, bi, ui.x, ui.y, d1r, d1i);
      } else {
        const float2 ua = HQ_LDG(&u[2 * m]), ub = HQ_LDG(&u[2 * m + 1]);
        cmac(ar, ai, ua.x, ua.y, d0r, d0i);
        cmac(br, bi, ub.x, ub.y, d0r, d0i);
      }
      tile[slot] = make_float4(ar, ai, br, bi);
    }
  }
}

template <int KK>
HQ_DEV void gate_dr1_f64(double2* tile, const HqGateDesc* __restrict__ g, const double2* __restrict__ p, int Tu, int tid) {
  const int DIM = 1 << KK;
  const int nq = Tu - KK;
  const uint32_t nwork = 1u << nq;
  if (uint32_t(tid) >= nwork) return;
  const uint32_t niter = nq > HQ_THREADS_LOG2 ? (nwork >> HQ_THREADS_LOG2) : 1u;
  const uint32_t st = HQ_LDG(&g->tbl_thread[tid]);
  const IterBasis ib = load_iter_basis(g->tbl_iter);
  const double2 lam = HQ_LDG(&p[0]);
  const double2* __restrict__ u = p + 1;
  const double2* __restrict__ v = p + 1 + DIM;
  const double2* __restrict__ sp = p + 1 + 2 * DIM;           // sparse form, see gate_dr1_f32
  if (HQ_LDG(&sp[0]).y != 0.) {
    uint32_t xm[4];
    double2 va[4], ua[4];
    HQ_UNROLL
    for (int s = 0; s < 4; ++s) {
      xm[s] = uint32_t(HQ_LDG(&sp[1 + 5 * s]).x);
      va[s] = HQ_LDG(&sp[2 + 5 * s]);
      ua[s] = HQ_LDG(&sp[4 + 5 * s]);
    }
    HQ_NOUNROLL
    for (uint32_t it = 0; it < niter; ++it) {
      const uint32_t sb = st ^ iter_offset(ib, it);
      double2 x[4];
      HQ_UNROLL
      for (int s = 0; s < 4; ++s) x[s] = tile[sb ^ xm[s]];
      double dr = 0., di = 0.;
      HQ_UNROLL
      for (int s = 0; s < 4; ++s) cmac(dr, di, va[s].x, va[s].y, x[s].x, x[s].y);
      HQ_UNROLL
      for (int s = 0; s < 4; ++s) {
        cmac(x[s].x, x[s].y, ua[s].x, ua[s].y, dr, di);
        tile[sb ^ xm[s]] = x[s];
      }
    }
    return;
  }
  HQ_NOUNROLL
  for (uint32_t it = 0; it < niter; ++it) {
    const uint32_t sb = st ^ iter_offset(ib, it);
    double dr = 0., di = 0.;
    HQ_UNROLL
    for (int m = 0; m < DIM; ++m) {
      const double2 x = tile[sb ^ uint32_t(HQ_LDG(&g->tbl_x[m]))];
      const double2 vj = HQ_LDG(&v[m]);
      cmac(dr, di, vj.x, vj.y, x.x, x.y);
    }
    HQ_UNROLL
    for (int m = 0; m < DIM; ++m) {
      const uint32_t slot = sb ^ uint32_t(HQ_LDG(&g->tbl_x[m]));
      const double2 x = tile[slot];
      double ar = 0., ai = 0.;
      const double2 ui = HQ_LDG(&u[m]);
      cmac(ar, ai, lam.x, lam.y, x.x, x.y);
      cmac(ar, ai, ui.x, ui.y, dr, di);
      tile[slot] = make_double2(ar, ai);
    }
  }
}

HQ_DEV void gate_dr1_dispatch(float4* tile, const HqGateDesc* g, uint32_t k, const unsigned char* prog, uint32_t mat_off,
                              int Tu, int tid) {
  const float2* p = reinterpret_cast<const float2*>(prog + mat_off);
  const bool low = HQ_LDG(&g->tpos[0]) == 0;
  if (!low) {
    if (k == 3) gate_dr1_f32<3, false>(tile, g, p, Tu, tid);
    else if (k == 4) gate_dr1_f32<4, false>(tile, g, p, Tu, tid);
  } else {
    if (k == 3) gate_dr1_f32<2, true>(tile, g, p, Tu, tid);
    else if (k == 4) gate_dr1_f32<3, true>(tile, g, p, Tu, tid);
  }
}
HQ_DEV void gate_dr1_dispatch(double2* tile, const HqGateDesc* g, uint32_t k, const unsigned char* prog, uint32_t mat_off,
                              int Tu, int tid) {
  const double2* p = reinterpret_cast<const double2*>(prog + mat_off);
  if (k == 3) gate_dr1_f64<3>(tile, g, p, Tu, tid);
  else if (k == 4) gate_dr1_f64<4>(tile, g, p, Tu, tid);
}

// ---------------------------------------------------------------------------------------
// dispatch of one register-path gate
// ---------------------------------------------------------------------------------------
// MAXK prunes the switch so that a pass made of small gates only is compiled with few
// registers (the kernel is instantiated per gate class, see hq_kernels.cu).
template <int MAXK>
HQ_DEV void gate_small_dispatch(float4* tile, const HqGateDesc* g, uint32_t k, bool low,
                                const unsigned char* prog, uint32_t mat_off, int Tu, int tid) {
  const float2* U = reinterpret_cast<const float2*>(prog + mat_off);
  if (!low) {
    switch (k) {
      case 1: gate_small_f32<1>(tile, g, U, Tu, tid); break;
      case 2: gate_small_f32<2>(tile, g, U, Tu, tid); break;
      case 3: if (MAXK >= 3) gate_small_f32<3>(tile, g, U, Tu, tid); break;
      case 4: if (MAXK >= 4) gate_small_f32<4>(tile, g, U, Tu, tid); break;
      default: break;
    }
  } else {
    switch (k) {
      case 1: gate_small_f32_low<0>(tile, g, U, Tu, tid); break;
      case 2: gate_small_f32_low<1>(tile, g, U, Tu, tid); break;
      case 3: if (MAXK >= 3) gate_small_f32_low<2>(tile, g, U, Tu, tid); break;
      case 4: if (MAXK >= 4) gate_small_f32_low<3>(tile, g, U, Tu, tid); break;
      default: break;
    }
  }
}

template <int MAXK>
HQ_DEV void gate_small_dispatch(double2* tile, const HqGateDesc* g, uint32_t k, bool /*low*/,
                                const unsigned char* prog, uint32_t mat_off, int Tu, int tid) {
  const double2* U = reinterpret_cast<const double2*>(prog + mat_off);
  switch (k) {
    case 1: gate_small_f64<1>(tile, g, U, Tu, tid); break;
    case 2: gate_small_f64<2>(tile, g, U, Tu, tid); break;
    case 3: if (MAXK >= 3) gate_small_f64<3>(tile, g, U, Tu, tid); break;
    case 4: if (MAXK >= 4) gate_small_f64<4>(tile, g, U, Tu, tid); break;
    default: break;
  }
}

// ---------------------------------------------------------------------------------------
// tile fill / drain address math (the copies themselves are in the kernel / the emulator)
// ---------------------------------------------------------------------------------------
// local unit c of tile -> global unit index, given the tile's first unit and the run table.
// Offset (in units) of local unit c from the tile's first unit: the low Lu bits stay, the
// bits above go to the tile's high positions.  Linear over disjoint bit sets.
HQ_HD uint64_t unit_offset(uint32_t c, int Lu, int V, const uint8_t* high_pos, int n_high) {
  return (deposit(c >> Lu, high_pos, n_high) >> V) + (c & ((1u << Lu) - 1u));
}

HQ_DEV float4 make_unit(const float2* a) { return make_float4(a[0].x, a[0].y, a[1].x, a[1].y); }
HQ_DEV double2 make_unit(const double2* a) { return a[0]; }

// Permuted drain: value of local amplitude j after the in-tile bit permutation.
HQ_DEV uint32_t perm_src(uint32_t j, const uint8_t* perm, int Tbits) {
  uint32_t y = 0;
  for (int i = 0; i < Tbits; ++i) y |= ((j >> i) & 1u) << perm[i];
  return y;
}

}  // namespace hq
