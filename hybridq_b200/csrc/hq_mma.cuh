// hq_mma.cuh -- tensor-core gate arithmetic of the tile kernel (device code only, sm_100a).
//
// A k-qubit gate applied to a tile is the real matrix product  D = A * B  with
//   A[row][kk]   row = one group of 2^k amplitudes, kk = the 2 * 2^k reals of the group
//   B[kk][n]     the real form of U^T:  (re, im) x (re, im) blocks  [[Ur, Ui], [-Ui, Ur]]
// executed with warp-level mma.sync on the amplitudes where they already sit in shared memory:
//   complex64    mma.m16n8k8 TF32 x3 (hi*hi + hi*lo + lo*hi with fp32 accumulation: the "3xTF32"
//                split keeps fp32-level accuracy)
//   complex128   mma.m8n8k4 FP64 (DMMA)
// The arithmetic replaces the per-group mat-vec of /root/reference/include/U.h:77-99 (k <= 4)
// and :174-199 (k >= 5); nothing of the reference's structure survives.
//
// Fragment layouts (PTX ISA, verified numerically on the B200 by tools/microbench_mma.cu):
//   m16n8k8 tf32  lane = 4 g + t
//       A: a0 = A[g][t]     a1 = A[g+8][t]    a2 = A[g][t+4]    a3 = A[g+8][t+4]
//       B: b0 = B[t][g]     b1 = B[t+4][g]
//       D: d0 = D[g][2t]    d1 = D[g][2t+1]   d2 = D[g+8][2t]   d3 = D[g+8][2t+1]
//   m8n8k4 f64
//       A: a0 = A[g][t]     B: b0 = B[t][g]   D: d0 = D[g][2t]  d1 = D[g][2t+1]
//
// Column order is ours to choose.  With  kk = t -> re(amp 4s+t),  kk = t+4 -> im(amp 4s+t)  and
// n = 2t, 2t+1 -> (re, im) of output amplitude 4j+t,  lane (g, t) loads exactly the amplitudes
// m = t (mod 4) of its rows and stores exactly the same ones: the update is in place with no
// synchronisation beyond the warp-synchronous mma itself.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hq {

__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

// x = hi + lo for the 3xTF32 product.  SPLIT selects how:
//   0  hi = x as it is (the tensor core reads the upper 19 bits of a tf32 operand, i.e. truncates),
//      lo = x - trunc(x): LOP3 + FADD, no extra register for hi
//   1  hi = round-to-nearest by integer add + mask, lo = x - hi: 3 instructions
//   2  hi = cvt.rna.tf32.f32 (ptxas expands it to 4 instructions with an inf/nan guard), lo = x - hi
//   3  like 1, and lo is rounded to nearest TF32 too (5 instructions)
// lo always goes to the tensor core as raw fp32 bits.  Relative error per product: about 2^-20
// (SPLIT 0) or 2^-21 (1, 2).  Measured on B200 (profiles/r01/microbench_mma_b.jsonl, 16 k = 2 gates
// on a tile): max-abs error 3.6e-7 (SPLIT 0) vs 1.2e-7 (SPLIT 1) at the same speed, so 1 is the default.
#ifndef HQ_TF32_SPLIT
#define HQ_TF32_SPLIT 1
#endif
// The tensor core adds into its fp32 accumulator with truncation (round toward zero): chaining the big
// hi*hi products through the accumulator shrinks every amplitude by about one ulp per gate -- measured
// norm^2 - 1 = -1.0e-4 after 600 k = 3 gates against -4e-8 on the FMA path (tools/norm_drift.py).  With
// HQ_MMA_ACC_OUTSIDE (default) every hi*hi product is issued with a zero accumulator and summed with
// round-to-nearest FADDs, and only the small correction terms (lo*hi, hi*lo) are chained inside the
// tensor core (the scheme of Ootomo & Yokota, "Recovering single precision accuracy from Tensor Cores").
#ifndef HQ_MMA_ACC_OUTSIDE
#define HQ_MMA_ACC_OUTSIDE 1
#endif
template <int SPLIT>
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  if (SPLIT == 0) {
    hi = __float_as_uint(x);
    lo = __float_as_uint(x - __uint_as_float(hi & 0xffffe000u));
  } else if (SPLIT == 1) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
  } else if (SPLIT == 2) {
    hi = tf32_rna(x);
    lo = __float_as_uint(x - __uint_as_float(hi));
  } else {
    // 3: hi and lo both rounded to nearest TF32 by add + mask, so that nothing is left to the tensor
    // core's truncation of its operands (which is biased toward zero): 5 instructions
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
  }
}

// the hi*hi product of one k-step: chained in the tensor core (d) or, by default, computed against a zero
// accumulator and added to `big` with round-to-nearest FADDs
__device__ __forceinline__ void mma_main(float (&d)[4], float (&big)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#if HQ_MMA_ACC_OUTSIDE == 1
  float m[4] = {0.f, 0.f, 0.f, 0.f};
  mma_tf32(m, a, b0, b1);
#pragma unroll
  for (int e = 0; e < 4; ++e) big[e] += m[e];
#elif HQ_MMA_ACC_OUTSIDE == 2
  (void)d;
  mma_tf32(big, a, b0, b1);     // the big products chained among themselves, apart from the small terms
#else
  (void)big;
  mma_tf32(d, a, b0, b1);
#endif
}

// Order of a lane's four A values inside a 16-byte complex64 unit (re_even, im_even, re_odd,
// im_odd): rows g / g+8 are the even / odd amplitude of the unit, columns t / t+4 are re / im.
__device__ __forceinline__ void unit_to_afrag(const float4& v, float (&a)[4]) {
  a[0] = v.x; a[1] = v.z; a[2] = v.y; a[3] = v.w;
}

// ------------------------------------------------------------------------------------------
// Common shape of the three gate loops below.  One call handles UNR independent row sets (UNR
// warp-iterations) so that their shared-memory loads, mma chains and stores overlap; KS = 2^k / 4
// k-steps and as many n-tiles.
//   sb[u]  slot of (row g, m = 0) of row set u for this lane
//   xo[s]  XOR offset of amplitude/unit m = 4 s + t
//   bf     B fragments of this lane: bf[(s * KS + j) * 32] = block (k-step s, n-tile j);
//          BREG = they are already in registers (breg[s * KS + j])
//   xtab   &tbl_x[t]: the n-tile loop of the larger gates (KS >= 8) is kept rolled (code size,
//          registers), so the store offset of n-tile j is re-read from the table: xtab[4 j]
// ------------------------------------------------------------------------------------------

// complex64, UNIT granularity (amplitude bit 0 is not a target): a row set is 8 unit-groups = 16
// groups; B fragment = (b0_hi, b1_hi, b0_lo, b1_lo).
template <int KS, int UNR, bool BREG, bool FLY, int SPLIT = HQ_TF32_SPLIT>
__device__ __forceinline__ void mma_iter_f32_unit(float4* tile, const uint32_t (&sb)[UNR], const uint32_t (&xo)[KS],
                                                  const float4* __restrict__ bf, const float4* breg,
                                                  const uint16_t* __restrict__ xtab) {
  float raw[UNR][FLY ? KS : 1][4];
  uint32_t hi[UNR][FLY ? 1 : KS][4], lo[UNR][FLY ? 1 : KS][4];
#pragma unroll
  for (int u = 0; u < UNR; ++u)
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const float4 v = tile[sb[u] ^ xo[s]];
      float a[4];
      unit_to_afrag(v, a);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (FLY) raw[u][s][e] = a[e];
        else tf32_split<SPLIT>(a[e], hi[u][s][e], lo[u][s][e]);
      }
    }
#pragma unroll(KS >= 8 ? 1 : KS)
  for (int j = 0; j < KS; ++j) {
    float d[UNR][4], big[UNR][4];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      d[u][0] = d[u][1] = d[u][2] = d[u][3] = 0.f;
      big[u][0] = big[u][1] = big[u][2] = big[u][3] = 0.f;
    }
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const float4 b = BREG ? breg[(KS >= 8 ? 0 : s * KS + j)] : __ldg(&bf[(s * KS + j) * 32]);
      const uint32_t bh0 = __float_as_uint(b.x), bh1 = __float_as_uint(b.y);
      const uint32_t bl0 = __float_as_uint(b.z), bl1 = __float_as_uint(b.w);
      if (FLY) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) tf32_split<SPLIT>(raw[u][s][e], h[e], l[e]);
          mma_tf32(d[u], l, bh0, bh1);
          mma_tf32(d[u], h, bl0, bl1);
          mma_main(d[u], big[u], h, bh0, bh1);
        }
      } else {
#pragma unroll
        for (int u = 0; u < UNR; ++u) mma_tf32(d[u], lo[u][s], bh0, bh1);
#pragma unroll
        for (int u = 0; u < UNR; ++u) mma_tf32(d[u], hi[u][s], bl0, bl1);
#pragma unroll
        for (int u = 0; u < UNR; ++u) mma_main(d[u], big[u], hi[u][s], bh0, bh1);
      }
    }
#if HQ_MMA_ACC_OUTSIDE
#pragma unroll
    for (int u = 0; u < UNR; ++u)
#pragma unroll
      for (int e = 0; e < 4; ++e) d[u][e] += big[u][e];
#endif
    const uint32_t xj = KS >= 8 ? uint32_t(__ldg(&xtab[4 * j])) : xo[KS >= 8 ? 0 : j];
#pragma unroll
    for (int u = 0; u < UNR; ++u) tile[sb[u] ^ xj] = make_float4(d[u][0], d[u][1], d[u][2], d[u][3]);
  }
}

// complex64, AMPLITUDE granularity (any targets; used when amplitude bit 0 is a target): a row set
// is 16 groups, rows g and g+8 being two different groups; `tile` is viewed as float2 amplitudes,
// sb[u] is the float2 slot of (row g, m = 0), sb[u] ^ row8 that of (row g+8, m = 0).
template <int KS, int UNR, bool BREG, bool FLY, int SPLIT = HQ_TF32_SPLIT>
__device__ __forceinline__ void mma_iter_f32_amp(float2* tile, const uint32_t (&sb)[UNR], uint32_t row8,
                                                 const uint32_t (&xo)[KS], const float4* __restrict__ bf,
                                                 const float4* breg, const uint16_t* __restrict__ xtab) {
  float raw[UNR][FLY ? KS : 1][4];
  uint32_t hi[UNR][FLY ? 1 : KS][4], lo[UNR][FLY ? 1 : KS][4];
#pragma unroll
  for (int u = 0; u < UNR; ++u)
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const float2 p = tile[sb[u] ^ xo[s]];
      const float2 q = tile[sb[u] ^ row8 ^ xo[s]];
      const float a[4] = {p.x, q.x, p.y, q.y};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (FLY) raw[u][s][e] = a[e];
        else tf32_split<SPLIT>(a[e], hi[u][s][e], lo[u][s][e]);
      }
    }
#pragma unroll(KS >= 8 ? 1 : KS)
  for (int j = 0; j < KS; ++j) {
    float d[UNR][4], big[UNR][4];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      d[u][0] = d[u][1] = d[u][2] = d[u][3] = 0.f;
      big[u][0] = big[u][1] = big[u][2] = big[u][3] = 0.f;
    }
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const float4 b = BREG ? breg[(KS >= 8 ? 0 : s * KS + j)] : __ldg(&bf[(s * KS + j) * 32]);
      const uint32_t bh0 = __float_as_uint(b.x), bh1 = __float_as_uint(b.y);
      const uint32_t bl0 = __float_as_uint(b.z), bl1 = __float_as_uint(b.w);
      if (FLY) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) tf32_split<SPLIT>(raw[u][s][e], h[e], l[e]);
          mma_tf32(d[u], l, bh0, bh1);
          mma_tf32(d[u], h, bl0, bl1);
          mma_main(d[u], big[u], h, bh0, bh1);
        }
      } else {
#pragma unroll
        for (int u = 0; u < UNR; ++u) mma_tf32(d[u], lo[u][s], bh0, bh1);
#pragma unroll
        for (int u = 0; u < UNR; ++u) mma_tf32(d[u], hi[u][s], bl0, bl1);
#pragma unroll
        for (int u = 0; u < UNR; ++u) mma_main(d[u], big[u], hi[u][s], bh0, bh1);
      }
    }
#if HQ_MMA_ACC_OUTSIDE
#pragma unroll
    for (int u = 0; u < UNR; ++u)
#pragma unroll
      for (int e = 0; e < 4; ++e) d[u][e] += big[u][e];
#endif
    const uint32_t xj = KS >= 8 ? uint32_t(__ldg(&xtab[4 * j])) : xo[KS >= 8 ? 0 : j];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      tile[sb[u] ^ xj] = make_float2(d[u][0], d[u][1]);
      tile[sb[u] ^ row8 ^ xj] = make_float2(d[u][2], d[u][3]);
    }
  }
}

// complex128 (unit = amplitude): a row set is 8 groups.  One 16-byte load feeds two k-steps (the
// re column and the im column of amplitude 4 s + t); B fragment = (b of the re step, b of the im step).
template <int KS, int UNR, bool BREG>
__device__ __forceinline__ void dmma_iter_f64(double2* tile, const uint32_t (&sb)[UNR], const uint32_t (&xo)[KS],
                                              const double2* __restrict__ bf, const double2* breg,
                                              const uint16_t* __restrict__ xtab) {
  double2 x[UNR][KS];
#pragma unroll
  for (int u = 0; u < UNR; ++u)
#pragma unroll
    for (int s = 0; s < KS; ++s) x[u][s] = tile[sb[u] ^ xo[s]];
#pragma unroll(KS >= 8 ? 1 : KS)
  for (int j = 0; j < KS; ++j) {
    double d0[UNR], d1[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) d0[u] = d1[u] = 0.;
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const double2 b = BREG ? breg[(KS >= 8 ? 0 : s * KS + j)] : __ldg(&bf[(s * KS + j) * 32]);
#pragma unroll
      for (int u = 0; u < UNR; ++u) dmma(d0[u], d1[u], x[u][s].x, b.x);
#pragma unroll
      for (int u = 0; u < UNR; ++u) dmma(d0[u], d1[u], x[u][s].y, b.y);
    }
    const uint32_t xj = KS >= 8 ? uint32_t(__ldg(&xtab[4 * j])) : xo[KS >= 8 ? 0 : j];
#pragma unroll
    for (int u = 0; u < UNR; ++u) tile[sb[u] ^ xj] = make_double2(d0[u], d1[u]);
  }
}

}  // namespace hq
