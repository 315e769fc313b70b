// hq_emu.cpp -- TEST INFRASTRUCTURE ONLY (built into libhq_emu.so, never loaded by the
// product).  A host loop plays the threads of one CTA and runs the very same phase
// functions (hq_tile.cuh) and the very same planner (hq_plan.cpp) as the CUDA build, so the
// CPU-only test-suite can check the tile/lane/swizzle index math, the fusion planner and
// the bit-permutation passes against the oracle without a GPU.  It is NOT a fallback: the
// Python package never imports it and hybridq_b200 fails loudly without the CUDA library.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "hq_plan.h"
#include "hq_tile.cuh"

namespace {

void emu_fast_slot(float4* tile, int s, const HqGateDesc* g, const HqPassHeader& ph, int Tu, int tid) {
  hq::gate_fast_f32<-1, 3>(tile, hq::load_stream_regs(g, tid), ph, uint32_t(s), Tu, tid);
}
void emu_fast_slot(double2*, int, const HqGateDesc*, const HqPassHeader&, int, int) {}

template <int KK>
void emu_rowpair_k(double2* tile, const HqGateDesc* g, const double2* U, int Tu) {
  std::vector<hq::RowPairRegs<KK>> regs(HQ_THREADS);
  for (int tid = 0; tid < HQ_THREADS; ++tid) {
    hq::rowpair_load_rows<KK>(regs[size_t(tid)], U, tid);
    hq::rowpair_load_offsets<KK>(regs[size_t(tid)], g);
  }
  const uint32_t niter = hq::rowpair_iters(Tu, KK);
  for (uint32_t it = 0; it < niter; ++it) {
    std::vector<double2> o0(HQ_THREADS), o1(HQ_THREADS);
    std::vector<uint32_t> s0(HQ_THREADS), s1(HQ_THREADS);
    std::vector<char> ok(HQ_THREADS);
    for (int tid = 0; tid < HQ_THREADS; ++tid)
      ok[size_t(tid)] = hq::rowpair_compute<KK>(tile, g, regs[size_t(tid)], Tu, tid, it, o0[size_t(tid)],
                                                o1[size_t(tid)], s0[size_t(tid)], s1[size_t(tid)]);
    for (int tid = 0; tid < HQ_THREADS; ++tid)
      if (ok[size_t(tid)]) {
        tile[s0[size_t(tid)]] = o0[size_t(tid)];
        tile[s1[size_t(tid)]] = o1[size_t(tid)];
      }
  }
}
void emu_rowpair(double2* tile, const HqGateDesc* g, const unsigned char* prog, int Tu) {
  const double2* U = reinterpret_cast<const double2*>(prog + g->mat_off);
  if (g->k == 2) emu_rowpair_k<2>(tile, g, U, Tu);
  else emu_rowpair_k<3>(tile, g, U, Tu);
}
void emu_rowpair(float4*, const HqGateDesc*, const unsigned char*, int) {}

// ---- tensor-core gates: the warp-collective mma.sync is modelled on whole fragment sets --------
// (PTX fragment layouts as listed in hq_mma.cuh; tools/microbench_mma.cu checks them on the GPU).
inline float tf32_trunc(float x) {
  uint32_t b;
  memcpy(&b, &x, 4);
  b &= 0xffffe000u;
  memcpy(&x, &b, 4);
  return x;
}

// D(16x8) += A(16x8) * B(8x8), operands truncated to TF32, fp32 accumulation
void emu_mma_tf32(float (*d)[4], const float (*a)[4], const float (*b)[2]) {   // [lane][reg]
  float A[16][8], B[8][8];
  for (int lane = 0; lane < 32; ++lane) {
    const int g = lane >> 2, t = lane & 3;
    A[g][t] = tf32_trunc(a[lane][0]); A[g + 8][t] = tf32_trunc(a[lane][1]);
    A[g][t + 4] = tf32_trunc(a[lane][2]); A[g + 8][t + 4] = tf32_trunc(a[lane][3]);
    B[t][g] = tf32_trunc(b[lane][0]); B[t + 4][g] = tf32_trunc(b[lane][1]);
  }
  for (int lane = 0; lane < 32; ++lane) {
    const int g = lane >> 2, t = lane & 3;
    const int rows[4] = {g, g, g + 8, g + 8}, cols[4] = {2 * t, 2 * t + 1, 2 * t, 2 * t + 1};
    for (int e = 0; e < 4; ++e) {
      float acc = d[lane][e];
      for (int kk = 0; kk < 8; ++kk) acc += A[rows[e]][kk] * B[kk][cols[e]];
      d[lane][e] = acc;
    }
  }
}

// D(8x8) += A(8x4) * B(4x8) in double
void emu_dmma(double (*d)[2], const double* a, const double* b) {   // [lane]
  double A[8][4], B[4][8];
  for (int lane = 0; lane < 32; ++lane) {
    A[lane >> 2][lane & 3] = a[lane];
    B[lane & 3][lane >> 2] = b[lane];
  }
  for (int lane = 0; lane < 32; ++lane) {
    const int g = lane >> 2, t = lane & 3;
    for (int e = 0; e < 2; ++e) {
      double acc = d[lane][e];
      for (int kk = 0; kk < 4; ++kk) acc += A[g][kk] * B[kk][2 * t + e];
      d[lane][e] = acc;
    }
  }
}

void emu_mma_gate(float4* tile, const HqGateDesc* g, const unsigned char* prog) {
  const int KS = (1 << g->k) / 4;
  const float4* bfr = reinterpret_cast<const float4*>(prog + g->mat_off);
  float2* amps = reinterpret_cast<float2*>(tile);
  for (uint32_t warp = 0; warp < g->mma_warps; ++warp)
    for (uint32_t it = 0; it < g->mma_n_iter; ++it) {
      std::vector<float> raw(size_t(32 * KS * 4));
      uint32_t sb[32];
      for (int lane = 0; lane < 32; ++lane) {
        const int t = lane & 3;
        sb[lane] = uint32_t(g->tbl_thread[warp * 32 + uint32_t(lane)]) ^ uint32_t(g->tbl_iter[it]);
        for (int s = 0; s < KS; ++s) {
          const uint32_t xo = g->tbl_x[t + 4 * s];
          float* a = &raw[size_t((lane * KS + s) * 4)];
          if (g->mma_amp) {
            const float2 p = amps[sb[lane] ^ xo], q = amps[sb[lane] ^ g->mma_row8 ^ xo];
            a[0] = p.x; a[1] = q.x; a[2] = p.y; a[3] = q.y;
          } else {
            const float4 v = tile[sb[lane] ^ xo];
            a[0] = v.x; a[1] = v.z; a[2] = v.y; a[3] = v.w;
          }
        }
      }
      for (int j = 0; j < KS; ++j) {
        float d[32][4] = {};
        for (int s = 0; s < KS; ++s) {
          float hi[32][4], lo[32][4], bh[32][2], bl[32][2];
          for (int lane = 0; lane < 32; ++lane) {
            for (int e = 0; e < 4; ++e) {
              const float x = raw[size_t((lane * KS + s) * 4 + e)];
              uint32_t b;                              // HQ_TF32_SPLIT 1: hi = round-to-nearest by add + mask
              memcpy(&b, &x, 4);
              b = (b + 0x1000u) & 0xffffe000u;
              memcpy(&hi[lane][e], &b, 4);
              lo[lane][e] = x - hi[lane][e];           // the tensor core truncates lo
            }
            const float4 b = bfr[(s * KS + j) * 32 + lane];
            bh[lane][0] = b.x; bh[lane][1] = b.y; bl[lane][0] = b.z; bl[lane][1] = b.w;
          }
          emu_mma_tf32(d, lo, bh);
          emu_mma_tf32(d, hi, bl);
          emu_mma_tf32(d, hi, bh);
        }
        for (int lane = 0; lane < 32; ++lane) {
          const uint32_t xj = g->tbl_x[(lane & 3) + 4 * j];
          if (g->mma_amp) {
            amps[sb[lane] ^ xj] = make_float2(d[lane][0], d[lane][1]);
            amps[sb[lane] ^ g->mma_row8 ^ xj] = make_float2(d[lane][2], d[lane][3]);
          } else {
            tile[sb[lane] ^ xj] = make_float4(d[lane][0], d[lane][1], d[lane][2], d[lane][3]);
          }
        }
      }
    }
}

void emu_mma_gate(double2* tile, const HqGateDesc* g, const unsigned char* prog) {
  const int KS = (1 << g->k) / 4;
  const double2* bfr = reinterpret_cast<const double2*>(prog + g->mat_off);
  for (uint32_t warp = 0; warp < g->mma_warps; ++warp)
    for (uint32_t it = 0; it < g->mma_n_iter; ++it) {
      std::vector<double2> x(size_t(32 * KS));
      uint32_t sb[32];
      for (int lane = 0; lane < 32; ++lane) {
        sb[lane] = uint32_t(g->tbl_thread[warp * 32 + uint32_t(lane)]) ^ uint32_t(g->tbl_iter[it]);
        for (int s = 0; s < KS; ++s) x[size_t(lane * KS + s)] = tile[sb[lane] ^ g->tbl_x[(lane & 3) + 4 * s]];
      }
      for (int j = 0; j < KS; ++j) {
        double d[32][2] = {};
        for (int s = 0; s < KS; ++s) {
          double a[32], b[32];
          for (int lane = 0; lane < 32; ++lane) { a[lane] = x[size_t(lane * KS + s)].x; b[lane] = bfr[(s * KS + j) * 32 + lane].x; }
          emu_dmma(d, a, b);
          for (int lane = 0; lane < 32; ++lane) { a[lane] = x[size_t(lane * KS + s)].y; b[lane] = bfr[(s * KS + j) * 32 + lane].y; }
          emu_dmma(d, a, b);
        }
        for (int lane = 0; lane < 32; ++lane)
          tile[sb[lane] ^ g->tbl_x[(lane & 3) + 4 * j]] = make_double2(d[lane][0], d[lane][1]);
      }
    }
}

template <typename T>
void emu_pass(typename hq::Traits<T>::Unit* state, unsigned n, const unsigned char* prog, const HqPassHeader& ph) {
  typedef typename hq::Traits<T>::Unit Unit;
  typedef typename hq::Traits<T>::Cplx Cplx;
  const int V = hq::Traits<T>::V;
  const int Tbits = int(ph.tile_bits), h = int(ph.n_high);
  const int Tu = Tbits - V, Lu = Tbits - h - V;
  const uint32_t n_units = 1u << Tu;
  std::vector<Unit> tile(n_units);
  const HqGateDesc* gates = reinterpret_cast<const HqGateDesc*>(prog + ph.gates_off);
  const uint64_t n_tiles = uint64_t(1) << (n - ph.tile_bits);
  for (uint64_t t = 0; t < n_tiles; ++t) {
    const uint64_t base_unit = hq::tile_base(t, Tbits, h, ph.high_pos) >> V;
    // fill, exactly as the kernel addresses it
    for (int tid = 0; tid < HQ_THREADS; ++tid) {
      const int npt = Tu > HQ_THREADS_LOG2 ? (1 << (Tu - HQ_THREADS_LOG2)) : (uint32_t(tid) < n_units ? 1 : 0);
      const uint64_t off_t = hq::unit_offset(uint32_t(tid), Lu, V, ph.high_pos, h);
      const uint32_t swz_t = hq::swz(uint32_t(tid));
      for (int i = 0; i < npt; ++i) tile[swz_t ^ ph.iter_swz[i]] = state[base_unit + off_t + ph.iter_off[i]];
    }
    for (uint32_t gi = 0; gi < ph.n_gates; ++gi) {
      const HqGateDesc* g = gates + gi;
      if (V == 1 && gi < 32u && ((ph.fast_mask >> gi) & 1u)) {
        const int slot = __builtin_popcount(ph.fast_mask & ((1u << gi) - 1u));
        for (int tid = 0; tid < HQ_THREADS; ++tid) emu_fast_slot(tile.data(), slot, g, ph, Tu, tid);
      } else if (g->kind == HQ_GATE_DR1) {
        for (int tid = 0; tid < HQ_THREADS; ++tid) hq::gate_dr1_dispatch(tile.data(), g, g->k, prog, g->mat_off, Tu, tid);
      } else if (g->kind == HQ_GATE_MMA) {
        emu_mma_gate(tile.data(), g, prog);
      } else if (V == 0 && g->kind == HQ_GATE_ROWPAIR) {
        emu_rowpair(tile.data(), g, prog, Tu);
      } else if (g->kind != HQ_GATE_BIG) {
        const bool low = V == 1 && g->tpos[0] == 0;
        for (int tid = 0; tid < HQ_THREADS; ++tid)
          hq::gate_small_dispatch<4>(tile.data(), g, g->k, low, prog, g->mat_off, Tu, tid);
      } else {
        const Cplx* Ut = reinterpret_cast<const Cplx*>(prog + g->mat_off);
        const int rounds = hq::big_rounds(Tbits, int(g->k));
        std::vector<hq::BigAcc<T>> acc(HQ_THREADS);
        for (int r = 0; r < rounds; ++r) {
          for (int tid = 0; tid < HQ_THREADS; ++tid)
            hq::gate_big_phaseA<T>(reinterpret_cast<const Cplx*>(tile.data()), *g, Ut, Tbits, tid, r, acc[size_t(tid)]);
          for (int tid = 0; tid < HQ_THREADS; ++tid)
            hq::gate_big_phaseB<T>(reinterpret_cast<Cplx*>(tile.data()), *g, acc[size_t(tid)]);
        }
      }
    }
    const Cplx* amps = reinterpret_cast<const Cplx*>(tile.data());
    for (int tid = 0; tid < HQ_THREADS; ++tid) {
      const int npt = Tu > HQ_THREADS_LOG2 ? (1 << (Tu - HQ_THREADS_LOG2)) : (uint32_t(tid) < n_units ? 1 : 0);
      const uint64_t off_t = hq::unit_offset(uint32_t(tid), Lu, V, ph.high_pos, h);
      const uint32_t swz_t = hq::swz(uint32_t(tid));
      for (int i = 0; i < npt; ++i) {
        Unit out;
        if (!ph.has_perm) {
          out = tile[swz_t ^ ph.iter_swz[i]];
        } else {
          const uint32_t c = uint32_t(tid) + (uint32_t(i) << HQ_THREADS_LOG2);
          Cplx o[1 << V];
          for (uint32_t e = 0; e < (1u << V); ++e)
            o[e] = amps[hq::amp_slot<T>(hq::perm_src((c << V) | e, ph.perm, Tbits))];
          out = hq::make_unit(o);
        }
        state[base_unit + off_t + ph.iter_off[i]] = out;
      }
    }
  }
}

void emu_plan(hq::Plan& plan, void* state) {
  for (const hq::PassInfo& pi : plan.passes) {
    if (pi.header.n_gates == 0 && !pi.header.has_perm) continue;
    if (plan.dtype == HQ_DTYPE_C64)
      emu_pass<float>(static_cast<float4*>(state), plan.n_qubits, plan.program.data(), pi.header);
    else
      emu_pass<double>(static_cast<double2*>(state), plan.n_qubits, plan.program.data(), pi.header);
  }
}

hq::PlanOptions make_opts(const int* o) {
  hq::PlanOptions p;
  if (o) {
    p.tile_bits = o[0];
    p.min_run_bits = o[1];
    p.fuse = o[2];
    p.max_gates_per_pass = o[3];
    p.lookahead = o[4];
    p.merge_max_k = o[5];
    p.merge_pass_cost = o[6];
    p.fast_slots = o[7];
    p.mma_min_k = o[8];
  }
  return p;
}

}  // namespace

extern "C" {

// opts = {tile_bits, min_run_bits, fuse, max_gates_per_pass, lookahead, merge_max_k, merge_pass_cost,
// fast_slots, mma_min_k} or NULL.
// info_out (optional, >= 4 ints) receives {n_passes, n_gates, n_kernel_gates, n_mma_gates}.
int hq_emu_run_circuit(int dtype, unsigned n, unsigned n_gates, const unsigned* ks, const unsigned* pos_flat,
                       const double* U_flat, const int* opts, void* state_interleaved, int* info_out,
                       char* err, int err_len) {
  std::vector<hq::GateIn> gates(n_gates);
  size_t po = 0, uo = 0;
  for (unsigned g = 0; g < n_gates; ++g) {
    const unsigned k = ks[g];
    gates[g].k = k;
    gates[g].pos.assign(pos_flat + po, pos_flat + po + k);
    const size_t e = size_t(1) << (2 * k);
    gates[g].U.resize(e);
    for (size_t i = 0; i < e; ++i) gates[g].U[i] = std::complex<double>(U_flat[uo + 2 * i], U_flat[uo + 2 * i + 1]);
    po += k;
    uo += 2 * e;
  }
  hq::Plan plan;
  if (hq::plan_build(plan, dtype, n, gates, make_opts(opts))) {
    if (err) snprintf(err, size_t(err_len), "%s", plan.error.c_str());
    return 1;
  }
  emu_plan(plan, state_interleaved);
  if (info_out) {
    info_out[0] = int(plan.passes.size());
    info_out[1] = int(plan.n_gates);
    info_out[2] = int(plan.n_kernel_gates);
    int n_mma = 0;
    for (unsigned i = 0; i < plan.n_kernel_gates; ++i)
      n_mma += reinterpret_cast<const HqGateDesc*>(plan.program.data())[i].kind == HQ_GATE_MMA;
    info_out[3] = n_mma;
  }
  return 0;
}

int hq_emu_bitperm(int dtype, unsigned n, const unsigned* perm, const int* opts, void* state_interleaved,
                   int* info_out, char* err, int err_len) {
  hq::Plan plan;
  std::vector<unsigned> pf(perm, perm + n);
  if (hq::plan_build_bitperm(plan, dtype, n, pf, make_opts(opts))) {
    if (err) snprintf(err, size_t(err_len), "%s", plan.error.c_str());
    return 1;
  }
  emu_plan(plan, state_interleaved);
  if (info_out) info_out[0] = int(plan.passes.size());
  return 0;
}

// Shared-memory bank model of the gate loops (host-logic tests): for every kernel matrix of the plan, the
// number of wavefronts its 16-byte accesses need per warp instruction against the ideal of 4 (a quarter-warp
// is served in one wavefront when its 8 lanes hit 8 distinct 16-byte bank groups = slot & 7; the complex64
// amplitude path issues 8-byte accesses, served per half-warp of 16 lanes over 16 bank pairs).
// out[3 g] = gate kind, out[3 g + 1] = ideal wavefronts, out[3 g + 2] = modelled wavefronts (first row set /
// first iteration of every warp, all k-steps / group members).  Returns the number of gates or -1.
int hq_emu_bank_model(int dtype, unsigned n, unsigned n_gates, const unsigned* ks, const unsigned* pos_flat,
                      const int* opts, unsigned* out, int out_len) {
  std::vector<hq::GateIn> gates(n_gates);
  size_t po = 0;
  for (unsigned g = 0; g < n_gates; ++g) {
    const unsigned k = ks[g];
    gates[g].k = k;
    gates[g].pos.assign(pos_flat + po, pos_flat + po + k);
    gates[g].U.assign(size_t(1) << (2 * k), std::complex<double>(0, 0));
    for (size_t i = 0; i < (size_t(1) << k); ++i) gates[g].U[i * ((size_t(1) << k) + 1)] = 1.0;
    po += k;
  }
  hq::Plan plan;
  if (hq::plan_build(plan, dtype, n, gates, make_opts(opts))) return -1;
  if (int(plan.n_kernel_gates) * 3 > out_len) return -1;
  const HqGateDesc* gd = reinterpret_cast<const HqGateDesc*>(plan.program.data());
  for (unsigned gi = 0; gi < plan.n_kernel_gates; ++gi) {
    const HqGateDesc& g = gd[gi];
    unsigned ideal = 0, actual = 0;
    auto quarter = [&](const uint32_t* slots) {       // 8 lanes, 16-byte accesses
      unsigned cnt[8] = {0};
      unsigned worst = 0;
      for (int l = 0; l < 8; ++l) worst = std::max(worst, ++cnt[slots[l] & 7u]);
      ideal += 1;
      actual += worst;
    };
    auto half = [&](const uint32_t* aslots) {         // 16 lanes, 8-byte accesses: bank pair = amplitude slot & 15
      unsigned cnt[16] = {0};
      unsigned worst = 0;
      for (int l = 0; l < 16; ++l) worst = std::max(worst, ++cnt[aslots[l] & 15u]);
      ideal += 1;
      actual += worst;
    };
    if (g.kind == HQ_GATE_MMA) {
      const int KS = (1 << g.k) / 4;
      for (uint32_t warp = 0; warp < g.mma_warps; ++warp)
        for (int s = 0; s < KS; ++s) {
          uint32_t slots[32];
          for (int lane = 0; lane < 32; ++lane)
            slots[lane] = uint32_t(g.tbl_thread[warp * 32 + uint32_t(lane)]) ^ uint32_t(g.tbl_x[(lane & 3) + 4 * s]);
          if (g.mma_amp) {
            for (int h = 0; h < 2; ++h) {
              half(slots + 16 * h);
              uint32_t other[16];
              for (int l = 0; l < 16; ++l) other[l] = slots[16 * h + l] ^ g.mma_row8;
              half(other);
            }
          } else {
            for (int q = 0; q < 4; ++q) quarter(slots + 8 * q);
          }
        }
    } else if (g.kind == HQ_GATE_SMALL) {
      const bool low = dtype == HQ_DTYPE_C64 && g.tpos[0] == 0;
      const int KK = int(g.k) - (low ? 1 : 0);
      for (int m = 0; m < (1 << KK); ++m)
        for (int w = 0; w < HQ_THREADS / 8; ++w) {
          uint32_t slots[8];
          for (int l = 0; l < 8; ++l) slots[l] = uint32_t(g.tbl_thread[w * 8 + l]) ^ uint32_t(g.tbl_x[m]);
          quarter(slots);
        }
    }
    out[3 * gi] = g.kind;
    out[3 * gi + 1] = ideal;
    out[3 * gi + 2] = actual;
  }
  return int(plan.n_kernel_gates);
}

// planner introspection for host-logic tests: passes as flat records
// {tile_bits, n_high, n_kernel_gates, has_perm, n_ids, high_pos..., gate ids...}; returns words written or -1.
int hq_emu_plan_dump(int dtype, unsigned n, unsigned n_gates, const unsigned* ks, const unsigned* pos_flat,
                     const int* opts, unsigned* out, int out_len) {
  std::vector<hq::GateIn> gates(n_gates);
  size_t po = 0;
  for (unsigned g = 0; g < n_gates; ++g) {
    const unsigned k = ks[g];
    gates[g].k = k;
    gates[g].pos.assign(pos_flat + po, pos_flat + po + k);
    gates[g].U.assign(size_t(1) << (2 * k), std::complex<double>(0, 0));
    po += k;
  }
  hq::Plan plan;
  if (hq::plan_build(plan, dtype, n, gates, make_opts(opts))) return -1;
  int w = 0;
  for (const hq::PassInfo& pi : plan.passes) {
    const int need = 5 + int(pi.header.n_high) + int(pi.gate_ids.size());
    if (w + need > out_len) return -1;
    out[w++] = pi.header.tile_bits;
    out[w++] = pi.header.n_high;
    out[w++] = pi.header.n_gates;
    out[w++] = pi.header.has_perm;
    out[w++] = unsigned(pi.gate_ids.size());
    for (unsigned i = 0; i < pi.header.n_high; ++i) out[w++] = pi.header.high_pos[i];
    for (unsigned id : pi.gate_ids) out[w++] = id;
  }
  return w;
}

// the tcgen05 operand blocks of a 2^k x 2^k matrix (hq_plan.cpp umma_pack_matrix): U_flat = row-major (re, im) doubles,
// hi / lo receive (2 * 2^k)^2 floats each
void hq_emu_umma_pack(const double* U_flat, unsigned k, float* hi, float* lo) {
  const size_t dim = size_t(1) << k;
  std::vector<std::complex<double>> U(dim * dim);
  for (size_t e = 0; e < dim * dim; ++e) U[e] = std::complex<double>(U_flat[2 * e], U_flat[2 * e + 1]);
  hq::umma_pack_matrix(U.data(), k, hi, lo);
}

}  // extern "C"
