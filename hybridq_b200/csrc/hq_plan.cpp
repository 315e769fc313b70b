// hq_plan.cpp -- see hq_plan.h.
#include "hq_plan.h"
#include "hq_tile.cuh"

#include <algorithm>
#include <cstring>

namespace hq {

// Defaults come from the round-1 sweep on B200 (profiles/r01/sweep_b_after_kernel_opt.jsonl):
// 64 KiB tiles (3 resident CTAs per SM) and 256 B / 512 B minimum runs.
int default_tile_bits(int dtype) { return dtype == HQ_DTYPE_C64 ? 13 : 12; }
int default_min_run_bits(int dtype) { return dtype == HQ_DTYPE_C64 ? 5 : 5; }
int default_merge_max_k(int) { return 2; }
// complex128: FP64 mma.sync (exact IEEE accumulation) beats the DFMA paths for every k >= 2 (profiles/r01).
// complex64: k <= 3 runs on the constant-bank FFMA2 slots -- the reference's own fp32 FMA arithmetic, no norm
// drift (3xTF32 on mma.sync is ~15 % faster at k = 3 but its truncating accumulator loses 4e-5 of norm^2 per 600
// gates); k >= 4, where the tile really is a dense contraction (16x16 complex and up) and
// the FMA path is 2-7x slower, goes to the tensor cores.  PlanOptions.mma_min_k overrides (0 = never, 3 = r01).
int default_mma_min_k(int dtype) { return dtype == HQ_DTYPE_C64 ? 4 : 2; }

namespace {

// Warp-closed chains (see plan_build): measured on B200 in round 2 -- 105 of the bench circuit's 112 gate-to-gate
// barriers become __syncwarp(), results identical, but no gain (128.1 vs 126.0 ms/step; a 2 x 4-gate pass 4.60 vs
// 4.13 ms, profiles/r02/sweep_ring_g.jsonl): the CTA barrier is not what limits the gate loop.  Kept as a
// compile-time experiment, off by default.
#ifndef HQ_WARP_CHAINS
#define HQ_WARP_CHAINS 0
#endif
#ifndef HQ_MERGE_SLACK
#define HQ_MERGE_SLACK 10
#endif
const int kMaxGatesPerPass = 48;   // gate-applies per pass before merging

int vbits(int dtype) { return dtype == HQ_DTYPE_C64 ? 1 : 0; }
bool vectors_independent(const std::vector<uint32_t>& v);
int max_tile_bits(int dtype) { return HQ_MAX_UNIT_BITS + vbits(dtype); }

// Largest run length L in [Lmin, T] such that the members of `bits` at or above L fit in the
// T - L high slots.  Returns -1 if there is none.
int choose_run_bits(const std::vector<unsigned>& bits_sorted, int T, int Lmin) {
  for (int L = T; L >= Lmin; --L) {
    int high = 0;
    for (unsigned b : bits_sorted) high += (int(b) >= L);
    if (high <= T - L && high <= HQ_MAX_HIGH) return L;
  }
  return -1;
}

struct Canon {               // gate with ascending positions and accordingly permuted matrix
  unsigned k;
  std::vector<unsigned> pos;
  std::vector<std::complex<double>> U;
  // "scalar + rank one" form U = lambda * 1 + u v^T (detect_dr1), e.g. a depolarizing channel as a super-operator
  bool dr1 = false;
  std::complex<double> lambda;
  std::vector<std::complex<double>> u, v;
  // sparse form (fold_dr1_scalars): lambda moved into another matrix of the plan (u already divided by it), the gate
  // is 1 + u v^T and touches only the amplitudes where u or v is non-zero
  bool dr1_folded = false;
};

// Is U = lambda * 1 + u v^T?  u and v are read off the row and the column of the largest OFF-diagonal entry
// (scale: u[i0] = 1), their two missing components off a second row / column, lambda = U_ii - u_i v_i must then be
// the same for every i, and every entry is checked against the reconstruction.  Only worth it from k = 3 up.
// Tolerance: the reconstruction must reproduce every entry to within what the kernel's own storage of the matrix
// would lose anyway -- a few fp32 ulps of the largest entry for complex64 plans (so channels whose matrices were
// built in single precision are still recognised), 1e-12 relative for complex128.
bool detect_dr1(Canon& c, int dtype) {
  c.dr1 = false;
  if (c.k < 3 || c.k > HQ_DR1_MAX_K) return false;
  const size_t dim = size_t(1) << c.k;
  const std::vector<std::complex<double>>& U = c.U;
  double scale = 0;
  for (const auto& x : U) scale = std::max(scale, std::abs(x));
  if (scale == 0) return false;
  const double tol = (dtype == HQ_DTYPE_C64 ? 6e-8 : 1e-12) * scale;
  size_t i0 = 0, j0 = 0;
  double big = 0;
  for (size_t i = 0; i < dim; ++i)
    for (size_t j = 0; j < dim; ++j)
      if (i != j && std::abs(U[i * dim + j]) > big) { big = std::abs(U[i * dim + j]); i0 = i; j0 = j; }
  std::vector<std::complex<double>> u(dim, 0.0), v(dim, 0.0);
  if (big > tol) {
    u[i0] = 1.0;
    for (size_t j = 0; j < dim; ++j)
      if (j != i0) v[j] = U[i0 * dim + j];
    for (size_t i = 0; i < dim; ++i)
      if (i != j0 && i != i0) u[i] = U[i * dim + j0] / v[j0];
    // u[j0] from another column, v[i0] from another row (the largest usable ones)
    size_t j1 = dim, i1 = dim;
    for (size_t j = 0; j < dim; ++j)
      if (j != j0 && j != i0 && (j1 == dim || std::abs(v[j]) > std::abs(v[j1]))) j1 = j;
    for (size_t i = 0; i < dim; ++i)
      if (i != i0 && i != j0 && (i1 == dim || std::abs(u[i]) > std::abs(u[i1]))) i1 = i;
    if (j1 == dim || i1 == dim) return false;
    if (std::abs(v[j1]) > tol) u[j0] = U[j0 * dim + j1] / v[j1];
    else {
      for (size_t j = 0; j < dim; ++j)
        if (j != j0 && std::abs(U[j0 * dim + j]) > tol) return false;      // row j0 has entries a zero u[j0] cannot make
    }
    if (std::abs(u[i1]) > tol) v[i0] = U[i1 * dim + i0] / u[i1];
    else {
      for (size_t i = 0; i < dim; ++i)
        if (i != i0 && std::abs(U[i * dim + i0]) > tol) return false;
    }
  }
  const std::complex<double> lam = U[0] - u[0] * v[0];
  for (size_t i = 0; i < dim; ++i)
    for (size_t j = 0; j < dim; ++j) {
      const std::complex<double> want = u[i] * v[j] + (i == j ? lam : std::complex<double>(0, 0));
      if (std::abs(U[i * dim + j] - want) > 8 * tol) return false;
    }
  c.dr1 = true;
  c.lambda = lam;
  c.u.swap(u);
  c.v.swap(v);
  return true;
}

bool canonicalise(const GateIn& g, unsigned n, Canon& out, std::string& err) {
  const unsigned k = g.k;
  if (g.pos.size() != k || g.U.size() != (size_t(1) << (2 * k))) {
    err = "gate has inconsistent k / pos / U sizes";
    return false;
  }
  std::vector<unsigned> order(k);
  for (unsigned i = 0; i < k; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](unsigned a, unsigned b) { return g.pos[a] < g.pos[b]; });
  out.k = k;
  out.pos.resize(k);
  for (unsigned i = 0; i < k; ++i) {
    out.pos[i] = g.pos[order[i]];
    if (out.pos[i] >= n) { err = "gate position out of range"; return false; }
    if (i && out.pos[i] == out.pos[i - 1]) { err = "gate has duplicate positions"; return false; }
  }
  const size_t dim = size_t(1) << k;
  // new matrix bit i is old matrix bit order[i]
  std::vector<size_t> map(dim);
  for (size_t a = 0; a < dim; ++a) {
    size_t o = 0;
    for (unsigned i = 0; i < k; ++i) o |= ((a >> i) & 1u) << order[i];
    map[a] = o;
  }
  out.U.resize(dim * dim);
  for (size_t a = 0; a < dim; ++a)
    for (size_t b = 0; b < dim; ++b) out.U[a * dim + b] = g.U[map[a] * dim + map[b]];
  return true;
}

// Choose the tile of a pass: pad the mandatory high bits up to exactly T - L entries.
void make_tile(const std::vector<unsigned>& bits_sorted, int T, int L, unsigned n, HqPassHeader& ph) {
  std::vector<unsigned> high;
  for (unsigned b : bits_sorted)
    if (int(b) >= L) high.push_back(b);
  const int want = T - L;
  for (unsigned b = unsigned(L); int(high.size()) < want && b < n; ++b)
    if (std::find(high.begin(), high.end(), b) == high.end()) high.push_back(b);
  std::sort(high.begin(), high.end());
  memset(&ph, 0, sizeof(ph));
  ph.tile_bits = uint32_t(L + int(high.size()));
  ph.n_high = uint32_t(high.size());
  for (size_t i = 0; i < high.size(); ++i) ph.high_pos[i] = uint8_t(high[i]);
}

// Fill/drain addressing tables of a pass (see HqPassHeader).
void make_iter_tables(HqPassHeader& ph, int V) {
  const int Tu = int(ph.tile_bits) - V;
  const int Lu = int(ph.tile_bits) - int(ph.n_high) - V;
  const int npt = Tu > HQ_THREADS_LOG2 ? (1 << (Tu - HQ_THREADS_LOG2)) : 1;
  for (int i = 0; i < HQ_MAX_PER_THREAD; ++i) {
    const uint32_t c = i < npt ? (uint32_t(i) << HQ_THREADS_LOG2) : 0u;
    ph.iter_off[i] = i < npt && Tu > HQ_THREADS_LOG2 ? unit_offset(c, Lu, V, ph.high_pos, int(ph.n_high)) : 0;
    ph.iter_swz[i] = i < npt && Tu > HQ_THREADS_LOG2 ? swz(c) : 0u;
  }
}

// Lane tables of a register-path gate (see HqGateDesc).
void make_lane_tables(HqGateDesc& gd, int Tu, int V) {
  const bool low = V == 1 && gd.tpos[0] == 0;
  const int KK = int(gd.k) - (low ? 1 : 0);
  const int nq = Tu - KK;
  const int tb = nq < HQ_THREADS_LOG2 ? nq : HQ_THREADS_LOG2;
  for (int t = 0; t < HQ_THREADS; ++t)
    gd.tbl_thread[t] = uint16_t(swz(scatter_bits(uint32_t(t) & ((1u << tb) - 1u), gd.q, 0, tb)));
  const int niter = 1 << (nq - tb);
  for (int it = 0; it < 16; ++it)
    gd.tbl_iter[it] = it < niter ? uint16_t(swz(scatter_bits(uint32_t(it), gd.q, tb, nq))) : uint16_t(0);
  uint8_t upos[16] = {0};
  for (int i = 0; i < KK; ++i) upos[i] = uint8_t(gd.tpos[i + (low ? 1 : 0)] - V);
  for (int m = 0; m < 16; ++m)
    gd.tbl_x[m] = m < (1 << KK) ? uint16_t(swz(uint32_t(deposit(uint32_t(m), upos, KK)))) : uint16_t(0);
}

// Lane tables of the complex128 row-pair scheme (see HqGateDesc).
void make_rowpair_tables(HqGateDesc& gd, int Tu) {
  const int KK = int(gd.k);
  if (KK < 2 || KK > 3) return;
  const int lrp = KK - 1;
  const int nq = Tu - KK;
  const int gbits = HQ_THREADS_LOG2 - lrp;                 // group slots per iteration = 2^gbits
  const int tb = nq < gbits ? nq : gbits;
  for (int gs = 0; gs < HQ_THREADS; ++gs)
    gd.tbl_rthread[gs] = uint16_t(swz(scatter_bits(uint32_t(gs) & ((1u << tb) - 1u), gd.q, 0, tb)));
  const int niter = 1 << (nq - tb);
  for (int it = 0; it < 16; ++it)
    gd.tbl_riter[it] = it < niter ? uint16_t(swz(scatter_bits(uint32_t(it), gd.q, tb, nq))) : uint16_t(0);
}

// ---- tensor-core gates (HQ_GATE_MMA, device side in hq_mma.cuh) -------------------------------
// Lane (g, t) = (lane >> 2, lane & 3) of a warp owns row g of a row set (8 rows; the complex64
// amplitude path adds row g + 8) and the amplitudes / units m = t + 4 s of that row.  The two
// matrix-index bits carried by t ("lane bits") and the lowest row bits are chosen so that a
// quarter-warp of 16-byte accesses (half-warp of 8-byte accesses on the amplitude path) hits
// distinct bank groups under swz: their unit bits must have distinct residues mod 3.
struct MmaLayout {
  bool amp = false;                 // complex64 amplitude granularity
  int forder[HQ_MMA_MAX_K] = {0};   // fragment-index bit b <-> matrix-index bit forder[b]
  std::vector<unsigned> rows;       // row bits in work-item order (local unit bits, or amplitude bits if amp)
  int row_lane_bits = 3;            // 3 (8 rows) or 4 (16 rows: g and the g + 8 half)
  int warp_bits = 0, iter_bits = 0;
};

bool mma_layout(const uint8_t* tpos, int k, int Tbits, int V, MmaLayout& L) {
  if (k < 2 || k > HQ_MMA_MAX_K) return false;
  L.amp = V == 1 && tpos[0] == 0;
  const int bits = L.amp ? Tbits : Tbits - V;          // index bits of the granularity in use
  std::vector<int> tb(size_t(k), 0);                   // target bits at that granularity
  for (int i = 0; i < k; ++i) tb[size_t(i)] = L.amp ? int(tpos[i]) : int(tpos[i]) - V;
  auto vec = [&](int b) { return swz_vec(L.amp ? b - 1 : b); };          // bank vector of the bit's unit bit
  // lane bits
  int c0 = 0, c1 = 1;
  if (!L.amp) {
    bool found = false;
    for (int i = 0; i < k && !found; ++i)
      for (int j = i + 1; j < k && !found; ++j)
        if (vec(tb[size_t(i)]) != vec(tb[size_t(j)])) { c0 = i; c1 = j; found = true; }
  }
  int nf = 0;
  L.forder[nf++] = c0;
  L.forder[nf++] = c1;
  for (int i = 0; i < k; ++i)
    if (i != c0 && i != c1) L.forder[nf++] = i;
  // row bits
  std::vector<unsigned> freeb;
  std::vector<bool> is_t(size_t(bits), false);
  for (int i = 0; i < k; ++i) is_t[size_t(tb[size_t(i)])] = true;
  for (int b = 0; b < bits; ++b)
    if (!is_t[size_t(b)]) freeb.push_back(unsigned(b));
  L.row_lane_bits = L.amp ? 4 : 3;
  if (int(freeb.size()) < L.row_lane_bits) return false;
  // unit path: a quarter-warp spans (t0, t1, g0); amplitude path: a half-warp of 8-byte accesses spans
  // (t0 = the half of the unit, t1, g0, g1) -- the unit-level lane bits must have independent vectors
  std::vector<uint32_t> vecs;
  if (L.amp) vecs.push_back(vec(tb[size_t(c1)]));
  else { vecs.push_back(vec(tb[size_t(c0)])); vecs.push_back(vec(tb[size_t(c1)])); }
  const int want = L.amp ? 2 : 1;
  for (int c = 0; c < want; ++c)
    for (size_t i = size_t(c); i < freeb.size(); ++i) {
      std::vector<uint32_t> trial = vecs;
      trial.push_back(vec(int(freeb[i])));
      if (vectors_independent(trial)) {
        const unsigned b = freeb[i];
        freeb.erase(freeb.begin() + long(i));
        freeb.insert(freeb.begin() + c, b);
        vecs.swap(trial);
        break;
      }
    }
  L.rows = freeb;
  const int nrow = int(freeb.size());
  L.warp_bits = std::min(HQ_THREADS_LOG2 - 5, nrow - L.row_lane_bits);
  L.iter_bits = nrow - L.row_lane_bits - L.warp_bits;
  return L.iter_bits <= 4;
}

uint32_t mma_slot(const MmaLayout& L, uint32_t x) {
  return L.amp ? ((swz(x >> 1) << 1) | (x & 1u)) : swz(x);
}

size_t mma_mat_index(const MmaLayout& L, int k, uint32_t m) {
  size_t r = 0;
  for (int b = 0; b < k; ++b) r |= size_t((m >> b) & 1u) << L.forder[b];
  return r;
}

void make_mma_tables(HqGateDesc& gd, const MmaLayout& L, int V) {
  const int k = int(gd.k);
  auto scat = [&](uint32_t w, int from) {
    uint32_t u = 0;
    for (size_t i = 0; size_t(from) + i < L.rows.size(); ++i) u |= ((w >> i) & 1u) << L.rows[size_t(from) + i];
    return u;
  };
  const uint32_t wmask = (1u << L.warp_bits) - 1u;
  for (int tid = 0; tid < HQ_THREADS; ++tid) {
    const uint32_t lane = uint32_t(tid) & 31u, warp = uint32_t(tid) >> 5, g = lane >> 2;
    gd.tbl_thread[tid] = uint16_t(mma_slot(L, scat(g | ((warp & wmask) << L.row_lane_bits), 0)));
  }
  gd.mma_n_iter = 1u << L.iter_bits;
  gd.mma_warps = 1u << L.warp_bits;
  gd.mma_amp = L.amp ? 1u : 0u;
  for (uint32_t it = 0; it < 16; ++it)
    gd.tbl_iter[it] = it < gd.mma_n_iter ? uint16_t(mma_slot(L, scat(it, L.row_lane_bits + L.warp_bits))) : uint16_t(0);
  for (uint32_t m = 0; m < 64; ++m) {
    uint32_t x = 0;
    if (m < (1u << k))
      for (int b = 0; b < k; ++b) x |= ((m >> b) & 1u) << (L.amp ? int(gd.tpos[L.forder[b]]) : int(gd.tpos[L.forder[b]]) - V);
    gd.tbl_x[m] = uint16_t(mma_slot(L, x));
  }
  gd.mma_row8 = L.amp ? mma_slot(L, 1u << L.rows[3]) : 0u;
}

float tf32_rna_host(float x) {
  uint32_t b;
  memcpy(&b, &x, 4);
  b = (b + 0x1000u) & 0xffffe000u;
  float r;
  memcpy(&r, &b, 4);
  return r;
}

size_t mma_frag_bytes(unsigned k) { return size_t(32) << (2 * k); }   // KS^2 blocks x 32 lanes x 16 B

// B fragments (see hq_mma.cuh): block (k-step s, n-tile j), lane (g, t): input amplitude m = 4 s + t,
// output amplitude 4 j + (g >> 1), output component re / im = g & 1;
//   b0 = B[re(m)][n], b1 = B[im(m)][n]   with   B = [[Ur, Ui], [-Ui, Ur]]  (rows re/im in, columns re/im out)
void write_mma_fragments(std::vector<unsigned char>& prog, size_t off, const Canon& c, const MmaLayout& L, int dtype) {
  const int k = int(c.k);
  const size_t dim = size_t(1) << k;
  const int KS = int(dim / 4);
  for (int s = 0; s < KS; ++s)
    for (int j = 0; j < KS; ++j)
      for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3, odd = g & 1;
        const size_t mi = mma_mat_index(L, k, uint32_t(4 * s + t)), mo = mma_mat_index(L, k, uint32_t(4 * j + (g >> 1)));
        const std::complex<double> u = c.U[mo * dim + mi];
        const double b0 = odd ? u.imag() : u.real(), b1 = odd ? u.real() : -u.imag();
        unsigned char* dst = prog.data() + off + (size_t(s * KS + j) * 32 + size_t(lane)) * 16;
        if (dtype == HQ_DTYPE_C64) {
          const float h0 = tf32_rna_host(float(b0)), h1 = tf32_rna_host(float(b1));
          const float f[4] = {h0, h1, tf32_rna_host(float(b0 - double(h0))), tf32_rna_host(float(b1 - double(h1)))};
          memcpy(dst, f, 16);
        } else {
          const double d[2] = {b0, b1};
          memcpy(dst, d, 16);
        }
      }
}

int local_bit(const HqPassHeader& ph, unsigned global_bit) {
  const int L = int(ph.tile_bits) - int(ph.n_high);
  if (int(global_bit) < L) return int(global_bit);
  for (unsigned i = 0; i < ph.n_high; ++i)
    if (ph.high_pos[i] == global_bit) return L + int(i);
  return -1;
}

// true when the 3-bit bank vectors in `v` (see swz_vec) are linearly independent over GF(2)
bool vectors_independent(const std::vector<uint32_t>& v) {
  for (size_t m = 1; m < (size_t(1) << v.size()); ++m) {
    uint32_t x = 0;
    for (size_t i = 0; i < v.size(); ++i)
      if ((m >> i) & 1u) x ^= v[i];
    if (x == 0) return false;
  }
  return true;
}

// Order the free unit bits so that the three lowest work-item bits land on bits with linearly
// independent bank vectors (conflict-free quarter-warps under swz), lowest bits first otherwise.
std::vector<unsigned> lane_order(const std::vector<unsigned>& free_sorted) {
  std::vector<unsigned> first;
  std::vector<uint32_t> vecs;
  std::vector<bool> taken(free_sorted.size(), false);
  for (size_t i = 0; i < free_sorted.size() && first.size() < 3; ++i) {
    std::vector<uint32_t> trial = vecs;
    trial.push_back(swz_vec(int(free_sorted[i])));
    if (vectors_independent(trial)) {
      vecs.swap(trial);
      taken[i] = true;
      first.push_back(free_sorted[i]);
    }
  }
  std::vector<unsigned> out = first;
  for (size_t i = 0; i < free_sorted.size(); ++i)
    if (!taken[i]) out.push_back(free_sorted[i]);
  return out;
}

template <typename R>
void write_matrix(std::vector<unsigned char>& prog, size_t off, const Canon& c, bool transposed) {
  const size_t dim = size_t(1) << c.k;
  R* dst = reinterpret_cast<R*>(prog.data() + off);
  for (size_t i = 0; i < dim; ++i)
    for (size_t j = 0; j < dim; ++j) {
      const std::complex<double> v = c.U[i * dim + j];
      const size_t e = transposed ? (j * dim + i) : (i * dim + j);
      dst[2 * e] = R(v.real());
      dst[2 * e + 1] = R(v.imag());
    }
}

// ---- in-pass merging -------------------------------------------------------------------
// Embed `c` (ascending positions) into the ascending superset `pos_u`.
std::vector<std::complex<double>> embed(const Canon& c, const std::vector<unsigned>& pos_u) {
  const unsigned ku = unsigned(pos_u.size());
  const size_t dim = size_t(1) << ku;
  std::vector<unsigned> slot(c.k);          // matrix bit i of c -> matrix bit slot[i] of the union
  for (unsigned i = 0; i < c.k; ++i)
    slot[i] = unsigned(std::find(pos_u.begin(), pos_u.end(), c.pos[i]) - pos_u.begin());
  size_t cmask = 0;
  for (unsigned i = 0; i < c.k; ++i) cmask |= size_t(1) << slot[i];
  auto sub = [&](size_t a) {
    size_t r = 0;
    for (unsigned i = 0; i < c.k; ++i) r |= ((a >> slot[i]) & 1u) << i;
    return r;
  };
  const size_t cd = size_t(1) << c.k;
  std::vector<std::complex<double>> out(dim * dim, std::complex<double>(0, 0));
  for (size_t a = 0; a < dim; ++a)
    for (size_t b = 0; b < dim; ++b)
      if ((a & ~cmask) == (b & ~cmask)) out[a * dim + b] = c.U[sub(a) * cd + sub(b)];
  return out;
}

template <typename T>
void write_dr1(std::vector<unsigned char>& prog, size_t off, const Canon& c, const uint16_t* tbl_x) {
  T* out = reinterpret_cast<T*>(prog.data() + off);
  const size_t dim = size_t(1) << c.k;
  out[0] = T(c.lambda.real());
  out[1] = T(c.lambda.imag());
  for (size_t i = 0; i < dim; ++i) {
    out[2 + 2 * i] = T(c.u[i].real());
    out[2 + 2 * i + 1] = T(c.u[i].imag());
    out[2 + 2 * dim + 2 * i] = T(c.v[i].real());
    out[2 + 2 * dim + 2 * i + 1] = T(c.v[i].imag());
  }
  // trailer of the sparse form (gate_dr1 in hq_tile.cuh): [0] = (number of listed units, folded flag), then per listed
  // unit s < 4 five complex slots: (slot offset of the unit = tbl_x[m], 0), v of its first amplitude, v of its second
  // (complex64 with a target on amplitude bit 0: a unit holds two amplitudes of the group), u likewise.  Unused
  // entries repeat unit 0 with zero u and v.
  T* tr = out + 2 + 4 * dim;
  for (int i = 0; i < 2 * (1 + 5 * 4); ++i) tr[i] = T(0);
  const bool pairs = sizeof(T) == 4 && c.pos[0] == 0;
  const size_t per = pairs ? 2 : 1, units = dim / per;
  std::vector<size_t> listed;
  for (size_t m = 0; m < units; ++m) {
    bool nzu = false;
    for (size_t e = 0; e < per; ++e)
      nzu = nzu || c.u[m * per + e] != std::complex<double>(0, 0) || c.v[m * per + e] != std::complex<double>(0, 0);
    if (nzu) listed.push_back(m);
  }
  const bool sparse = c.dr1_folded && !listed.empty() && listed.size() <= 4;
  tr[0] = T(listed.size());
  tr[1] = T(sparse ? 1 : 0);
  if (sparse)
    for (size_t s4 = 0; s4 < 4; ++s4) {
      const bool real = s4 < listed.size();
      const size_t m = real ? listed[s4] : listed[0];
      T* e = tr + 2 * (1 + 5 * s4);
      e[0] = T(tbl_x[m]);
      if (!real) continue;
      for (size_t a = 0; a < per; ++a) {
        e[2 + 2 * a] = T(c.v[m * per + a].real());
        e[3 + 2 * a] = T(c.v[m * per + a].imag());
        e[6 + 2 * a] = T(c.u[m * per + a].real());
        e[7 + 2 * a] = T(c.u[m * per + a].imag());
      }
    }
}

// first := second * first on the union of their positions (second is applied after first).
void merge_into(Canon& first, const Canon& second) {
  std::vector<unsigned> pos_u = first.pos;
  for (unsigned p : second.pos)
    if (std::find(pos_u.begin(), pos_u.end(), p) == pos_u.end()) pos_u.push_back(p);
  std::sort(pos_u.begin(), pos_u.end());
  const size_t dim = size_t(1) << pos_u.size();
  const std::vector<std::complex<double>> A = embed(second, pos_u), B = embed(first, pos_u);
  std::vector<std::complex<double>> C(dim * dim, std::complex<double>(0, 0));
  for (size_t i = 0; i < dim; ++i)
    for (size_t l = 0; l < dim; ++l) {
      const std::complex<double> a = A[i * dim + l];
      if (a == std::complex<double>(0, 0)) continue;
      for (size_t j = 0; j < dim; ++j) C[i * dim + j] += a * B[l * dim + j];
    }
  first.k = unsigned(pos_u.size());
  first.pos = pos_u;
  first.U.swap(C);
  first.dr1 = false;
}

struct Cluster {
  Canon gate;
  std::vector<unsigned> ids;       // canonical-gate indices merged in, in application order
  uint64_t mask = 0;
};

int union_k(uint64_t a, uint64_t b) { return __builtin_popcountll(a | b); }

// Greedy merging of an ordered gate list (all inside one tile).
// Cost of one kernel matrix of k qubits inside a saturated pass, in microseconds at n = 30 (complex64)
// / n = 29 (complex128), measured on B200 (profiles/r01/sweep_slope_b_breg4_k1d3.jsonl, sweep_slope_e_acc_outside.jsonl): the FMA
// paths (constant-bank FFMA2 slots for complex64 k = 2, row pairs for complex128) and the tensor-core
// path (3xTF32 / FP64 mma.sync).  Only the ratios matter.
int measured_cost(int dtype, bool mma_on, int mma_min_k, int k) {
  // complex64 FMA column, round 2: k <= 3 are the constant-bank FFMA2 slots (gate_stream_f32; k = 3 measured at
  // 1.36 ms per matrix at n = 30, profiles/r02/sweep_ring_b.jsonl), k = 4 the generic register path
  static const int c64_fma[5] = {0, 600, 620, 1360, 5280}, c64_mma[5] = {0, 600, 940, 1255, 2420};
  static const int c128_fma[5] = {0, 620, 1255, 2650, 7190}, c128_mma[5] = {0, 620, 720, 1160, 2320};
  if (k < 1) return 0;
  if (k > 4) return 1 << 30;
  const bool mma = mma_on && k >= mma_min_k;
  return dtype == HQ_DTYPE_C64 ? (mma ? c64_mma[k] : c64_fma[k]) : (mma ? c128_mma[k] : c128_fma[k]);
}

std::vector<Cluster> merge_pass(const std::vector<Canon>& canon, const std::vector<unsigned>& ids, int max_k,
                                int pass_cost, int dtype, bool mma_on, int mma_min_k) {
  std::vector<Cluster> cl;
  // pass_cost >= 0: the analytic model 4 * 2^k + pass_cost (FMA per amplitude + one shared-memory round
  // trip); pass_cost < 0: the measured table above
  auto cost = [&](int k) { return pass_cost >= 0 ? 4 * (1 << k) + pass_cost : measured_cost(dtype, mma_on, mma_min_k, k); };
  for (unsigned id : ids) {
    const Canon& g = canon[id];
    uint64_t gm = 0;
    for (unsigned p : g.pos) gm |= uint64_t(1) << p;
    int target = -1;
    // a scalar + rank-one gate is cheaper on its own than anything it could be multiplied into
    if (max_k > 0 && int(g.k) <= max_k && !g.dr1) {
      int best_gain = -1;
      for (int c = int(cl.size()) - 1; c >= 0; --c) {
        const int ku = union_k(cl[size_t(c)].mask, gm);
        if (ku <= max_k && ku <= HQ_SMALL_K && !cl[size_t(c)].gate.dr1) {
          // a merged matrix that costs up to HQ_MERGE_SLACK % more than its two parts is still taken: the merge
          // is pairwise and greedy, and a cluster that has grown to k bits absorbs every later gate on those
          // bits for free (a triangle of k = 2 gates becomes ONE k = 3 matrix only if the first pair may merge)
          const int gain = cost(int(cl[size_t(c)].gate.k)) + cost(int(g.k)) - cost(ku) + cost(ku) * HQ_MERGE_SLACK / 100;
          // prefer the cluster that shares bits with g (later ones are disjoint by construction)
          if (gain >= 0 && gain > best_gain) {
            best_gain = gain;
            target = c;
          }
        }
        if (cl[size_t(c)].mask & gm) break;       // g cannot move before a gate it overlaps
      }
    }
    if (target >= 0) {
      merge_into(cl[size_t(target)].gate, g);
      cl[size_t(target)].ids.push_back(id);
      cl[size_t(target)].mask |= gm;
    } else {
      Cluster c;
      c.gate = g;
      c.ids.push_back(id);
      c.mask = gm;
      cl.push_back(std::move(c));
    }
  }
  return cl;
}

}  // namespace

int plan_build(Plan& plan, int dtype, unsigned n, const std::vector<GateIn>& gates_in,
               const PlanOptions& opts) {
  plan.dtype = dtype;
  plan.n_qubits = n;
  plan.passes.clear();
  plan.program.clear();
  plan.error.clear();
  plan.n_gates = 0;
  const int V = vbits(dtype);
  if (dtype != HQ_DTYPE_C64 && dtype != HQ_DTYPE_C128) { plan.error = "bad dtype"; return 1; }
  if (n < unsigned(V) + 0u || n == 0 || n > 48) { plan.error = "unsupported number of qubits"; return 1; }

  int T = opts.tile_bits > 0 ? opts.tile_bits : default_tile_bits(dtype);
  T = std::min(T, max_tile_bits(dtype));
  T = std::min<int>(T, int(n));
  if (T < V) { plan.error = "tile too small"; return 1; }
  const int hard_min_run = V;                       // a unit must not straddle two runs
  int fuse_min_run = opts.min_run_bits >= 0 ? opts.min_run_bits : default_min_run_bits(dtype);
  fuse_min_run = std::max(hard_min_run, std::min(fuse_min_run, T));
  const int max_per_pass = std::min(kMaxGatesPerPass, opts.max_gates_per_pass > 0 ? opts.max_gates_per_pass : kMaxGatesPerPass);
  const size_t lookahead = opts.lookahead > 0 ? size_t(opts.lookahead) : size_t(4096);

  // canonical gates (k = 0 gates are no-ops exactly as in the reference, python_U.cpp:38-39)
  std::vector<Canon> canon;
  std::vector<unsigned> canon_id;
  for (size_t i = 0; i < gates_in.size(); ++i) {
    if (gates_in[i].k == 0) continue;
    if (gates_in[i].k > HQ_MAX_K) { plan.error = "gate with k > HQ_MAX_K"; return 1; }
    Canon c;
    if (!canonicalise(gates_in[i], n, c, plan.error)) return 1;
    std::vector<unsigned> b = c.pos;
    if (choose_run_bits(b, T, hard_min_run) < 0) { plan.error = "gate does not fit in a tile"; return 1; }
    detect_dr1(c, dtype);
    canon.push_back(std::move(c));
    canon_id.push_back(unsigned(i));
  }
  plan.n_gates = unsigned(canon.size());

  // ---- greedy fusion: a gate joins the open pass if it touches no bit a deferred gate
  // touches and the union of target bits still fits in a tile with runs >= fuse_min_run.
  // A dense complex64 gate of UMMA_MIN_K .. UMMA_MAX_K qubits is cheaper as a pass of its own on the tcgen05 kernel
  // (hq_umma.cuh: 2.7 / 3.6 / 3.9 ms at n = 30 for k = 4 / 5 / 6 whatever was multiplied into it) than as one more
  // matrix of a tile pass on the mma.sync path (5.3 ms and up per matrix): such a "solo" gate opens its own pass, which
  // then only admits gates acting inside its qubits -- they are multiplied into its matrix below (so the rule is off
  // when the caller switched merging off, merge_max_k = 0).
  const int fuse_mma_min_k = opts.mma_min_k < 0 ? default_mma_min_k(dtype) : opts.mma_min_k;
  auto is_solo = [&](const Canon& g) {
    return opts.fuse && opts.merge_max_k != 0 && dtype == HQ_DTYPE_C64 && fuse_mma_min_k >= 2 && int(g.k) >= fuse_mma_min_k && g.k >= UMMA_MIN_K &&
           g.k <= UMMA_MAX_K && g.k <= unsigned(HQ_MMA_MAX_K) && !g.dr1 && n >= g.k + UMMA_ROW_BITS;
  };
  std::vector<bool> done(canon.size(), false);
  size_t first = 0;
  struct Draft { std::vector<unsigned> ids; std::vector<unsigned> bits; bool solo = false; };
  std::vector<Draft> drafts;
  while (first < canon.size()) {
    if (done[first]) { ++first; continue; }
    Draft d;
    uint64_t blocked = 0, solo_mask = 0;
    std::vector<unsigned> bits;
    size_t scanned_blocked = 0;
    for (size_t i = first; i < canon.size(); ++i) {
      if (done[i]) continue;
      uint64_t mask = 0;
      for (unsigned p : canon[i].pos) mask |= uint64_t(1) << p;
      bool take = !(mask & blocked) && int(d.ids.size()) < max_per_pass;
      if (take && !d.ids.empty()) {
        if (d.solo) take = !(mask & ~solo_mask) && !canon[i].dr1;
        else if (is_solo(canon[i])) take = false;
      }
      if (take && d.ids.empty() && is_solo(canon[i])) {
        d.solo = true;
        solo_mask = mask;
      }
      if (take) {
        std::vector<unsigned> u = bits;
        for (unsigned p : canon[i].pos)
          if (std::find(u.begin(), u.end(), p) == u.end()) u.push_back(p);
        std::sort(u.begin(), u.end());
        const int minrun = d.ids.empty() ? hard_min_run : fuse_min_run;
        if (choose_run_bits(u, T, minrun) >= 0) {
          bits.swap(u);
          d.ids.push_back(unsigned(i));
          done[i] = true;
        } else {
          take = false;
        }
      }
      if (!take) {
        blocked |= mask;
        if (++scanned_blocked >= lookahead) break;
      }
      if (!opts.fuse && !d.ids.empty()) break;
      if (blocked == ((n >= 64) ? ~uint64_t(0) : ((uint64_t(1) << n) - 1))) break;
    }
    d.bits = bits;
    drafts.push_back(std::move(d));
  }

  // ---- merge inside each pass, then serialise
  const int mma_min_k = opts.mma_min_k < 0 ? default_mma_min_k(dtype) : opts.mma_min_k;
  const bool mma_on = mma_min_k >= 2;
  // Merging is decided by the measured per-matrix costs (measured_cost above) unless the caller asks for
  // the analytic model with merge_pass_cost >= 0.  Without the tensor-core path nothing above k = 2 pays.
  const int merge_default = mma_on ? HQ_SMALL_K : default_merge_max_k(dtype);
  const int merge_max_k = opts.merge_max_k < 0 ? merge_default : std::min(opts.merge_max_k, HQ_SMALL_K);
  const int merge_pass_cost = opts.merge_pass_cost;
  std::vector<std::vector<Cluster>> merged(drafts.size());
  std::vector<PassInfo> infos(drafts.size());
  struct GateLayout { bool mma = false; bool dr1 = false; MmaLayout L; uint8_t tpos[16] = {0}; size_t bytes = 0; };
  std::vector<std::vector<GateLayout>> layouts(drafts.size());
  size_t total_gates = 0;
  size_t mat_bytes = 0;
  const size_t esz = dtype == HQ_DTYPE_C64 ? 8 : 16;
  for (size_t d = 0; d < drafts.size(); ++d) {
    if (drafts[d].solo) {
      // everything in a solo pass acts inside the big gate's qubits: one matrix, in application order
      Cluster c;
      c.gate = canon[drafts[d].ids[0]];
      for (unsigned p : c.gate.pos) c.mask |= uint64_t(1) << p;
      c.ids.push_back(drafts[d].ids[0]);
      for (size_t j = 1; j < drafts[d].ids.size(); ++j) {
        merge_into(c.gate, canon[drafts[d].ids[j]]);
        c.ids.push_back(drafts[d].ids[j]);
      }
      merged[d].push_back(std::move(c));
    } else
      merged[d] = merge_pass(canon, drafts[d].ids, merge_max_k, merge_pass_cost, dtype, mma_on, mma_min_k);
    if (!drafts[d].solo && merge_pass_cost < 0 && merge_max_k > 2) {
      // second look at every cluster that grew beyond k = 2 on the strength of the slack: keep it only if it
      // really is cheaper than the same gates merged no further than k = 2 (a chain of two k = 2 gates is not, a
      // triangle of three is)
      std::vector<Cluster> refined;
      for (Cluster& c : merged[d]) {
        if (c.gate.k > 2 && c.ids.size() > 1) {
          std::vector<Cluster> sub = merge_pass(canon, c.ids, 2, merge_pass_cost, dtype, mma_on, mma_min_k);
          long split = 0;
          for (const Cluster& sc : sub) split += measured_cost(dtype, mma_on, mma_min_k, int(sc.gate.k));
          if (split < measured_cost(dtype, mma_on, mma_min_k, int(c.gate.k))) {
            for (Cluster& sc : sub) refined.push_back(std::move(sc));
            continue;
          }
        }
        refined.push_back(std::move(c));
      }
      merged[d].swap(refined);
    }
  }
  // ---- sparse "scalar + rank one" gates (a depolarizing channel as a super-operator: u = v = the vectorised identity,
  // 4 of 16 entries): lambda * 1 + u v^T = lambda * (1 + (u / lambda) v^T), and a scalar commutes with everything, so
  // the product of all such lambdas goes into ONE dense matrix of the plan and each gate only touches the amplitudes
  // where u or v is non-zero (gate_dr1 sparse form).  Needs a dense matrix somewhere in the plan to carry the scalar.
  {
    Canon* carrier = nullptr;
    for (size_t d = 0; d < drafts.size() && !carrier; ++d)
      for (Cluster& c : merged[d])
        if (!c.gate.dr1) { carrier = &c.gate; break; }
    std::complex<double> total(1, 0);
    if (carrier)
      for (size_t d = 0; d < drafts.size(); ++d)
        for (Cluster& c : merged[d]) {
          Canon& g = c.gate;
          if (!g.dr1) continue;
          double umax = 0, vmax = 0;
          for (const auto& x : g.u) umax = std::max(umax, std::abs(x));
          for (const auto& x : g.v) vmax = std::max(vmax, std::abs(x));
          size_t nz = 0;
          for (size_t i = 0; i < g.u.size(); ++i) {
            if (std::abs(g.u[i]) <= 1e-14 * umax) g.u[i] = 0;
            if (std::abs(g.v[i]) <= 1e-14 * vmax) g.v[i] = 0;
            if (g.u[i] != std::complex<double>(0, 0) || g.v[i] != std::complex<double>(0, 0)) ++nz;
          }
          if (2 * nz > g.u.size() || std::abs(g.lambda) < 1e-3) continue;
          for (auto& x : g.u) x /= g.lambda;
          total *= g.lambda;
          g.lambda = 1;
          g.dr1_folded = true;
        }
    if (carrier && total != std::complex<double>(1, 0))
      for (auto& x : carrier->U) x *= total;
  }
  for (size_t d = 0; d < drafts.size(); ++d) {
    total_gates += merged[d].size();
    // single gates may use shorter runs than the fuser is allowed to create
    const int L = choose_run_bits(drafts[d].bits, T, drafts[d].ids.size() > 1 ? fuse_min_run : hard_min_run);
    if (L < 0) { plan.error = "internal: pass does not fit"; return 1; }
    PassInfo& pi = infos[d];
    make_tile(drafts[d].bits, T, L, n, pi.header);
    make_iter_tables(pi.header, V);
    const int Tbits = int(pi.header.tile_bits);
    layouts[d].resize(merged[d].size());
    for (size_t ci = 0; ci < merged[d].size(); ++ci) {
      const Canon& c = merged[d][ci].gate;
      GateLayout& gl = layouts[d][ci];
      for (unsigned i = 0; i < c.k; ++i) {
        const int lb = local_bit(pi.header, c.pos[i]);
        if (lb < 0) { plan.error = "internal: target bit outside tile"; return 1; }
        gl.tpos[i] = uint8_t(lb);
      }
      // a lone k <= 3 gate keeps its plain matrix: such passes go to the direct kernel (hq_abi.cu)
      const bool lone_small = merged[d].size() == 1 && c.k <= 3;
      gl.dr1 = c.dr1 && !lone_small;
      gl.mma = !gl.dr1 && mma_on && !lone_small && int(c.k) >= mma_min_k && int(c.k) <= HQ_MMA_MAX_K &&
               mma_layout(gl.tpos, int(c.k), Tbits, V, gl.L);
      gl.bytes = gl.dr1 ? (((esz * (1 + (size_t(2) << c.k) + 21)) + 15) & ~size_t(15))
                        : (gl.mma ? mma_frag_bytes(c.k) : (((esz << (2 * c.k)) + 15) & ~size_t(15)));
      mat_bytes += gl.bytes;
    }
  }
  plan.n_kernel_gates = unsigned(total_gates);
  size_t off = total_gates * sizeof(HqGateDesc);
  off = (off + 15) & ~size_t(15);
  const size_t mat_base = off;
  plan.program.assign(mat_base + mat_bytes + 16, 0);

  size_t gate_cursor = 0;
  size_t mat_cursor = mat_base;
  for (size_t di = 0; di < drafts.size(); ++di) {
    PassInfo& pi = infos[di];
    pi.header.n_gates = uint32_t(merged[di].size());
    pi.header.gates_off = uint32_t(gate_cursor * sizeof(HqGateDesc));
    const int Tbits = int(pi.header.tile_bits);
    const int Tu = Tbits - V;
    // ---- warp-closed chains (complex64 FFMA2 slots): consecutive slot gates whose unit-level target bits
    // leave at least three unit bits untouched by ALL of them get those three bits as the warp-index bits of
    // their work-item maps.  The units a warp works on are then the same set for every gate of the chain, so
    // __syncwarp() replaces the CTA barrier between them (HqPassHeader::chain_mask).
    const size_t ng = merged[di].size();
    std::vector<uint32_t> chain_w(ng, 0);          // mask of the three warp-index unit bits (0 = not chained)
    std::vector<size_t> chain_first(ng, 0);
#if HQ_WARP_CHAINS
    if (dtype == HQ_DTYPE_C64 && opts.fast_slots != 0 && Tu >= HQ_THREADS_LOG2 + 3) {
      auto slot_gate = [&](size_t ci, uint32_t& tmask) {
        const Canon& c = merged[di][ci].gate;
        const GateLayout& gl = layouts[di][ci];
        if (ci >= HQ_FAST_SLOTS || gl.mma || gl.dr1 || c.k > HQ_FAST_MAX_K || (ng == 1 && c.k <= 3)) return false;
        tmask = 0;
        int kk = 0;
        for (unsigned i = 0; i < c.k; ++i)
          if (int(gl.tpos[i]) >= V) { tmask |= 1u << (gl.tpos[i] - V); ++kk; }
        return Tu - kk >= HQ_THREADS_LOG2;          // at least one work item per thread
      };
      size_t a = 0;
      while (a < ng) {
        uint32_t un = 0;
        if (!slot_gate(a, un)) { ++a; continue; }
        size_t b = a + 1;
        for (; b < ng; ++b) {
          uint32_t tm = 0;
          if (!slot_gate(b, tm)) break;
          if (__builtin_popcount(un | tm) > Tu - 3) break;
          un |= tm;
        }
        if (b - a >= 2) {
          uint32_t w = 0;
          int cnt = 0;
          for (int u = Tu - 1; u >= 0 && cnt < 3; --u)
            if (!((un >> u) & 1u)) { w |= 1u << u; ++cnt; }
          for (size_t ci = a; ci < b; ++ci) { chain_w[ci] = w; chain_first[ci] = a; }
        }
        a = b;
      }
    }
#endif
    std::vector<std::vector<uint16_t>> warp_sets;      // closure self-check: sorted slots of warp 0 per chained gate
    warp_sets.resize(ng);
    for (size_t ci = 0; ci < merged[di].size(); ++ci) {
      const Cluster& cluster = merged[di][ci];
      const GateLayout& gl = layouts[di][ci];
      const Canon& c = cluster.gate;
      HqGateDesc gd;
      memset(&gd, 0, sizeof(gd));
      gd.k = c.k;
      gd.kind = gl.dr1 ? HQ_GATE_DR1 : (gl.mma ? HQ_GATE_MMA : (c.k <= HQ_SMALL_K ? HQ_GATE_SMALL : HQ_GATE_BIG));
      gd.mat_off = uint32_t(mat_cursor);
      std::vector<bool> is_t;
      is_t.assign(size_t(Tbits), false);
      for (unsigned i = 0; i < c.k; ++i) {
        gd.tpos[i] = gl.tpos[i];
        is_t[size_t(gl.tpos[i])] = true;
      }
      // canonical positions are ascending globally, hence ascending locally too
      std::vector<unsigned> free_bits;
      if (gd.kind == HQ_GATE_SMALL || gd.kind == HQ_GATE_DR1) {
        for (int u = 0; u < Tu; ++u)
          if (!is_t[size_t(u + V)] && !((chain_w[ci] >> u) & 1u)) free_bits.push_back(unsigned(u));
        free_bits = lane_order(free_bits);
        if (chain_w[ci]) {
          // work-item bits 0..4 = lanes, 5..7 = warp index (the chain's common bits), 8.. = iterations
          std::vector<unsigned> q(free_bits.begin(), free_bits.begin() + 5);
          for (int u = 0; u < Tu; ++u)
            if ((chain_w[ci] >> u) & 1u) q.push_back(unsigned(u));
          q.insert(q.end(), free_bits.begin() + 5, free_bits.end());
          free_bits.swap(q);
        }
      } else if (gd.kind == HQ_GATE_BIG) {
        for (int a = 0; a < Tbits; ++a)
          if (!is_t[size_t(a)]) free_bits.push_back(unsigned(a));
      }
      gd.n_free = uint32_t(free_bits.size());
      for (size_t i = 0; i < free_bits.size() && i < 16; ++i) gd.q[i] = uint8_t(free_bits[i]);
      if (gd.kind == HQ_GATE_SMALL || gd.kind == HQ_GATE_DR1) make_lane_tables(gd, Tu, V);
      if (chain_w[ci]) {
        // the units warp 0 touches in this gate (all its lanes, all iterations, all units of a work item)
        const bool low = V == 1 && gd.tpos[0] == 0;
        const int KK = int(gd.k) - (low ? 1 : 0);
        const int niter = 1 << (Tu - KK - HQ_THREADS_LOG2);
        std::vector<uint16_t>& ws = warp_sets[ci];
        for (int t = 0; t < 32; ++t)
          for (int it = 0; it < niter; ++it)
            for (int m = 0; m < (1 << KK); ++m) ws.push_back(uint16_t(gd.tbl_thread[t] ^ gd.tbl_iter[it] ^ gd.tbl_x[m]));
        std::sort(ws.begin(), ws.end());
        if (ci > chain_first[ci] && chain_w[ci - 1] == chain_w[ci]) {
          if (ws == warp_sets[ci - 1]) pi.header.chain_mask |= 1u << (ci - 1);      // gate ci-1 -> ci: warp sync only
        }
      }
      if (gd.kind == HQ_GATE_MMA) make_mma_tables(gd, gl.L, V);
      if (gd.kind == HQ_GATE_SMALL && dtype == HQ_DTYPE_C128 && opts.fast_slots != 0 && (gd.k == 2 || gd.k == 3)) {
        make_rowpair_tables(gd, Tu);
        gd.kind = HQ_GATE_ROWPAIR;
      }
      memcpy(plan.program.data() + gate_cursor * sizeof(HqGateDesc), &gd, sizeof(gd));
      if (gd.kind == HQ_GATE_DR1) {
        if (dtype == HQ_DTYPE_C64) write_dr1<float>(plan.program, mat_cursor, c, gd.tbl_x);
        else write_dr1<double>(plan.program, mat_cursor, c, gd.tbl_x);
      } else if (gd.kind == HQ_GATE_MMA)
        write_mma_fragments(plan.program, mat_cursor, c, gl.L, dtype);
      else if (dtype == HQ_DTYPE_C64)
        write_matrix<float>(plan.program, mat_cursor, c, gd.kind == HQ_GATE_BIG);
      else
        write_matrix<double>(plan.program, mat_cursor, c, gd.kind == HQ_GATE_BIG);
      mat_cursor += gl.bytes;
      ++gate_cursor;
      // slot s = the s-th eligible matrix of the pass (bit ci of fast_mask marks it; the kernel counts the bits below)
      const unsigned slots_used = unsigned(__builtin_popcount(pi.header.fast_mask));
      if (dtype == HQ_DTYPE_C64 && gd.kind == HQ_GATE_SMALL && opts.fast_slots != 0 && ci < 32 && slots_used < HQ_FAST_SLOTS &&
          c.k <= HQ_FAST_MAX_K && Tbits - int(c.k) >= 1) {
        pi.header.fast_mask |= 1u << ci;
        pi.header.fast_k[slots_used] = uint8_t(c.k | (gd.tpos[0] == 0 ? 4u : 0u));
        for (size_t e = 0; e < (size_t(1) << (2 * c.k)); ++e) {
          pi.header.fast_u[slots_used][2 * e] = float(c.U[e].real());
          pi.header.fast_u[slots_used][2 * e + 1] = float(c.U[e].imag());
        }
      }
      for (unsigned id : cluster.ids) pi.gate_ids.push_back(canon_id[id]);
      // kernel class: a scalar + rank-one gate needs the registers of a k = 3 pass, not of a k = 4 one
      pi.header.max_k = std::max<uint32_t>(pi.header.max_k, gd.kind == HQ_GATE_DR1 ? 3u : c.k);
    }
    plan.passes.push_back(std::move(pi));
  }
  // ---- tcgen05 operand blocks for passes made of one dense complex64 k = 4 .. 6 matrix (hq_umma.cuh) ----
  if (dtype == HQ_DTYPE_C64) {
    for (size_t di = 0; di < drafts.size(); ++di) {
      if (merged[di].size() != 1 || !layouts[di][0].mma || layouts[di][0].dr1) continue;
      const Canon& c = merged[di][0].gate;
      if (c.k < UMMA_MIN_K || c.k > UMMA_MAX_K || n < c.k + UMMA_ROW_BITS || plan.passes[di].header.has_perm) continue;
      const size_t R = size_t(2) << c.k, floats = R * R;
      const size_t at = (plan.program.size() + 15) & ~size_t(15);
      if (at + 2 * floats * 4 + 16 > 0xffffffffull) break;
      plan.program.resize(at + 2 * floats * 4 + 16, 0);
      float* hi = reinterpret_cast<float*>(plan.program.data() + at);
      umma_pack_matrix(c.U.data(), c.k, hi, hi + floats);
      plan.passes[di].umma_off = uint32_t(at);
    }
  }
  return 0;
}

namespace {
float tf32_round_nearest(float x) {      // cvt.rna.tf32.f32: nearest, ties away from zero
  uint32_t b;
  memcpy(&b, &x, 4);
  b = (b + 0x1000u) & 0xffffe000u;
  memcpy(&x, &b, 4);
  return x;
}
}  // namespace

void umma_pack_matrix(const std::complex<double>* U, unsigned k, float* hi, float* lo) {
  const size_t dim = size_t(1) << k, R = 2 * dim;
  for (size_t nn = 0; nn < R; ++nn)
    for (size_t kk = 0; kk < R; ++kk) {
      const size_t i = nn >> 1, ri = nn & 1, j = kk >> 1, rj = kk & 1;
      const float re = float(U[i * dim + j].real()), im = float(U[i * dim + j].imag());
      const float v = ri == rj ? re : (ri == 0 ? -im : im);      // (re, im) of the output from (re, im) of the input
      const float h = tf32_round_nearest(v);
      const size_t at = ((kk >> 2) * R + nn) * 4 + (kk & 3);
      hi[at] = h;
      lo[at] = tf32_round_nearest(v - h);
    }
}

void make_identity_pass(int dtype, unsigned n, HqPassHeader& ph) {
  const int V = vbits(dtype);
  int T = std::min(default_tile_bits(dtype), max_tile_bits(dtype));
  T = std::min<int>(T, int(n));
  make_tile({}, T, T, n, ph);      // one contiguous run of 2^T amplitudes, no high bits
  make_iter_tables(ph, V);
}

int plan_build_bitperm(Plan& plan, int dtype, unsigned n, const std::vector<unsigned>& perm_full,
                       const PlanOptions& opts) {
  plan.dtype = dtype;
  plan.n_qubits = n;
  plan.passes.clear();
  plan.program.assign(16, 0);
  plan.error.clear();
  plan.n_gates = 0;
  const int V = vbits(dtype);
  if (perm_full.size() != n) { plan.error = "permutation must list every bit"; return 1; }
  {
    std::vector<bool> seen(n, false);
    for (unsigned p : perm_full) {
      if (p >= n || seen[p]) { plan.error = "not a permutation"; return 1; }
      seen[p] = true;
    }
  }
  int T = opts.tile_bits > 0 ? opts.tile_bits : default_tile_bits(dtype);
  T = std::min(T, max_tile_bits(dtype));
  T = std::min<int>(T, int(n));
  const int hard_min_run = V;
  int min_run = opts.min_run_bits >= 0 ? opts.min_run_bits : default_min_run_bits(dtype);
  min_run = std::max(hard_min_run, std::min(min_run, T));

  // decompose into transpositions of positions: cur[b] = old bit currently at position b
  std::vector<unsigned> cur(n);
  for (unsigned b = 0; b < n; ++b) cur[b] = b;
  std::vector<std::pair<unsigned, unsigned>> swaps;
  for (unsigned i = 0; i < n; ++i) {
    if (cur[i] == perm_full[i]) continue;
    unsigned j = i;
    while (cur[j] != perm_full[i]) ++j;
    std::swap(cur[i], cur[j]);
    swaps.push_back({i, j});
  }
  // group consecutive transpositions whose bits fit one tile
  size_t s = 0;
  while (s < swaps.size()) {
    std::vector<unsigned> bits;
    size_t e = s;
    while (e < swaps.size()) {
      std::vector<unsigned> u = bits;
      for (unsigned b : {swaps[e].first, swaps[e].second})
        if (std::find(u.begin(), u.end(), b) == u.end()) u.push_back(b);
      std::sort(u.begin(), u.end());
      const int mr = (e == s) ? hard_min_run : min_run;
      if (choose_run_bits(u, T, mr) < 0) break;
      bits.swap(u);
      ++e;
    }
    if (e == s) { plan.error = "transposition does not fit in a tile"; return 1; }
    PassInfo pi;
    const int L = choose_run_bits(bits, T, (e - s) > 1 ? min_run : hard_min_run);
    make_tile(bits, T, L, n, pi.header);
    make_iter_tables(pi.header, V);
    const int Tbits = int(pi.header.tile_bits);
    // compose: total[i] = p1[p2[...pk[i]]] where stage r has new bit a <- old bit b and vice versa
    std::vector<unsigned> total;
    total.resize(size_t(Tbits));
    for (int i = 0; i < Tbits; ++i) total[size_t(i)] = unsigned(i);
    for (size_t r = e; r-- > s;) {
      const int a = local_bit(pi.header, swaps[r].first);
      const int b = local_bit(pi.header, swaps[r].second);
      for (int i = 0; i < Tbits; ++i) {
        if (total[size_t(i)] == unsigned(a)) total[size_t(i)] = unsigned(b);
        else if (total[size_t(i)] == unsigned(b)) total[size_t(i)] = unsigned(a);
      }
    }
    pi.header.has_perm = 1;
    for (int i = 0; i < Tbits; ++i) pi.header.perm[i] = uint8_t(total[size_t(i)]);
    pi.header.n_gates = 0;
    pi.header.gates_off = 0;
    plan.passes.push_back(std::move(pi));
    s = e;
  }
  return 0;
}

}  // namespace hq
