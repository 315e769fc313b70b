// umma_lane_map_check.cu -- exhaustive HOST-side check of the tcgen05 gate kernel's lane map (csrc/hq_umma.cuh:
// umma_lane_map, umma_lane_split, umma_unit_slot).  No GPU needed: only the __host__ __device__ helpers run.
//
// For every target set of k = 4, 5, 6 bits among the 12 lowest amplitude bits (and a few high ones) it verifies that
//   1. the (warp, lane, iteration) -> (row, K-chunk) map is a bijection onto the tile, for the fill (K-chunks of the A
//      buffer) and for the staged epilogue (all K-chunks);
//   2. the 8 lanes of every quarter-warp hit 8 different 16-byte bank groups of shared memory with the skewed chunk
//      stride (slot = chunk * lbo + row);
//   3. the lanes of a quarter-warp never touch more 128-byte lines of global memory than with the row-only map, and
//      exactly one run of consecutive units when the low bits allow it.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -o tools/bin/umma_lane_map_check \
//        tools/umma_lane_map_check.cu && tools/bin/umma_lane_map_check        (tests/test_host_logic.py does this)
#include <cstdio>
#include <set>
#include <vector>

#include "../hybridq_b200/csrc/hq_umma.cuh"

static unsigned long long spread(unsigned long long g, const hq::UmmaPos& p, int k) {
  for (int b = 0; b < k; ++b) {
    const unsigned long long low = (1ull << p.tpos[b]) - 1ull;
    g = ((g & ~low) << 1) | (g & low);
  }
  return g;
}

int main() {
  long sets = 0, mixed = 0, failures = 0;
  long lines_rows_total = 0, lines_mem_total = 0, quarters = 0;
  for (int k = 4; k <= 6; ++k) {
    const int dim = 1 << k, CH = dim / 2, ksplit = k == 6 ? 2 : 1, CHH = CH / ksplit;
    int chh_bits = 0;
    while ((1 << chh_bits) < CHH) ++chh_bits;
    std::vector<std::vector<int>> tsets;
    for (unsigned m = 0; m < (1u << 12); ++m)
      if (__builtin_popcount(m) == k) {
        std::vector<int> t;
        for (int b = 0; b < 12; ++b)
          if ((m >> b) & 1u) t.push_back(b);
        tsets.push_back(t);
      }
    for (int base : {13, 20}) {            // high targets, and one low target with the rest high
      std::vector<int> t, u = {0};
      for (int i = 0; i < k; ++i) t.push_back(base + i);
      for (int i = 1; i < k; ++i) u.push_back(base + i);
      tsets.push_back(t);
      tsets.push_back(u);
    }
    for (const auto& t : tsets) {
      ++sets;
      hq::UmmaPos p;
      for (int i = 0; i < 8; ++i) p.tpos[i] = i < k ? (unsigned char)t[size_t(i)] : 0;
      hq::umma_lane_map(p, k, chh_bits);
      const int A = p.nchunk, lbo = p.lbo;
      if (A) ++mixed;
      if (__builtin_popcount(p.cmask) != A || A > chh_bits) { ++failures; printf("bad chunk count\n"); continue; }
      std::vector<unsigned long long> dep(size_t(dim), 0), rowoff(128);
      for (int j = 0; j < dim; ++j)
        for (int b = 0; b < k; ++b) dep[size_t(j)] |= (unsigned long long)((j >> b) & 1) << p.tpos[b];
      for (int r = 0; r < 128; ++r) rowoff[size_t(r)] = spread((unsigned long long)r, p, k);
      for (int nch : {CHH, CH}) {
        std::vector<int> seen(size_t(nch) * 128, 0);
        for (int warp = 0; warp < 4; ++warp)
          for (int i = 0; i < nch; ++i)
            for (int quarter = 0; quarter < 4; ++quarter) {
              std::set<int> banks;
              std::set<unsigned long long> lines_mem, lines_rows;
              for (int l8 = 0; l8 < 8; ++l8) {
                const int lane = quarter * 8 + l8;
                int c_lane, r_lane;
                hq::umma_lane_split(p.cmask, lane, c_lane, r_lane);
                const int slot = hq::umma_unit_slot(warp, i, nch, A, c_lane, r_lane);
                const int c = slot >> 7, r = slot & 127;
                if (c < 0 || c >= nch) { ++failures; continue; }
                ++seen[size_t(slot)];
                banks.insert((c * lbo + r) & 7);
                lines_mem.insert((rowoff[size_t(r)] | dep[size_t(2 * c)]) >> 4);
                // the row-only map: thread = row, iteration = chunk
                lines_rows.insert((rowoff[size_t(warp * 32 + lane)] | dep[size_t(2 * (i % nch))]) >> 4);
              }
              if (banks.size() != 8) { ++failures; if (failures < 10) printf("bank conflict k=%d cmask=%u lbo=%d\n", k, p.cmask, lbo); }
              if (lines_mem.size() > lines_rows.size()) { ++failures; if (failures < 10) printf("more lines than the row map k=%d\n", k); }
              ++quarters;
              lines_rows_total += long(lines_rows.size());
              lines_mem_total += long(lines_mem.size());
            }
        for (int v : seen)
          if (v != 1) { ++failures; if (failures < 10) printf("not a bijection k=%d cmask=%u A=%d nch=%d\n", k, p.cmask, A, nch); break; }
      }
    }
  }
  printf("{\"target_sets\": %ld, \"with_chunk_lanes\": %ld, \"failures\": %ld, \"lines_per_quarter_rows_map\": %.3f, "
         "\"lines_per_quarter_memory_order_map\": %.3f}\n",
         sets, mixed, failures, double(lines_rows_total) / double(quarters), double(lines_mem_total) / double(quarters));
  return failures ? 1 : 0;
}
