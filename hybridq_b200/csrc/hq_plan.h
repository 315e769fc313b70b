// hq_plan.h -- host-side planner: turns a stream of gate-applies into passes of the tile
// kernel (which gates share a pass, which index bits form the tile, how work items map to
// lanes) and serialises them into the program buffer the kernel reads.  Pure C++ (no CUDA),
// so the CPU-only test-suite exercises it directly.
//
// This is the B200-native counterpart of the bookkeeping the reference does per gate in
// Python (/root/reference/hybridq/circuit/simulation/simulation.py:556-646: low-bit swap
// windows, _map/_inv_map, one ctypes call per gate): here nothing is ever permuted in
// memory for a gate, high target bits become tile bits instead.
#pragma once
#include <complex>
#include <cstdint>
#include <string>
#include <vector>

#include "hq_common.h"

namespace hq {

struct GateIn {
  unsigned k = 0;
  std::vector<unsigned> pos;                 // pos[i] = amplitude-index bit of matrix bit i
  std::vector<std::complex<double>> U;       // row-major 2^k x 2^k
};

struct PlanOptions {
  int tile_bits = 0;        // T in amplitudes; 0 = default for the dtype
  int min_run_bits = -1;    // smallest run (log2 amplitudes) the fuser may create; -1 = default
  int fuse = 1;             // 0: one gate per pass
  int max_gates_per_pass = 0;   // 0 = default (48)
  int lookahead = 0;        // how many gates past the first blocked one the fuser scans; 0 = default
  // In-pass gate merging (the reference does this on the host too: utils.compress +
  // to_matrix_gate, /root/reference/hybridq/circuit/utils.py:467, :419, default max 4 qubits).
  // Two gates of a pass are multiplied into one matrix when that does not raise the cost
  // cost(k) of a merged matrix: the per-matrix times measured on B200 (hq_plan.cpp, measured_cost), or,
  // with merge_pass_cost >= 0, the analytic model 4 * 2^k + merge_pass_cost.
  int merge_max_k = -1;     // largest merged gate; 0 = no merging; -1 = default (4 with the tensor-core path, else 2)
  int merge_pass_cost = -1; // -1 = measured cost table
  int fast_slots = 1;       // 0 = never use the constant-bank fast slots (measurements)
  // Tensor-core path (hq_mma.cuh): gates with mma_min_k <= k <= HQ_MMA_MAX_K are applied with
  // mma.sync (3xTF32 for complex64, FP64 for complex128).  0 = never, -1 = default for the dtype.
  int mma_min_k = -1;
};

struct PassInfo {
  HqPassHeader header;
  std::vector<unsigned> gate_ids;            // indices into the input gate list
  // tcgen05 path (hq_umma.cuh): a complex64 pass made of ONE dense k = 4, 5 or 6 matrix also carries that matrix as
  // TF32 hi / lo operand blocks (B_hi then B_lo, each 2^(2k) * 4 floats) at this program offset; 0 = none
  uint32_t umma_off = 0;
};

struct Plan {
  int dtype = HQ_DTYPE_C64;
  unsigned n_qubits = 0;
  unsigned n_gates = 0;                      // gate-applies covered (k = 0 gates are dropped)
  unsigned n_kernel_gates = 0;               // matrices the kernels apply after in-pass merging
  std::vector<PassInfo> passes;
  std::vector<unsigned char> program;        // host copy of the device program buffer
  std::string error;
  // device side (filled by hq_abi.cu)
  void* d_program = nullptr;
  int device = -1;
};

// Smallest / largest k of the tcgen05 lone-gate kernel and the state size it needs (128 groups per tile).
constexpr unsigned UMMA_MIN_K = 4, UMMA_MAX_K = 6, UMMA_ROW_BITS = 7;
// Real form of a 2^k x 2^k complex matrix as the K-major operand blocks of tcgen05.mma kind::tf32:
// 16-byte unit (n, c) at index c * R + n holds Bs[n][4c .. 4c + 3], R = 2 * 2^k, Bs[2i + ri][2j + rj] = the real
// 2 x 2 block of U[i][j]; `hi` = rounded to TF32 (nearest), `lo` = the rounded remainder.  2 * R * R floats are written.
void umma_pack_matrix(const std::complex<double>* U, unsigned k, float* hi, float* lo);

int default_tile_bits(int dtype);
int default_min_run_bits(int dtype);
int default_merge_max_k(int dtype);
int default_mma_min_k(int dtype);

// Build a plan.  Returns 0 on success; on failure returns non-zero and sets plan.error.
int plan_build(Plan& plan, int dtype, unsigned n_qubits, const std::vector<GateIn>& gates,
               const PlanOptions& opts);

// A single pass whose only job is an in-place permutation of index bits:
// new bit i <- old bit perm[i] for the bits listed (a closed set of at most T bits).
int plan_build_bitperm(Plan& plan, int dtype, unsigned n_qubits, const std::vector<unsigned>& perm_full,
                       const PlanOptions& opts);

// Header of a pass that applies no gate and no permutation (a plain copy of the state through the tile kernel):
// the carrier of an exchange redirect when no local pass precedes the exchange (hq_plan_run_range_xchg).
void make_identity_pass(int dtype, unsigned n_qubits, HqPassHeader& ph);

}  // namespace hq
