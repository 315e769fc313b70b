// hq_kernels.h -- host-callable launchers of the sm_100a kernels (hq_kernels.cu).
// Every launcher is asynchronous on `stream` and returns a cudaError_t as int.
#pragma once
#include <cstddef>
#include <cstdint>

#include "hq_common.h"

namespace hq {

struct DeviceInfo {
  int sm_count;
  int max_smem_optin;
};
int device_info(DeviceInfo* out);

// One pass of the tile kernel over a 2^n-amplitude interleaved-complex state.
//   prog       device pointer to the plan's program buffer
//   ph         pass header (by value)
//   launches   if non-null, incremented by the number of kernel launches issued
int launch_tile_pass(int dtype, void* state, unsigned n_qubits, const unsigned char* prog,
                     const HqPassHeader& ph, void* stream, int grid_override);
size_t tile_pass_smem_bytes(const HqPassHeader& ph, int dtype, int nbuf);
// nbuf: 0 = auto (double-buffer when it costs no resident CTA), 1 = single-buffered tiles,
// 2 = next tile prefetched while the current one is processed;
// ctas_per_sm: cap on resident CTAs per SM (0 = occupancy limit)
void set_tuning(int nbuf, int ctas_per_sm);

// Exchange redirect of a pass's write-back (multi-GPU, hybridq_b200/dist.py): the amplitudes whose LOCAL index bits
// pos[0..s) spell the digit D are written to buffer dst[D] -- this GPU's second buffer for D == mine, a peer GPU's
// (mapped over NVLink) otherwise -- at the same local index with those bits replaced by `mine`.
struct HqXchgDesc {
  unsigned s;
  unsigned mine;
  unsigned pos[4];
  void* dst[8];
  // optional: the pass reads its tiles from `src` instead of the state buffer (same layout; e.g. a pinned host
  // array for the first pass of an end-to-end run); with s = 0, dst[0] is where the whole result goes
  const void* src;
};
// One pass over the state: the pipelined ring kernel (hq_ring_kernel) for large states, hq_tile_kernel otherwise.
// xchg may be null.  ring mode: -1 auto, 0 never, 1 whenever the tile allows it (tests).
int launch_pass(int dtype, void* state, unsigned n_qubits, const unsigned char* prog, const HqPassHeader& ph,
                const HqXchgDesc* xchg, void* stream, int grid_override);
void set_ring(int mode);

// Direct (no shared memory) single-gate kernel for k <= 3 with every target at or above
// amplitude bit `V` (see hq_kernels.cu); U is read from kernel parameters.
// tcgen05 / TMEM kernel for ONE dense complex64 gate, k = 4, 5 or 6, n_qubits >= k + 7 (hq_umma.cuh).  pos ascending;
// d_operands = device pointer to the B_hi, B_lo blocks written by umma_pack_matrix (hq_plan.h).  Returns a cudaError.
int launch_umma(void* state, unsigned n_qubits, const unsigned* pos, unsigned k, const void* d_operands, void* stream);

int launch_direct_gate(int dtype, void* state, unsigned n_qubits, const void* U_host,
                       const unsigned* pos_sorted, unsigned k, void* stream);

// split planes <-> interleaved complex (reference to_complex, python_U.cpp:114-123)
int launch_pack(int dtype, const void* re, const void* im, void* out, uint64_t n_amps, void* stream);
int launch_unpack(int dtype, const void* in, void* re, void* im, uint64_t n_amps, void* stream);

// out[j] = in[(j & ~(2^m-1)) | sigma(j)], sigma(j) = XOR_i bit_i(j) << pos[i]  (swap.h:28-33)
int launch_bitperm_oop(int elem_bytes, const void* in, void* out, unsigned n_bits,
                       const unsigned* pos, unsigned m, void* stream);

// |psi> = product state given by `spec` (chars 0,1,+,-; spec[0] is the most significant bit),
// reference prepare_state (/root/reference/hybridq/circuit/simulation/utils.py:40-156).
int launch_init_product(int dtype, void* state, unsigned n_qubits, const char* spec, void* stream);
// counter-based complex Gaussian state (unnormalised), for sizes no host can stage.
int launch_init_random(int dtype, void* state, unsigned n_qubits, uint64_t seed,
                       uint64_t index_offset, void* stream);
int launch_scale(int dtype, void* state, uint64_t n_amps, double factor, void* stream);
// partial[0..n_partial) <- per-block sums of |psi|^2 (double); caller adds them up.
int launch_norm2(int dtype, const void* state, uint64_t n_amps, double* partial,
                 unsigned n_partial, void* stream);
// partial[2b], partial[2b+1] <- per-block re/im of <a|b> = sum conj(a) * b
int launch_vdot(int dtype, const void* a, const void* b, uint64_t n_amps, double* partial,
                unsigned n_partial, void* stream);

// out_dev[2 s], out_dev[2 s + 1] += sum of re^2, im^2 over the amplitudes whose index bits pos[0..k) spell the
// outcome s (bit j of s = index bit pos[j]) and whose index satisfies (i & cond_mask) == cond_value; out_dev holds
// 2 * 2^k doubles and must be zeroed; k <= HQ_MARGINAL_MAX_K (measure.py:25-50)
#define HQ_MARGINAL_MAX_K 24
int launch_marginal(int dtype, const void* state, unsigned n_qubits, const unsigned* pos, unsigned k,
                    uint64_t cond_mask, uint64_t cond_value, double* out_dev, void* stream);
// projection onto the amplitudes with (index & mask) == value, plane-wise scale factors (projection.py:25-68)
int launch_project(int dtype, void* state, unsigned n_qubits, uint64_t mask, uint64_t value,
                   double scale_re, double scale_im, void* stream);

}  // namespace hq
