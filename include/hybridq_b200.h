/*
 * hybridq_b200.h -- C ABI of libhybridq_b200.so, the B200-native (sm_100a) state-vector
 * evolution core for HybridQ.
 *
 * Part 1 is the DROP-IN boundary: the exact eleven symbols the reference binds with
 * ctypes from 'hybridq.so' and 'hybridq_swap.so' (the same library file is installed
 * under both names).  Host pointers in, host pointers out, results
 * visible in host memory on return, return code 0 = OK / 1 = rejected, never throws.
 *
 * Part 2 is the device-resident extension used by hybridq_b200.simulate(): the state
 * stays in HBM across gates as ONE interleaved-complex array, circuits are planned once
 * (gate fusion into tile passes) and run as a sequence of kernel launches on a stream.
 *
 * All functions return 0 on success unless stated otherwise; hq_last_error() gives a
 * human-readable message for the calling thread's last failure.  No torch types, plain
 * pointers and sizes only.  `stream` arguments are cudaStream_t passed as void* (NULL =
 * default stream).  `dtype`: 0 = complex64 (float pairs), 1 = complex128 (double pairs).
 */
#ifndef HYBRIDQ_B200_H
#define HYBRIDQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ===================================================================================
 * Part 1 -- reference-compatible symbols
 * =================================================================================== */

/* Replaces get_log2_pack_size, /root/reference/include/python_U.cpp:129 (bound at
 * hybridq/utils/dot.py:65-66).  The reference's host logic permutes index bits away
 * whenever a gate touches a bit below this value (simulation.py:559) and treats 0 as
 * "library missing" (simulation.py:393-394).  The GPU kernels have no pack, so this
 * returns the smallest legal value, 1. */
unsigned int get_log2_pack_size(void);

/* Replace apply_U_float32/64, /root/reference/include/python_U.cpp:131-143 (bound at
 * dot.py:69-78; called from simulation.py:640-646 and dot.py:305-307).
 * In place on two host real planes re[2^n], im[2^n]; U = row-major 2^k x 2^k complex,
 * interleaved; pos[i] = index bit (LSB 0) of matrix bit i.  Returns 1 -- and touches
 * nothing -- if re or im is not 32-byte aligned (U.h:34-36) or any pos[i] is below
 * get_log2_pack_size() (U.h:48-54) or >= n_qubits / duplicated; n_pos = 0 is a no-op
 * (python_U.cpp:38-39).  Returns 2 on a CUDA failure (see hq_last_error). */
int apply_U_float32(float* psi_re, float* psi_im, const float* U, const unsigned int* pos,
                    unsigned int n_qubits, unsigned int n_pos);
int apply_U_float64(double* psi_re, double* psi_im, const double* U, const unsigned int* pos,
                    unsigned int n_qubits, unsigned int n_pos);

/* Replace to_complex64/128, /root/reference/include/python_U.cpp:145-153 (bound at
 * dot.py:81-89; called from simulation.py:669-674): out[2i] = re[i], out[2i+1] = im[i].
 * `size` is 32-bit exactly as in the reference (python_U.cpp:115-116). */
int to_complex64(float* psi_re, float* psi_im, float* psi, unsigned int size);
int to_complex128(double* psi_re, double* psi_im, double* psi, unsigned int size);

/* Replace swap_*, /root/reference/include/python_swap.cpp:70-98 (bound at
 * hybridq/utils/transpose.py:42-58; called from simulation.py:623-630, :658-663,
 * dot.py:291-317, transpose.py:148).  In place on a host array of 2^n_qubits elements:
 * inside every aligned block of 2^n_pos elements new[j] = old[sigma(j)],
 * sigma(j) = XOR_i bit_i(j) << pos[i]  (swap.h:28-33).  n_pos = 0 is a no-op. */
int swap_float32(float* array, const unsigned int* pos, unsigned int n_qubits, unsigned int n_pos);
int swap_float64(double* array, const unsigned int* pos, unsigned int n_qubits, unsigned int n_pos);
int swap_int32(int* array, const unsigned int* pos, unsigned int n_qubits, unsigned int n_pos);
int swap_int64(long* array, const unsigned int* pos, unsigned int n_qubits, unsigned int n_pos);
int swap_uint32(unsigned int* array, const unsigned int* pos, unsigned int n_qubits, unsigned int n_pos);
int swap_uint64(unsigned long* array, const unsigned int* pos, unsigned int n_qubits, unsigned int n_pos);

/* ===================================================================================
 * Part 2 -- device-resident extension (what `_simulate_evolution`'s hybridq branch,
 * simulation.py:464-678, becomes when the state lives in HBM)
 * =================================================================================== */

int hq_version(void);
const char* hq_last_error(void);
int hq_device_count(int* count);
int hq_set_device(int device);
int hq_device_props(int* sm_count, size_t* total_mem, int* cc_major, int* cc_minor);

/* raw memory (replaces hybridq.utils.aligned.empty for the state, aligned_array.py:69-143) */
int hq_malloc(void** dptr, size_t bytes);
int hq_free(void* dptr);
int hq_host_alloc(void** hptr, size_t bytes);      /* pinned */
int hq_host_free(void* hptr);
int hq_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream);
int hq_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream);
int hq_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);
int hq_stream_sync(void* stream);

/* one gate on a device-resident interleaved state; any distinct pos in [0, n) incl. bit 0 */
int hq_apply_U_dev(void* state, int dtype, unsigned int n_qubits, const void* U_host,
                   const unsigned int* pos, unsigned int n_pos, void* stream);
/* same, with the shared-memory-free kernel (k <= 3); for measurements */
int hq_apply_U_direct_dev(void* state, int dtype, unsigned int n_qubits, const void* U_host,
                          const unsigned int* pos, unsigned int n_pos, void* stream);

/* In-place permutation of the low n_pos index bits of a device-resident interleaved state,
 * same sigma convention as swap_* (replaces the pair of swap calls on the re and im planes,
 * simulation.py:623-630). */
int hq_swap_dev(void* state, int dtype, unsigned int n_qubits, const unsigned int* pos,
                unsigned int n_pos, void* stream);

/* split planes <-> interleaved, all pointers on the device */
int hq_pack_dev(const void* re, const void* im, void* out, int dtype, uint64_t n_amps, void* stream);
int hq_unpack_dev(const void* in, void* re, void* im, int dtype, uint64_t n_amps, void* stream);

/* state preparation (prepare_state, hybridq/circuit/simulation/utils.py:40-156): spec is a
 * string of n_qubits characters from "01+-", spec[0] = most significant index bit */
int hq_init_product_dev(void* state, int dtype, unsigned int n_qubits, const char* spec, void* stream);
/* seeded complex Gaussian, normalised to 1; amplitude i depends only on (seed, index_offset+i) */
int hq_init_random_dev(void* state, int dtype, unsigned int n_qubits, uint64_t seed,
                       uint64_t index_offset, double scale, void* stream);
int hq_norm2_dev(const void* state, int dtype, uint64_t n_amps, double* result_host, void* stream);
int hq_vdot_dev(const void* a, const void* b, int dtype, uint64_t n_amps, double* re_im_host, void* stream);
int hq_scale_dev(void* state, int dtype, uint64_t n_amps, double factor, void* stream);
/* measurement support -- device-native halves of the reference's FunctionalGates, which otherwise force the
 * state back to the host (hybridq/circuit/simulation/simulation.py:525-554):
 * hq_marginal_dev: out_host[2 s] / out_host[2 s + 1] = sum of re^2 / im^2 over the amplitudes whose index bits
 *   pos[0..k) spell the outcome s (bit j of s = index bit pos[j]); 2 * 2^k doubles; their pairwise sums are
 *   the probabilities of hybridq/gate/measure.py:25-50 (_Measure, get_probs_only);
 * hq_project_dev: amplitudes with another outcome become 0, the kept ones are scaled by scale_re / scale_im
 *   plane-wise (hybridq/gate/projection.py:25-68 projects and renormalises the re and im planes separately). */
int hq_marginal_dev(const void* state, int dtype, unsigned int n_qubits, const unsigned int* pos,
                    unsigned int k, double* out_host, void* stream);
int hq_project_dev(void* state, int dtype, unsigned int n_qubits, const unsigned int* pos, unsigned int k,
                   unsigned int outcome, double scale_re, double scale_im, void* stream);
/* The same two operations without the k <= 10 limit of a shared-memory histogram (the reference's Measure /
 * Projection accept any number of qubits, gate/measure.py:77, gate/projection.py:72):
 * hq_marginal_cond_dev: as hq_marginal_dev for k <= 24 outcome bits, restricted to the amplitudes whose index
 *   satisfies (index & cond_mask) == cond_value -- lets a caller sample more than 24 qubits chunk by chunk
 *   (k = 0 gives the re^2 / im^2 sums of one outcome of any width: what Projection needs);
 * hq_project_mask_dev: keep (and scale plane-wise) the amplitudes with (index & mask) == value, zero the rest. */
int hq_marginal_cond_dev(const void* state, int dtype, unsigned int n_qubits, const unsigned int* pos,
                         unsigned int k, uint64_t cond_mask, uint64_t cond_value, double* out_host, void* stream);
int hq_project_mask_dev(void* state, int dtype, unsigned int n_qubits, uint64_t mask, uint64_t value,
                        double scale_re, double scale_im, void* stream);

/* ---- circuit plans: fuse a gate stream into tile passes once, run many times ---- */
typedef struct hq_plan hq_plan;

typedef struct hq_plan_options {
  int tile_bits;            /* log2 amplitudes per tile; 0 = default (13 c64 / 12 c128 = 64 KiB) */
  int min_run_bits;         /* smallest contiguous run the fuser may create; -1 = default */
  int fuse;                 /* 0 = one pass per gate */
  int max_gates_per_pass;   /* 0 = default */
  int lookahead;            /* 0 = default */
  int merge_max_k;          /* in-pass merging of gates into one matrix of at most this many qubits
                               (the reference's host-side `compress`, circuit/utils.py:467);
                               0 = off, -1 = default (4 with the tensor-core path, else 2) */
  int merge_pass_cost;      /* >= 0: analytic cost model cost(k) = 4*2^k + merge_pass_cost;
                               -1 = per-matrix costs measured on B200 (default) */
  int fast_slots;           /* complex64: pass the first 8 k=2 matrices of a pass as kernel parameters
                               (constant-bank FFMA operands); 0 = off, anything else = on (default) */
  int mma_min_k;            /* tensor-core path: gates with mma_min_k <= k <= 6 run on mma.sync (3xTF32 for
                               complex64, FP64 for complex128); 0 = never, -1 = default (3 complex64, 2 complex128) */
} hq_plan_options;

/* gates: n_gates entries; ks[g] = number of target bits; pos_flat = concatenated positions;
 * U_flat = concatenated row-major matrices, complex128 (double pairs) regardless of dtype */
hq_plan* hq_plan_create(int dtype, unsigned int n_qubits, unsigned int n_gates,
                        const unsigned int* ks, const unsigned int* pos_flat,
                        const double* U_flat, const hq_plan_options* opts);
/* plan made of permutation passes only: new index bit i <- old index bit perm[i], i < n_qubits */
hq_plan* hq_plan_create_bitperm(int dtype, unsigned int n_qubits, const unsigned int* perm,
                                const hq_plan_options* opts);
void hq_plan_destroy(hq_plan* plan);
int hq_plan_num_passes(const hq_plan* plan);
int hq_plan_num_gates(const hq_plan* plan);
/* number of matrices the kernels apply after in-pass merging (<= hq_plan_num_gates) */
int hq_plan_num_kernel_gates(const hq_plan* plan);
/* real floating-point operations one run of the plan performs: sum over kernel matrices of
 * 8 * 2^k * 2^n (2^k complex multiply-adds per output amplitude) */
double hq_plan_flops(const hq_plan* plan);
/* per pass: {tile_bits, n_high, n_kernel_gates, has_perm, n_gate_ids, high_pos[0..n_high)}
 * -> out[0..5+n_high); if out_len allows two more: fast_mask (bit s: kernel matrix s runs on a constant-bank FFMA2
 * slot) and chain_mask (bit s: matrices s and s+1 are separated by a warp-level sync only) */
int hq_plan_pass_info(const hq_plan* plan, int pass, unsigned int* out, int out_len);
/* gate ids (indices into the creation arrays) of a pass, in execution order */
int hq_plan_pass_gates(const hq_plan* plan, int pass, unsigned int* out, int out_len);
/* launch every pass on `stream` (asynchronous); the program is uploaded on first use */
int hq_plan_run(hq_plan* plan, void* state, void* stream);
/* launch passes [first, last) only */
int hq_plan_run_range(hq_plan* plan, void* state, int first, int last, void* stream);
/* which arithmetic the kernel matrices of a plan run on: out[5 * (k - 1) + a], k = 1..8, a = 0 constant-bank FFMA2
 * slot, 1 tensor cores (mma.sync 3xTF32 / FP64), 2 generic FMA path (incl. the direct kernel), 3 two-phase path,
 * 4 scalar + rank-one form (a depolarizing channel's super-operator: lambda * 1 + u v^T); out has 40 entries */
int hq_plan_arith_counts(const hq_plan* plan, unsigned int* out, int out_len);
/* how many of the scalar + rank-one gates run in the sparse form: lambda * 1 + u v^T = lambda * (1 + (u / lambda) v^T),
 * the scalars of all such gates multiplied into one dense matrix of the plan, and each gate touching only the
 * amplitudes where u or v is non-zero (a 2-qubit depolarizing channel as a 4-qubit super-operator: 4 of 16,
 * hybridq/noise/channel/channel.py:413-529) */
int hq_plan_sparse_rank_one_gates(const hq_plan* plan);

/* ---- multi-GPU: rank-bit <-> local-bit exchange fused into a pass (no reference counterpart: the reference's
 * evolution path refuses MPI, hybridq/circuit/simulation/simulation.py:379-380) ----
 * hq_plan_run_range_xchg runs passes [first, last) like hq_plan_run_range, but the LAST pass writes its result
 * to other buffers instead of back in place: the amplitudes whose LOCAL index bits pos[0..s) (s <= 3, amplitude-bit
 * positions) spell the digit D go to dst[D] -- this GPU's second shard buffer for D == mine, a peer GPU's buffer
 * opened with hq_ipc_open otherwise (stores travel over NVLink) -- at the same local index with those bits
 * replaced by the digit `mine`.  After every rank has run it (and a barrier), rank r's second buffer holds the
 * shard with rank bits and local bits pos[] swapped.  s = 0 is hq_plan_run_range.
 * hq_ipc_*: cudaIpc plumbing for the peer buffers (handle = 64 bytes, exchanged by the host code). */
int hq_plan_run_range_xchg(hq_plan* plan, void* state, int first, int last, unsigned int s, unsigned int mine,
                           const unsigned int* pos, void* const* dst, void* stream);
/* End-to-end run over PINNED host arrays (same interleaved layout as the device state): the first pass reads its
 * tiles straight from host_src over PCIe and the last pass writes its result straight to host_dst, so the upload and
 * the download overlap the arithmetic of those passes instead of being separate copies (either may be NULL; `state`
 * is the device working buffer and holds the result too unless host_dst is given).  Asynchronous on the stream. */
int hq_plan_run_io(hq_plan* plan, void* state, const void* host_src, void* host_dst, void* stream);
int hq_host_is_pinned(const void* ptr);
int hq_ipc_get_handle(void* dptr, void* handle_out_64);
int hq_ipc_open(const void* handle_64, void** dptr);
int hq_ipc_close(void* dptr);

/* measurement knobs of the tile kernel: nbuf = 0 (auto, default: prefetch the next tile while the
 * current one is processed whenever the second buffer costs no resident CTA), 1 (single-buffered)
 * or 2 (always double-buffered); ctas_per_sm = cap on resident CTAs per SM (0 = occupancy limit);
 * use_direct = 1 (default): a pass holding one k <= 3 gate runs on the shared-memory-free kernel.
 * Negative values leave a knob unchanged. */
int hq_set_tuning(int nbuf, int ctas_per_sm, int use_direct);
/* which kernel runs a pass: -1 (default) / 0 = the tile kernel (three independent CTAs per SM), except for the
 * exchange-redirect passes of hq_plan_run_range_xchg, which always run on the pipelined ring kernel (one
 * persistent CTA per SM, a 3-stage shared-memory ring fed by cp.async + mbarrier, two consumer groups);
 * 1 = the ring kernel for every pass whose tile has >= 256 units (measurements, tests). */
int hq_set_ring(int mode);

/* Blackwell tensor-core path for lone dense gates (hq_umma.cuh): a complex64 pass that consists of ONE dense k = 4, 5
 * or 6 matrix (after in-pass merging) on a state of at least k + 7 qubits runs on `tcgen05.mma kind::tf32` (3xTF32,
 * operands split hi / lo in registers and staged in shared memory, accumulators in TMEM) instead of the mma.sync
 * tile-kernel path; same contraction as /root/reference/include/U.h:123-202 for those k.
 * hq_set_umma: 1 = on (default), 0 = off, negative = query only; returns the previous setting.
 * hq_umma_launch_count: launches of that kernel by this process.  hq_plan_umma_passes: passes of a plan that qualify. */
int hq_set_umma(int mode);
uint64_t hq_umma_launch_count(void);
/* launches of the shared-memory-free direct kernel (a pass made of one gate with k <= 3, either precision) from plans */
uint64_t hq_direct_launch_count(void);
int hq_plan_umma_passes(const hq_plan* plan);

/* counters: kernels launched by this library in this process since the last reset */
uint64_t hq_launch_count(void);
void hq_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* HYBRIDQ_B200_H */
