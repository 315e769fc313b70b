// hq_abi.cu -- the extern "C" boundary of libhybridq_b200.so (declared in
// include/hybridq_b200.h).  Part 1: the reference's eleven symbols with host-pointer
// semantics; Part 2: the device-resident extension.  Nothing here throws across the ABI.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/hybridq_b200.h"
#include "hq_kernels.h"
#include "hq_plan.h"

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
std::mutex g_scratch_mutex;

int fail(const std::string& msg, int rc = 2) {
  g_err = msg;
  return rc;
}
int cuda_fail(const char* what, int e) {
  g_err = std::string(what) + ": " + cudaGetErrorString(cudaError_t(e));
  return 2;
}
#define HQ_CUDA(call)                                      \
  do {                                                     \
    const int _e = int(call);                              \
    if (_e != 0) return cuda_fail(#call, _e);              \
  } while (0)

// grow-only device scratch for the host-pointer ABI, one per device
struct Scratch {
  void* ptr = nullptr;
  size_t cap = 0;
};
Scratch g_scratch[64];

int scratch_get(size_t bytes, void** out) {
  int dev = 0;
  HQ_CUDA(cudaGetDevice(&dev));
  Scratch& s = g_scratch[dev & 63];
  if (s.cap < bytes) {
    if (s.ptr) cudaFree(s.ptr);
    s.ptr = nullptr;
    s.cap = 0;
    HQ_CUDA(cudaMalloc(&s.ptr, bytes));
    s.cap = bytes;
  }
  *out = s.ptr;
  return 0;
}

bool positions_ok(const unsigned* pos, unsigned n, unsigned k, unsigned min_pos) {
  if (k > n) return false;
  for (unsigned i = 0; i < k; ++i) {
    if (pos[i] >= n || pos[i] < min_pos) return false;
    for (unsigned j = 0; j < i; ++j)
      if (pos[j] == pos[i]) return false;
  }
  return true;
}

template <typename T>
void fill_gate(hq::GateIn& g, const T* U, const unsigned* pos, unsigned k) {
  g.k = k;
  g.pos.assign(pos, pos + k);
  const size_t e = size_t(1) << (2 * k);
  g.U.resize(e);
  for (size_t i = 0; i < e; ++i) g.U[i] = std::complex<double>(double(U[2 * i]), double(U[2 * i + 1]));
}

int g_use_umma = 1;     // complex64 passes made of one dense k = 4 .. 6 matrix go to the tcgen05 kernel (hq_umma.cuh)
std::atomic<uint64_t> g_umma_launches{0};
std::atomic<uint64_t> g_direct_launches{0};
int g_use_direct = 1;   // single k <= 2 gate passes go to the shared-memory-free kernel

// xchg (may be null): exchange redirect applied to the write-back of the LAST pass of the range
int run_plan_passes(hq::Plan& plan, const unsigned char* d_prog, void* state, int first, int last, void* stream,
                    const hq::HqXchgDesc* xchg = nullptr, const void* io_src = nullptr, void* io_dst = nullptr) {
  int first_nonempty = -1, last_nonempty = -1;
  for (int p = first; p < last; ++p) {
    const HqPassHeader& ph = plan.passes[size_t(p)].header;
    if (ph.n_gates == 0 && !ph.has_perm) continue;
    if (first_nonempty < 0) first_nonempty = p;
    last_nonempty = p;
  }
  for (int p = first; p < last; ++p) {
    const HqPassHeader& ph = plan.passes[size_t(p)].header;
    const hq::HqXchgDesc* xg = (xchg && xchg->s && p == last - 1) ? xchg : nullptr;
    hq::HqXchgDesc io;
    if (io_src || io_dst) {
      // end-to-end run: the first pass reads the host array, the last pass writes the host array
      const bool is_first = p == first_nonempty, is_last = p == last_nonempty;
      if ((is_first && io_src) || (is_last && io_dst)) {
        memset(&io, 0, sizeof(io));
        io.src = is_first ? io_src : nullptr;
        io.dst[0] = (is_last && io_dst) ? io_dst : state;
        xg = &io;
      }
    }
    if (ph.n_gates == 0 && !ph.has_perm && !xg) continue;
    HqGateDesc gd;
    if (ph.n_gates == 1) memcpy(&gd, plan.program.data() + ph.gates_off, sizeof(gd));
    // (the gate's own k and kind decide, not the header's kernel class: a lone scalar + rank-one k = 4 gate carries
    // max_k = 3 -- found by tools/sanitize_target.py in round 2; SMALL and ROWPAIR -- the complex128 k = 2, 3 register
    // scheme -- both keep the plain row-major matrix the direct kernel reads)
    if (g_use_direct && !xg && ph.n_gates == 1 && !ph.has_perm && (gd.kind == HQ_GATE_SMALL || gd.kind == HQ_GATE_ROWPAIR) &&
        gd.k <= 3 &&
        plan.n_qubits >= gd.k + 1) {
      // measured (profiles/): the direct kernel runs a lone 1-/2-/3-qubit gate at copy bandwidth
      const unsigned L = ph.tile_bits - ph.n_high;
      unsigned pos[4];
      for (unsigned i = 0; i < gd.k; ++i) pos[i] = gd.tpos[i] < L ? gd.tpos[i] : ph.high_pos[gd.tpos[i] - L];
      const int rc = hq::launch_direct_gate(plan.dtype, state, plan.n_qubits, plan.program.data() + gd.mat_off, pos,
                                            gd.k, stream);
      if (rc) return cuda_fail("direct gate launch", rc);
      ++g_launches;
      ++g_direct_launches;
      continue;
    }
    if (g_use_umma && !xg && plan.passes[size_t(p)].umma_off && ph.n_gates == 1 && d_prog) {
      // measured (profiles/r02): 1.7-1.9x (k = 4), 2.4x (k = 5), 3x (k = 6) the bandwidth of the mma.sync tile-kernel path
      const unsigned L = ph.tile_bits - ph.n_high;
      unsigned pos[8];
      for (unsigned i = 0; i < gd.k; ++i) pos[i] = gd.tpos[i] < L ? gd.tpos[i] : ph.high_pos[gd.tpos[i] - L];
      const int rc = hq::launch_umma(state, plan.n_qubits, pos, gd.k, d_prog + plan.passes[size_t(p)].umma_off, stream);
      if (rc) return cuda_fail("tcgen05 gate launch", rc);
      ++g_launches;
      ++g_umma_launches;
      continue;
    }
    const int rc = hq::launch_pass(plan.dtype, state, plan.n_qubits, d_prog, ph, xg, stream, 0);
    if (rc) return cuda_fail("tile pass launch", rc);
    ++g_launches;
  }
  return 0;
}

// one-off plan: program uploaded with stream-ordered allocation and freed after the launches
int run_plan_once(hq::Plan& plan, void* state, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  void* d_prog = nullptr;
  HQ_CUDA(cudaMallocAsync(&d_prog, plan.program.size(), s));
  int rc = int(cudaMemcpyAsync(d_prog, plan.program.data(), plan.program.size(), cudaMemcpyHostToDevice, s));
  if (rc == 0) rc = run_plan_passes(plan, static_cast<const unsigned char*>(d_prog), state, 0, int(plan.passes.size()), stream);
  else rc = cuda_fail("program upload", rc);
  cudaFreeAsync(d_prog, s);
  return rc;
}

int apply_U_dev_impl(void* state, int dtype, unsigned n, const void* U_host, const unsigned* pos, unsigned k,
                     void* stream) {
  if (k == 0) return 0;
  if (!state || !U_host || !pos) return fail("null pointer", 1);
  if (!positions_ok(pos, n, k, 0)) return fail("invalid positions", 1);
  hq::GateIn g;
  if (dtype == HQ_DTYPE_C64) fill_gate(g, static_cast<const float*>(U_host), pos, k);
  else if (dtype == HQ_DTYPE_C128) fill_gate(g, static_cast<const double*>(U_host), pos, k);
  else return fail("bad dtype", 1);
  hq::Plan plan;
  hq::PlanOptions o;
  o.fuse = 0;
  if (hq::plan_build(plan, dtype, n, {g}, o)) return fail(plan.error, 1);
  return run_plan_once(plan, state, stream);
}

// host-pointer apply (Part 1)
template <typename T>
int apply_U_host(T* re, T* im, const T* U, const unsigned* pos, unsigned n, unsigned k) {
  const int dtype = sizeof(T) == 4 ? HQ_DTYPE_C64 : HQ_DTYPE_C128;
  if (k == 0) return 0;                                         // python_U.cpp:38-39
  if (!re || !im || !U || !pos) return fail("null pointer", 1);
  if (reinterpret_cast<uintptr_t>(re) % 32 || reinterpret_cast<uintptr_t>(im) % 32)
    return fail("psi_re/psi_im must be 32-byte aligned", 1);    // U.h:34-36
  if (!positions_ok(pos, n, k, get_log2_pack_size()))
    return fail("position below get_log2_pack_size(), out of range or duplicated", 1);   // U.h:48-54
  std::lock_guard<std::mutex> lock(g_scratch_mutex);
  const size_t amps = size_t(1) << n;
  const size_t plane = amps * sizeof(T);
  void* buf = nullptr;
  if (int rc = scratch_get(4 * plane, &buf)) return rc;
  char* d_re = static_cast<char*>(buf);
  char* d_im = d_re + plane;
  char* d_psi = d_re + 2 * plane;
  HQ_CUDA(cudaMemcpy(d_re, re, plane, cudaMemcpyHostToDevice));
  HQ_CUDA(cudaMemcpy(d_im, im, plane, cudaMemcpyHostToDevice));
  HQ_CUDA(hq::launch_pack(dtype, d_re, d_im, d_psi, amps, nullptr));
  ++g_launches;
  if (int rc = apply_U_dev_impl(d_psi, dtype, n, U, pos, k, nullptr)) return rc == 1 ? 1 : 2;
  HQ_CUDA(hq::launch_unpack(dtype, d_psi, d_re, d_im, amps, nullptr));
  ++g_launches;
  HQ_CUDA(cudaMemcpy(re, d_re, plane, cudaMemcpyDeviceToHost));
  HQ_CUDA(cudaMemcpy(im, d_im, plane, cudaMemcpyDeviceToHost));
  return 0;
}

template <typename T>
int to_complex_host(const T* re, const T* im, T* out, unsigned size) {
  const int dtype = sizeof(T) == 4 ? HQ_DTYPE_C64 : HQ_DTYPE_C128;
  if (size == 0) return 0;
  if (!re || !im || !out) return fail("null pointer", 1);
  std::lock_guard<std::mutex> lock(g_scratch_mutex);
  const size_t plane = size_t(size) * sizeof(T);
  void* buf = nullptr;
  if (int rc = scratch_get(4 * plane, &buf)) return rc;
  char* d_re = static_cast<char*>(buf);
  char* d_im = d_re + plane;
  char* d_out = d_re + 2 * plane;
  HQ_CUDA(cudaMemcpy(d_re, re, plane, cudaMemcpyHostToDevice));
  HQ_CUDA(cudaMemcpy(d_im, im, plane, cudaMemcpyHostToDevice));
  HQ_CUDA(hq::launch_pack(dtype, d_re, d_im, d_out, size, nullptr));
  ++g_launches;
  HQ_CUDA(cudaMemcpy(out, d_out, 2 * plane, cudaMemcpyDeviceToHost));
  return 0;
}

int swap_host(void* array, int elem_bytes, const unsigned* pos, unsigned n, unsigned m) {
  if (m == 0) return 0;                                         // python_swap.cpp:35-36
  if (!array || !pos) return fail("null pointer", 1);
  if (m > n || m > 32) return fail("n_pos out of range", 1);
  {
    unsigned long long seen = 0;
    for (unsigned i = 0; i < m; ++i) {
      if (pos[i] >= m || ((seen >> pos[i]) & 1ull)) return fail("pos is not a permutation of 0..n_pos-1", 1);
      seen |= 1ull << pos[i];
    }
  }
  std::lock_guard<std::mutex> lock(g_scratch_mutex);
  const size_t bytes = (size_t(1) << n) * size_t(elem_bytes);
  void* buf = nullptr;
  if (int rc = scratch_get(2 * bytes, &buf)) return rc;
  char* d_in = static_cast<char*>(buf);
  char* d_out = d_in + bytes;
  HQ_CUDA(cudaMemcpy(d_in, array, bytes, cudaMemcpyHostToDevice));
  HQ_CUDA(hq::launch_bitperm_oop(elem_bytes, d_in, d_out, n, pos, m, nullptr));
  ++g_launches;
  HQ_CUDA(cudaMemcpy(array, d_out, bytes, cudaMemcpyDeviceToHost));
  return 0;
}

hq::PlanOptions convert_opts(const hq_plan_options* o) {
  hq::PlanOptions p;
  if (o) {
    p.tile_bits = o->tile_bits;
    p.min_run_bits = o->min_run_bits;
    p.fuse = o->fuse;
    p.max_gates_per_pass = o->max_gates_per_pass;
    p.lookahead = o->lookahead;
    p.merge_max_k = o->merge_max_k;
    p.merge_pass_cost = o->merge_pass_cost;
    p.fast_slots = o->fast_slots;
    p.mma_min_k = o->mma_min_k;
  }
  return p;
}

}  // namespace

struct hq_plan {
  hq::Plan plan;
};

extern "C" {

// ------------------------------------------------------------------------------ Part 1
unsigned int get_log2_pack_size(void) { return 1u; }

int apply_U_float32(float* re, float* im, const float* U, const unsigned int* pos, unsigned int n, unsigned int k) {
  return apply_U_host<float>(re, im, U, pos, n, k);
}
int apply_U_float64(double* re, double* im, const double* U, const unsigned int* pos, unsigned int n, unsigned int k) {
  return apply_U_host<double>(re, im, U, pos, n, k);
}
int to_complex64(float* re, float* im, float* out, unsigned int size) { return to_complex_host<float>(re, im, out, size); }
int to_complex128(double* re, double* im, double* out, unsigned int size) { return to_complex_host<double>(re, im, out, size); }

int swap_float32(float* a, const unsigned int* pos, unsigned int n, unsigned int m) { return swap_host(a, 4, pos, n, m); }
int swap_float64(double* a, const unsigned int* pos, unsigned int n, unsigned int m) { return swap_host(a, 8, pos, n, m); }
int swap_int32(int* a, const unsigned int* pos, unsigned int n, unsigned int m) { return swap_host(a, 4, pos, n, m); }
int swap_int64(long* a, const unsigned int* pos, unsigned int n, unsigned int m) { return swap_host(a, 8, pos, n, m); }
int swap_uint32(unsigned int* a, const unsigned int* pos, unsigned int n, unsigned int m) { return swap_host(a, 4, pos, n, m); }
int swap_uint64(unsigned long* a, const unsigned int* pos, unsigned int n, unsigned int m) { return swap_host(a, 8, pos, n, m); }

// ------------------------------------------------------------------------------ Part 2
int hq_version(void) { return 100; }
const char* hq_last_error(void) { return g_err.c_str(); }

int hq_device_count(int* count) {
  HQ_CUDA(cudaGetDeviceCount(count));
  return 0;
}
int hq_set_device(int device) {
  HQ_CUDA(cudaSetDevice(device));
  return 0;
}
int hq_device_props(int* sm_count, size_t* total_mem, int* cc_major, int* cc_minor) {
  int dev = 0;
  HQ_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  HQ_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (total_mem) *total_mem = p.totalGlobalMem;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}

int hq_malloc(void** dptr, size_t bytes) {
  HQ_CUDA(cudaMalloc(dptr, bytes));
  return 0;
}
int hq_free(void* dptr) {
  HQ_CUDA(cudaFree(dptr));
  return 0;
}
int hq_host_alloc(void** hptr, size_t bytes) {
  HQ_CUDA(cudaMallocHost(hptr, bytes));
  return 0;
}
int hq_host_free(void* hptr) {
  HQ_CUDA(cudaFreeHost(hptr));
  return 0;
}
int hq_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
  HQ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}
int hq_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
  HQ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}
int hq_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
  HQ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}
int hq_stream_sync(void* stream) {
  HQ_CUDA(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int hq_apply_U_dev(void* state, int dtype, unsigned int n, const void* U_host, const unsigned int* pos,
                   unsigned int k, void* stream) {
  return apply_U_dev_impl(state, dtype, n, U_host, pos, k, stream);
}

int hq_apply_U_direct_dev(void* state, int dtype, unsigned int n, const void* U_host, const unsigned int* pos,
                          unsigned int k, void* stream) {
  if (k == 0) return 0;
  if (!state || !U_host || !pos) return fail("null pointer", 1);
  if (k > 3) return fail("direct kernel supports k <= 3", 1);
  if (!positions_ok(pos, n, k, 0)) return fail("invalid positions", 1);
  if (n < k + 1) return fail("state too small for the direct kernel", 1);
  // canonicalise through the planner's code path: build a 1-gate plan and read back the matrix
  hq::GateIn g;
  if (dtype == HQ_DTYPE_C64) fill_gate(g, static_cast<const float*>(U_host), pos, k);
  else if (dtype == HQ_DTYPE_C128) fill_gate(g, static_cast<const double*>(U_host), pos, k);
  else return fail("bad dtype", 1);
  hq::Plan plan;
  hq::PlanOptions o;
  o.fuse = 0;
  o.mma_min_k = 0;          // the direct kernel reads the plain row-major matrix
  if (hq::plan_build(plan, dtype, n, {g}, o)) return fail(plan.error, 1);
  HqGateDesc gd;
  memcpy(&gd, plan.program.data() + plan.passes[0].header.gates_off, sizeof(gd));
  unsigned sorted[4];
  for (unsigned i = 0; i < k; ++i) sorted[i] = pos[i];
  for (unsigned i = 1; i < k; ++i)
    for (unsigned j = i; j > 0 && sorted[j - 1] > sorted[j]; --j) { unsigned t = sorted[j]; sorted[j] = sorted[j - 1]; sorted[j - 1] = t; }
  const int rc = hq::launch_direct_gate(dtype, state, n, plan.program.data() + gd.mat_off, sorted, k, stream);
  if (rc) return cuda_fail("direct gate launch", rc);
  ++g_launches;
  return 0;
}

int hq_swap_dev(void* state, int dtype, unsigned int n, const unsigned int* pos, unsigned int m, void* stream) {
  if (m == 0) return 0;
  if (!state || !pos) return fail("null pointer", 1);
  if (m > n) return fail("n_pos out of range", 1);
  std::vector<unsigned> perm(n);
  for (unsigned b = 0; b < n; ++b) perm[b] = b < m ? pos[b] : b;
  hq::Plan plan;
  hq::PlanOptions o;
  if (hq::plan_build_bitperm(plan, dtype, n, perm, o)) return fail(plan.error, 1);
  return run_plan_once(plan, state, stream);
}

int hq_pack_dev(const void* re, const void* im, void* out, int dtype, uint64_t n_amps, void* stream) {
  HQ_CUDA(hq::launch_pack(dtype, re, im, out, n_amps, stream));
  ++g_launches;
  return 0;
}
int hq_unpack_dev(const void* in, void* re, void* im, int dtype, uint64_t n_amps, void* stream) {
  HQ_CUDA(hq::launch_unpack(dtype, in, re, im, n_amps, stream));
  ++g_launches;
  return 0;
}

int hq_init_product_dev(void* state, int dtype, unsigned int n, const char* spec, void* stream) {
  if (!state || !spec || strlen(spec) != n) return fail("spec must have n_qubits characters", 1);
  const int rc = hq::launch_init_product(dtype, state, n, spec, stream);
  if (rc) return cuda_fail("init_product (characters must be 0, 1, + or -)", rc);
  ++g_launches;
  return 0;
}

static int reduce_partials(int which, const void* a, const void* b, int dtype, uint64_t n_amps, double* out,
                           void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  hq::DeviceInfo di;
  HQ_CUDA(hq::device_info(&di));
  const unsigned np = unsigned(di.sm_count) * 8u;
  const size_t width = which == 0 ? 1 : 2;
  double* d_part = nullptr;
  HQ_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_part), np * width * sizeof(double), s));
  int rc = which == 0 ? hq::launch_norm2(dtype, a, n_amps, d_part, np, stream)
                      : hq::launch_vdot(dtype, a, b, n_amps, d_part, np, stream);
  std::vector<double> h(np * width);
  if (rc == 0) {
    ++g_launches;
    rc = int(cudaMemcpyAsync(h.data(), d_part, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
  }
  if (rc == 0) rc = int(cudaStreamSynchronize(s));
  cudaFreeAsync(d_part, s);
  if (rc) return cuda_fail("reduction", rc);
  for (size_t w = 0; w < width; ++w) {
    long double acc = 0;
    for (unsigned i = 0; i < np; ++i) acc += h[i * width + w];
    out[w] = double(acc);
  }
  return 0;
}

int hq_norm2_dev(const void* state, int dtype, uint64_t n_amps, double* result_host, void* stream) {
  if (!state || !result_host) return fail("null pointer", 1);
  return reduce_partials(0, state, nullptr, dtype, n_amps, result_host, stream);
}
int hq_vdot_dev(const void* a, const void* b, int dtype, uint64_t n_amps, double* re_im_host, void* stream) {
  if (!a || !b || !re_im_host) return fail("null pointer", 1);
  return reduce_partials(1, a, b, dtype, n_amps, re_im_host, stream);
}
int hq_marginal_cond_dev(const void* state, int dtype, unsigned int n, const unsigned int* pos, unsigned int k,
                         uint64_t cond_mask, uint64_t cond_value, double* out_host, void* stream) {
  if (!state || !out_host || (k && !pos)) return fail("null pointer", 1);
  if (k > HQ_MARGINAL_MAX_K || k > n) return fail("too many measured bits in one call (at most 24; sample in chunks)", 1);
  if (n < 64 && ((cond_mask >> n) || (cond_value & ~cond_mask))) return fail("bad condition mask / value", 1);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t bins = size_t(2) << k;
  double* d_out = nullptr;
  HQ_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_out), bins * sizeof(double), s));
  int rc = int(cudaMemsetAsync(d_out, 0, bins * sizeof(double), s));
  if (rc == 0) rc = hq::launch_marginal(dtype, state, n, pos, k, cond_mask, cond_value, d_out, stream);
  if (rc == 0) {
    ++g_launches;
    rc = int(cudaMemcpyAsync(out_host, d_out, bins * sizeof(double), cudaMemcpyDeviceToHost, s));
  }
  if (rc == 0) rc = int(cudaStreamSynchronize(s));
  cudaFreeAsync(d_out, s);
  if (rc) return cuda_fail("marginal", rc);
  return 0;
}
int hq_marginal_dev(const void* state, int dtype, unsigned int n, const unsigned int* pos, unsigned int k,
                    double* out_host, void* stream) {
  return hq_marginal_cond_dev(state, dtype, n, pos, k, 0, 0, out_host, stream);
}
int hq_project_mask_dev(void* state, int dtype, unsigned int n, uint64_t mask, uint64_t value, double scale_re,
                        double scale_im, void* stream) {
  if (!state) return fail("null pointer", 1);
  const int rc = hq::launch_project(dtype, state, n, mask, value, scale_re, scale_im, stream);
  if (rc) return cuda_fail("project", rc);
  ++g_launches;
  return 0;
}
int hq_project_dev(void* state, int dtype, unsigned int n, const unsigned int* pos, unsigned int k,
                   unsigned int outcome, double scale_re, double scale_im, void* stream) {
  if (!state || (k && !pos)) return fail("null pointer", 1);
  if (k > 32 || k > n || (k < 32 && (outcome >> k))) return fail("bad projection bits / outcome", 1);
  uint64_t mask = 0, value = 0;
  for (unsigned j = 0; j < k; ++j) {
    if (pos[j] >= n || ((mask >> pos[j]) & 1ull)) return fail("bad projection bits", 1);
    mask |= uint64_t(1) << pos[j];
    value |= uint64_t((outcome >> j) & 1u) << pos[j];
  }
  return hq_project_mask_dev(state, dtype, n, mask, value, scale_re, scale_im, stream);
}
int hq_scale_dev(void* state, int dtype, uint64_t n_amps, double factor, void* stream) {
  HQ_CUDA(hq::launch_scale(dtype, state, n_amps, factor, stream));
  ++g_launches;
  return 0;
}
int hq_init_random_dev(void* state, int dtype, unsigned int n, uint64_t seed, uint64_t index_offset, double scale,
                       void* stream) {
  if (!state) return fail("null pointer", 1);
  HQ_CUDA(hq::launch_init_random(dtype, state, n, seed, index_offset, stream));
  ++g_launches;
  if (scale == 0.0) {
    double n2 = 0;
    if (int rc = hq_norm2_dev(state, dtype, uint64_t(1) << n, &n2, stream)) return rc;
    scale = n2 > 0 ? 1.0 / std::sqrt(n2) : 1.0;
  }
  if (scale != 1.0) return hq_scale_dev(state, dtype, uint64_t(1) << n, scale, stream);
  return 0;
}

hq_plan* hq_plan_create(int dtype, unsigned int n, unsigned int n_gates, const unsigned int* ks,
                        const unsigned int* pos_flat, const double* U_flat, const hq_plan_options* opts) {
  hq_plan* p = new (std::nothrow) hq_plan();
  if (!p) { fail("out of memory"); return nullptr; }
  try {
    std::vector<hq::GateIn> gates(n_gates);
    size_t po = 0, uo = 0;
    for (unsigned g = 0; g < n_gates; ++g) {
      const unsigned k = ks[g];
      if (k > HQ_MAX_K) { fail("gate with k > 10"); delete p; return nullptr; }
      fill_gate(gates[g], U_flat + uo, pos_flat + po, k);
      po += k;
      uo += size_t(2) << (2 * k);
    }
    if (hq::plan_build(p->plan, dtype, n, gates, convert_opts(opts))) {
      fail(p->plan.error, 1);
      delete p;
      return nullptr;
    }
  } catch (const std::exception& e) {
    fail(e.what());
    delete p;
    return nullptr;
  }
  return p;
}

hq_plan* hq_plan_create_bitperm(int dtype, unsigned int n, const unsigned int* perm, const hq_plan_options* opts) {
  hq_plan* p = new (std::nothrow) hq_plan();
  if (!p) { fail("out of memory"); return nullptr; }
  std::vector<unsigned> pf(perm, perm + n);
  if (hq::plan_build_bitperm(p->plan, dtype, n, pf, convert_opts(opts))) {
    fail(p->plan.error, 1);
    delete p;
    return nullptr;
  }
  return p;
}

void hq_plan_destroy(hq_plan* plan) {
  if (!plan) return;
  if (plan->plan.d_program) cudaFree(plan->plan.d_program);
  delete plan;
}
int hq_plan_num_passes(const hq_plan* plan) { return plan ? int(plan->plan.passes.size()) : -1; }
int hq_plan_num_gates(const hq_plan* plan) { return plan ? int(plan->plan.n_gates) : -1; }
int hq_plan_num_kernel_gates(const hq_plan* plan) { return plan ? int(plan->plan.n_kernel_gates) : -1; }
double hq_plan_flops(const hq_plan* plan) {
  if (!plan) return -1.0;
  double f = 0;
  for (const hq::PassInfo& pi : plan->plan.passes)
    for (uint32_t g = 0; g < pi.header.n_gates; ++g) {
      HqGateDesc gd;
      memcpy(&gd, plan->plan.program.data() + pi.header.gates_off + size_t(g) * sizeof(HqGateDesc), sizeof(gd));
      f += 8.0 * double(1u << gd.k) * std::ldexp(1.0, int(plan->plan.n_qubits));
    }
  return f;
}

/* out[5 * (k - 1) + a] = number of kernel matrices of k = 1..8 qubits that run on arithmetic a:
 * 0 = constant-bank FFMA2 slot (complex64), 1 = tensor cores (mma.sync), 2 = generic FMA register / row-pair path,
 * 3 = two-phase path (k >= 5 without tensor cores), 4 = scalar + rank-one form (2 * 2^k MACs per group); lone
 * k <= 3 gates that take the direct kernel count as 2. */
int hq_plan_arith_counts(const hq_plan* plan, unsigned int* out, int out_len) {
  if (!plan || !out || out_len < 40) return fail("need 40 counters", 1);
  memset(out, 0, 40 * sizeof(unsigned));
  for (const hq::PassInfo& pi : plan->plan.passes)
    for (uint32_t g = 0; g < pi.header.n_gates; ++g) {
      HqGateDesc gd;
      memcpy(&gd, plan->plan.program.data() + pi.header.gates_off + size_t(g) * sizeof(HqGateDesc), sizeof(gd));
      if (gd.k < 1 || gd.k > 8) continue;
      const bool fast = plan->plan.dtype == HQ_DTYPE_C64 && pi.header.max_k <= 3 && g < 32u &&
                        ((pi.header.fast_mask >> g) & 1u) && pi.header.n_gates > 1;
      const unsigned a = fast ? 0u : (gd.kind == HQ_GATE_MMA ? 1u : (gd.kind == HQ_GATE_BIG ? 3u : (gd.kind == HQ_GATE_DR1 ? 4u : 2u)));
      ++out[5 * (gd.k - 1) + a];
    }
  return 0;
}

/* scalar + rank-one gates of the plan that run in the sparse form (scalar folded into another matrix, only the
 * amplitudes with a non-zero u / v component touched), e.g. depolarizing channels of a density-matrix circuit */
int hq_plan_sparse_rank_one_gates(const hq_plan* plan) {
  if (!plan) return -1;
  const size_t esz = plan->plan.dtype == HQ_DTYPE_C64 ? 4 : 8;
  int cnt = 0;
  for (const hq::PassInfo& pi : plan->plan.passes)
    for (uint32_t g = 0; g < pi.header.n_gates; ++g) {
      HqGateDesc gd;
      memcpy(&gd, plan->plan.program.data() + pi.header.gates_off + size_t(g) * sizeof(HqGateDesc), sizeof(gd));
      if (gd.kind != HQ_GATE_DR1) continue;
      const unsigned char* flag = plan->plan.program.data() + gd.mat_off + esz * (2 + (size_t(4) << gd.k) + 1);   // trailer[0].im
      double f = 0;
      if (esz == 4) { float x; memcpy(&x, flag, 4); f = x; } else memcpy(&f, flag, 8);
      cnt += f != 0 ? 1 : 0;
    }
  return cnt;
}

int hq_plan_pass_info(const hq_plan* plan, int pass, unsigned int* out, int out_len) {
  if (!plan || pass < 0 || pass >= int(plan->plan.passes.size())) return fail("bad pass index", 1);
  const HqPassHeader& ph = plan->plan.passes[size_t(pass)].header;
  if (out_len < 5 + int(ph.n_high)) return fail("output too small", 1);
  out[0] = ph.tile_bits;
  out[1] = ph.n_high;
  out[2] = ph.n_gates;
  out[3] = ph.has_perm;
  out[4] = unsigned(plan->plan.passes[size_t(pass)].gate_ids.size());
  for (unsigned i = 0; i < ph.n_high; ++i) out[5 + i] = ph.high_pos[i];
  if (out_len >= 7 + int(ph.n_high)) {        // optional tail: which gates sit on FFMA2 slots / are chained
    out[5 + ph.n_high] = ph.fast_mask;
    out[6 + ph.n_high] = ph.chain_mask;
  }
  return 0;
}
int hq_plan_pass_gates(const hq_plan* plan, int pass, unsigned int* out, int out_len) {
  if (!plan || pass < 0 || pass >= int(plan->plan.passes.size())) return fail("bad pass index", 1);
  const std::vector<unsigned>& ids = plan->plan.passes[size_t(pass)].gate_ids;
  if (out_len < int(ids.size())) return fail("output too small", 1);
  for (size_t i = 0; i < ids.size(); ++i) out[i] = ids[i];
  return 0;
}

int hq_plan_run_range(hq_plan* plan, void* state, int first, int last, void* stream) {
  if (!plan || !state) return fail("null pointer", 1);
  hq::Plan& pl = plan->plan;
  if (first < 0 || last > int(pl.passes.size()) || first > last) return fail("bad pass range", 1);
  int dev = 0;
  HQ_CUDA(cudaGetDevice(&dev));
  if (pl.d_program && pl.device != dev) {
    cudaFree(pl.d_program);
    pl.d_program = nullptr;
  }
  if (!pl.d_program) {
    HQ_CUDA(cudaMalloc(&pl.d_program, pl.program.size()));
    HQ_CUDA(cudaMemcpy(pl.d_program, pl.program.data(), pl.program.size(), cudaMemcpyHostToDevice));
    pl.device = dev;
  }
  return run_plan_passes(pl, static_cast<const unsigned char*>(pl.d_program), state, first, last, stream);
}
int hq_plan_run_range_xchg(hq_plan* plan, void* state, int first, int last, unsigned int s, unsigned int mine,
                           const unsigned int* pos, void* const* dst, void* stream) {
  if (!plan || !state) return fail("null pointer", 1);
  if (s == 0) return hq_plan_run_range(plan, state, first, last, stream);
  if (s > 3 || !pos || !dst) return fail("bad exchange descriptor", 1);
  hq::Plan& pl = plan->plan;
  if (first < 0 || last > int(pl.passes.size()) || first > last) return fail("bad pass range", 1);
  hq::HqXchgDesc xg;
  memset(&xg, 0, sizeof(xg));
  xg.s = s;
  xg.mine = mine;
  for (unsigned j = 0; j < s; ++j) xg.pos[j] = pos[j];
  for (unsigned d = 0; d < (1u << s); ++d) xg.dst[d] = dst[d];
  if (first == last) {
    // nothing to compute before the exchange: a gate-less pass carries the redirect
    HqPassHeader ph;
    hq::make_identity_pass(pl.dtype, pl.n_qubits, ph);
    const int rc = hq::launch_pass(pl.dtype, state, pl.n_qubits, nullptr, ph, &xg, stream, 0);
    if (rc) return cuda_fail("exchange pass launch", rc);
    ++g_launches;
    return 0;
  }
  int dev = 0;
  HQ_CUDA(cudaGetDevice(&dev));
  if (pl.d_program && pl.device != dev) {
    cudaFree(pl.d_program);
    pl.d_program = nullptr;
  }
  if (!pl.d_program) {
    HQ_CUDA(cudaMalloc(&pl.d_program, pl.program.size()));
    HQ_CUDA(cudaMemcpy(pl.d_program, pl.program.data(), pl.program.size(), cudaMemcpyHostToDevice));
    pl.device = dev;
  }
  return run_plan_passes(pl, static_cast<const unsigned char*>(pl.d_program), state, first, last, stream, &xg);
}

int hq_plan_run_io(hq_plan* plan, void* state, const void* host_src, void* host_dst, void* stream) {
  if (!plan || !state) return fail("null pointer", 1);
  hq::Plan& pl = plan->plan;
  bool any = false;
  for (const hq::PassInfo& pi : pl.passes) any = any || pi.header.n_gates || pi.header.has_perm;
  if (!any) return fail("plan without passes: copy the state instead", 1);
  for (const void* p : {host_src, (const void*)host_dst}) {
    if (!p) continue;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess || (at.type != cudaMemoryTypeHost && at.type != cudaMemoryTypeManaged)) {
      cudaGetLastError();
      return fail("hq_plan_run_io needs pinned (page-locked, mapped) host arrays", 1);
    }
  }
  int dev = 0;
  HQ_CUDA(cudaGetDevice(&dev));
  if (pl.d_program && pl.device != dev) {
    cudaFree(pl.d_program);
    pl.d_program = nullptr;
  }
  if (!pl.d_program) {
    HQ_CUDA(cudaMalloc(&pl.d_program, pl.program.size()));
    HQ_CUDA(cudaMemcpy(pl.d_program, pl.program.data(), pl.program.size(), cudaMemcpyHostToDevice));
    pl.device = dev;
  }
  return run_plan_passes(pl, static_cast<const unsigned char*>(pl.d_program), state, 0, int(pl.passes.size()), stream,
                         nullptr, host_src, host_dst);
}
int hq_host_is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (!p || cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return at.type == cudaMemoryTypeHost ? 1 : 0;
}

int hq_ipc_get_handle(void* dptr, void* handle_out_64) {
  if (!dptr || !handle_out_64) return fail("null pointer", 1);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  HQ_CUDA(cudaIpcGetMemHandle(&h, dptr));
  memcpy(handle_out_64, &h, sizeof(h));
  return 0;
}
int hq_ipc_open(const void* handle_64, void** dptr) {
  if (!handle_64 || !dptr) return fail("null pointer", 1);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_64, sizeof(h));
  HQ_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int hq_ipc_close(void* dptr) {
  HQ_CUDA(cudaIpcCloseMemHandle(dptr));
  return 0;
}

int hq_plan_run(hq_plan* plan, void* state, void* stream) {
  if (!plan) return fail("null pointer", 1);
  return hq_plan_run_range(plan, state, 0, int(plan->plan.passes.size()), stream);
}

int hq_set_tuning(int nbuf, int ctas_per_sm, int use_direct) {
  hq::set_tuning(nbuf, ctas_per_sm);
  if (use_direct >= 0) g_use_direct = use_direct;
  return 0;
}

int hq_set_ring(int mode) {
  hq::set_ring(mode);
  return 0;
}

int hq_set_umma(int mode) {
  const int old = g_use_umma;
  if (mode >= 0) g_use_umma = mode ? 1 : 0;
  return old;
}
uint64_t hq_umma_launch_count(void) { return g_umma_launches.load(); }
uint64_t hq_direct_launch_count(void) { return g_direct_launches.load(); }
int hq_plan_umma_passes(const hq_plan* plan) {
  if (!plan) return -1;
  int cnt = 0;
  for (const hq::PassInfo& pi : plan->plan.passes) cnt += pi.umma_off ? 1 : 0;
  return cnt;
}

uint64_t hq_launch_count(void) { return g_launches.load(); }
void hq_launch_count_reset(void) { g_launches.store(0); }

}  // extern "C"
