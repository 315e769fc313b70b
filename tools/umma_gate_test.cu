// umma_gate_test.cu -- stand-alone bring-up of the tcgen05 gate kernel (round 2): one dense k = 5 gate applied to a
// complex64 state with `tcgen05.mma kind::tf32` (3xTF32: hi*hi + lo*hi + hi*lo), accumulators in TMEM, operands in
// the K-major no-swizzle canonical layout built from 16-byte units (validated by microbench_tcgen05.cu).
//
//   rows (M = 128)   = 128 consecutive groups (one per thread)
//   K = N = 2 * 2^k  = the reals of a group, (re, im) interleaved, amplitude j of the group at K index 2j, 2j + 1
//   A unit (r, c)    = amplitudes 2c, 2c + 1 of group r  -> shared slot c * 128 + r        (LBO = 128, SBO = 8)
//   B unit (n, c)    = Bs[n][4c .. 4c+3], Bs = real form of U (row = output real, column = input real)
//                                                        -> shared slot c * N + n          (LBO = N, SBO = 8)
//   D[r][n] in TMEM lane r, column n; thread r reads its row back with tcgen05.ld and writes the group.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/umma_gate_test tools/umma_gate_test.cu
#include <cuda_runtime.h>

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

#include "../hybridq_b200/csrc/hq_umma.cuh"

static float trunc_tf32(float x) {
  uint32_t b;
  memcpy(&b, &x, 4);
  b = (b + 0x1000u) & 0xffffe000u;      // round to nearest, ties away (cvt.rna.tf32.f32)
  memcpy(&x, &b, 4);
  return x;
}

template <int KQ>
static void run(int n, const std::vector<unsigned>& pos, int reps, int mode = -1, int ctas = 0) {
  const int DIM = 1 << KQ, R = 2 * DIM, CH = R / 4;
  std::mt19937_64 rng(7 + n + pos[0]);
  std::normal_distribution<double> nd;
  const size_t namp = size_t(1) << n;
  std::vector<std::complex<float>> psi(namp);
  double nrm = 0;
  if (n <= 22) {
    for (auto& a : psi) { a = std::complex<float>(float(nd(rng)), float(nd(rng))); nrm += std::norm(a); }
    for (auto& a : psi) a /= float(std::sqrt(nrm));
  } else {      // timing only: a cheap pattern
    uint32_t z = 12345u;
    const float sc = 1.0f / float(std::sqrt(double(namp)));
    for (auto& a : psi) {
      z = z * 1664525u + 1013904223u;
      a = std::complex<float>((float(z >> 8) / 8388608.0f - 1.0f) * sc, (float((z * 2654435761u) >> 8) / 8388608.0f - 1.0f) * sc);
    }
  }
  std::vector<std::complex<double>> U(size_t(DIM) * DIM);
  for (auto& u : U) u = std::complex<double>(nd(rng), nd(rng)) / std::sqrt(2.0 * DIM);
  // real form, canonical units, hi / lo
  std::vector<float> Bhi(size_t(CH) * R * 4), Blo(size_t(CH) * R * 4);
  for (int nn = 0; nn < R; ++nn)
    for (int kk = 0; kk < R; ++kk) {
      const int i = nn >> 1, ri = nn & 1, j = kk >> 1, rj = kk & 1;
      const std::complex<float> u = std::complex<float>(U[size_t(i) * DIM + j]);
      const float v = ri == rj ? u.real() : (ri == 0 ? -u.imag() : u.imag());
      const float hi = trunc_tf32(v);
      const size_t unit = size_t(kk >> 2) * R + nn;
      Bhi[unit * 4 + (kk & 3)] = hi;
      Blo[unit * 4 + (kk & 3)] = trunc_tf32(v - hi);
    }
  float2* d_state;
  float4 *d_bhi, *d_blo;
  CK(cudaMalloc(&d_state, namp * 8));
  CK(cudaMalloc(&d_bhi, Bhi.size() * 4));
  CK(cudaMalloc(&d_blo, Blo.size() * 4));
  CK(cudaMemcpy(d_state, psi.data(), namp * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_bhi, Bhi.data(), Bhi.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_blo, Blo.data(), Blo.size() * 4, cudaMemcpyHostToDevice));
  hq::UmmaPos p;
  for (int i = 0; i < KQ; ++i) p.tpos[i] = (unsigned char)pos[size_t(i)];
  int rc = hq::launch_umma_gate<KQ>(d_state, unsigned(n), p, d_bhi, d_blo, nullptr, mode, ctas);
  if (rc) { printf("{\"error\": \"launch rc=%d\"}\n", rc); exit(1); }
  CK(cudaDeviceSynchronize());
  double err = 0;
  if (n <= 22) {
    std::vector<std::complex<float>> out(namp);
    CK(cudaMemcpy(out.data(), d_state, namp * 8, cudaMemcpyDeviceToHost));
    // host reference in double
    const size_t ngroups = namp >> KQ;
    for (size_t g = 0; g < ngroups; ++g) {
      size_t base = g;
      for (int i = 0; i < KQ; ++i) {
        const size_t low = (size_t(1) << pos[size_t(i)]) - 1;
        base = ((base & ~low) << 1) | (base & low);
      }
      std::complex<double> in[128];
      size_t idx[128];
      for (int j = 0; j < DIM; ++j) {
        size_t a = base;
        for (int b = 0; b < KQ; ++b) a |= size_t((j >> b) & 1) << pos[size_t(b)];
        idx[j] = a;
        in[j] = std::complex<double>(psi[a]);
      }
      for (int i = 0; i < DIM; ++i) {
        std::complex<double> s = 0;
        for (int j = 0; j < DIM; ++j) s += std::complex<double>(std::complex<float>(U[size_t(i) * DIM + j])) * in[j];
        err = std::max(err, std::abs(s - std::complex<double>(out[idx[i]])));
      }
    }
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r) hq::launch_umma_gate<KQ>(d_state, unsigned(n), p, d_bhi, d_blo, nullptr, mode, ctas);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= float(reps);
  char ps[64] = "";
  for (int i = 0; i < KQ; ++i) snprintf(ps + strlen(ps), sizeof(ps) - strlen(ps), i ? ",%u" : "%u", pos[size_t(i)]);
  printf("{\"test\": \"umma_gate\", \"k\": %d, \"n\": %d, \"pos\": [%s], \"mode\": %d, \"ctas\": %d, \"grid\": %u, \"max_abs_err\": %.3e, \"ms\": %.4f, \"GBps\": %.1f}\n",
         KQ, n, ps, mode, ctas, hq::umma_last_grid(), err, ms, 2.0 * double(namp) * 8 / (ms * 1e-3) / 1e9);
  cudaFree(d_state); cudaFree(d_bhi); cudaFree(d_blo);
}

int main(int argc, char** argv) {
  // umma_gate_test time <k> <n> <mode> <ctas> <pos...> : one timing run (for ncu); otherwise the whole suite
  if (argc > 6 && !strcmp(argv[1], "time")) {
    const int k = atoi(argv[2]), n = atoi(argv[3]), mode = atoi(argv[4]), ctas = atoi(argv[5]);
    std::vector<unsigned> ps;
    for (int i = 6; i < argc; ++i) ps.push_back(unsigned(atoi(argv[i])));
    if (int(ps.size()) != k) return 2;
    if (k == 6) run<6>(n, ps, 5, mode, ctas);
    if (k == 5) run<5>(n, ps, 5, mode, ctas);
    if (k == 4) run<4>(n, ps, 5, mode, ctas);
    if (k == 3) run<3>(n, ps, 5, mode, ctas);
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "k6")) {
    for (int mode = 0; mode < 2; ++mode) {
      run<6>(14, {1, 4, 7, 9, 12, 13}, 1, mode);
      run<6>(16, {0, 3, 8, 11, 13, 15}, 1, mode);
      run<6>(17, {0, 1, 2, 3, 4, 5}, 1, mode);
      run<6>(20, {2, 5, 6, 11, 17, 19}, 2, mode);
    }
    fflush(stdout);
    run<6>(28, {3, 7, 12, 20, 25, 27}, 5, -1);
    run<6>(28, {0, 7, 12, 20, 25, 27}, 5, -1);
    run<6>(28, {0, 1, 2, 3, 4, 5}, 5, 0);
    run<6>(28, {0, 1, 2, 3, 4, 5}, 5, 1);
    run<5>(28, {3, 7, 12, 20, 25}, 5, -1);
    return 0;
  }
  if (argc > 1 && !strcmp(argv[1], "quick")) {
    run<5>(16, {1, 4, 7, 9, 12}, 1, 0);
    for (int ctas = 1; ctas <= 2; ++ctas) run<5>(28, {3, 7, 12, 20, 25}, 5, 0, ctas);
    for (int ctas = 1; ctas <= 4; ctas *= 2) run<4>(28, {3, 7, 12, 20}, 5, 0, ctas);
    for (int ctas = 1; ctas <= 8; ctas *= 2) run<3>(28, {3, 7, 12}, 5, 0, ctas);
    return 0;
  }
  const int big = argc > 1 ? atoi(argv[1]) : 28;
  // correctness: every kernel variant (mode x pair16) for k = 3, 4, 5
  for (int mode = 0; mode < 2; ++mode) {
    run<5>(16, {1, 4, 7, 9, 12}, 1, mode);
    run<5>(16, {0, 3, 8, 13, 15}, 1, mode);
    run<5>(17, {0, 1, 2, 3, 4}, 1, mode);
    run<5>(20, {2, 5, 6, 11, 19}, 2, mode);
    run<4>(15, {1, 2, 7, 14}, 1, mode);
    run<4>(16, {0, 1, 5, 9}, 1, mode);
    run<3>(14, {2, 3, 13}, 1, mode);
    run<3>(14, {0, 6, 11}, 1, mode);
  }
  fflush(stdout);
  // timing
  const std::vector<std::vector<unsigned>> sets5 = {{3, 7, 12, 20, 25}, {0, 1, 2, 3, 4}, {2, 3, 4, 5, 6}, {8, 9, 10, 11, 12}};
  for (const auto& ps : sets5)
    for (int mode = 0; mode < 2; ++mode) run<5>(big, ps, 5, mode);
  for (int ctas = 1; ctas <= 2; ++ctas) {
    run<5>(big, {3, 7, 12, 20, 25}, 5, 0, ctas);
    run<5>(big, {0, 1, 2, 3, 4}, 5, 0, ctas);
  }
  fflush(stdout);
  for (int mode = 0; mode < 2; ++mode) {
    run<4>(big, {3, 7, 12, 20}, 5, mode);
    run<4>(big, {0, 1, 2, 3}, 5, mode);
    run<3>(big, {3, 7, 12}, 5, mode);
    run<3>(big, {0, 1, 2}, 5, mode);
  }
  for (int ctas = 2; ctas <= 6; ctas += 2) run<4>(big, {3, 7, 12, 20}, 5, 0, ctas);
  for (int ctas = 2; ctas <= 8; ctas += 2) run<3>(big, {3, 7, 12}, 5, 0, ctas);
  return 0;
}
