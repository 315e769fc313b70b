/*
 * hq_oracle.c -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * A plain-C (scalar, optional OpenMP) restatement of the arithmetic that the
 * reference's native core performs for the state-vector evolution hot path:
 *
 *   oracle_apply_U_f32/f64   <- hybridq::U::apply          /root/reference/include/U.h:28-102 (k<=4)
 *                                                           /root/reference/include/U.h:123-202 (k>=5)
 *                               index expansion             /root/reference/include/utils.h:78-105
 *   oracle_swap_b32/b64      <- hybridq::swap::swap_array  /root/reference/include/swap.h:47-95
 *   oracle_to_complex_f32/64 <- hybridq::python::to_complex /root/reference/include/python_U.cpp:114-123
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's shared object.  The product
 * (hybridq_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here
 * against (a) the reference's own C++ core compiled from /root/reference into
 * oracle/_ref/ (when present) and (b) the golden vectors under tests/golden/
 * that were produced by the unmodified reference (tests/golden/make_golden.py).
 *
 * Conventions (identical to the reference ABI):
 *   - the state is two real planes re[2^n], im[2^n] (split, not interleaved);
 *   - U is a row-major 2^k x 2^k complex matrix stored interleaved [re,im];
 *   - pos[i] is the bit (LSB = 0) of the flat amplitude index that carries bit i
 *     of the matrix row/column index;
 *   - everything is done in place.
 *
 * Unlike the reference there is no SIMD pack, hence no "pos >= log2_pack_size"
 * restriction: any distinct positions in [0, n) are accepted.  Accumulation
 * order over j is the reference's (j ascending, re and im updated together).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAX_K 16

/* Scatter the bits of `m` to the positions pos[0..k): bit i of m -> bit pos[i]. */
static inline uint64_t deposit(uint64_t m, const unsigned *pos, unsigned k) {
  uint64_t y = 0;
  for (unsigned i = 0; i < k; ++i) y |= ((m >> i) & 1u) << pos[i];
  return y;
}

/* Spread counter c over the index bits NOT in `sorted` (ascending target
 * positions): for each target position, open a zero gap there.  This is the
 * closed form of the reference's sequential gap insertion (utils.h:90-99). */
static inline uint64_t open_gaps(uint64_t c, const unsigned *sorted,
                                 unsigned k) {
  for (unsigned i = 0; i < k; ++i) {
    const uint64_t low = ((uint64_t)1 << sorted[i]) - 1;
    c = ((c & ~low) << 1) | (c & low);
  }
  return c;
}

static int check_pos(const unsigned *pos, unsigned n, unsigned k,
                     unsigned *sorted) {
  if (k > n || k > ORACLE_MAX_K) return 1;
  for (unsigned i = 0; i < k; ++i) {
    if (pos[i] >= n) return 1;
    for (unsigned j = 0; j < i; ++j)
      if (pos[j] == pos[i]) return 1;
    sorted[i] = pos[i];
  }
  for (unsigned i = 1; i < k; ++i) { /* insertion sort, ascending */
    unsigned v = sorted[i], j = i;
    while (j > 0 && sorted[j - 1] > v) {
      sorted[j] = sorted[j - 1];
      --j;
    }
    sorted[j] = v;
  }
  return 0;
}

#define DEFINE_APPLY(NAME, T)                                                  \
  int NAME(T *re, T *im, const T *U, const unsigned *pos, unsigned n,          \
           unsigned k) {                                                       \
    unsigned sorted[ORACLE_MAX_K];                                             \
    if (k == 0) return 0;                                                      \
    if (check_pos(pos, n, k, sorted)) return 1;                                \
    const size_t dim = (size_t)1 << k;                                         \
    const uint64_t groups = (uint64_t)1 << (n - k);                            \
    uint64_t *off = (uint64_t *)malloc(dim * sizeof(uint64_t));                \
    if (!off) return 2;                                                        \
    for (size_t m = 0; m < dim; ++m) off[m] = deposit(m, pos, k);              \
    int fail = 0;                                                              \
    _Pragma("omp parallel")                                                    \
    {                                                                          \
      T *xr = (T *)malloc(dim * sizeof(T));                                    \
      T *xi = (T *)malloc(dim * sizeof(T));                                    \
      if (!xr || !xi) {                                                        \
        _Pragma("omp atomic write") fail = 1;                                  \
      } else {                                                                 \
        _Pragma("omp for schedule(static)")                                    \
        for (uint64_t g = 0; g < groups; ++g) {                                \
          const uint64_t base = open_gaps(g, sorted, k);                       \
          for (size_t j = 0; j < dim; ++j) {                                   \
            xr[j] = re[base | off[j]];                                         \
            xi[j] = im[base | off[j]];                                         \
          }                                                                    \
          for (size_t i = 0; i < dim; ++i) {                                   \
            T ar = 0, ai = 0;                                                  \
            const T *row = U + 2 * i * dim;                                    \
            for (size_t j = 0; j < dim; ++j) {                                 \
              const T ur = row[2 * j], ui = row[2 * j + 1];                    \
              ar += ur * xr[j] - ui * xi[j];                                   \
              ai += ur * xi[j] + ui * xr[j];                                   \
            }                                                                  \
            re[base | off[i]] = ar;                                            \
            im[base | off[i]] = ai;                                            \
          }                                                                    \
        }                                                                      \
      }                                                                        \
      free(xr);                                                                \
      free(xi);                                                                \
    }                                                                          \
    free(off);                                                                 \
    return fail ? 2 : 0;                                                       \
  }

DEFINE_APPLY(oracle_apply_U_f32, float)
DEFINE_APPLY(oracle_apply_U_f64, double)

/* Bit permutation of the lowest m index bits, per aligned block of 2^m
 * elements: new[j] = old[sigma(j)], sigma(j) = XOR_i bit_i(j) << pos[i].
 * (swap.h:28-33 defines sigma; :83-92 the block loop.) */
#define DEFINE_SWAP(NAME, T)                                                   \
  int NAME(T *a, const unsigned *pos, unsigned n, unsigned m) {                \
    if (m == 0) return 0;                                                      \
    if (m > n || m > 30) return 1;                                             \
    const size_t bs = (size_t)1 << m;                                          \
    size_t *src = (size_t *)malloc(bs * sizeof(size_t));                       \
    if (!src) return 2;                                                        \
    for (size_t j = 0; j < bs; ++j) {                                          \
      size_t y = 0;                                                            \
      for (unsigned i = 0; i < m; ++i) y ^= ((j >> i) & 1u) << pos[i];         \
      src[j] = y;                                                              \
    }                                                                          \
    const uint64_t blocks = (uint64_t)1 << (n - m);                            \
    int fail = 0;                                                              \
    _Pragma("omp parallel")                                                    \
    {                                                                          \
      T *buf = (T *)malloc(bs * sizeof(T));                                    \
      if (!buf) {                                                              \
        _Pragma("omp atomic write") fail = 1;                                  \
      } else {                                                                 \
        _Pragma("omp for schedule(static)")                                    \
        for (uint64_t b = 0; b < blocks; ++b) {                                \
          T *blk = a + (b << m);                                               \
          for (size_t j = 0; j < bs; ++j) buf[j] = blk[src[j]];                \
          memcpy(blk, buf, bs * sizeof(T));                                    \
        }                                                                      \
      }                                                                        \
      free(buf);                                                               \
    }                                                                          \
    free(src);                                                                 \
    return fail ? 2 : 0;                                                       \
  }

DEFINE_SWAP(oracle_swap_b32, uint32_t)
DEFINE_SWAP(oracle_swap_b64, uint64_t)

#define DEFINE_TO_COMPLEX(NAME, T)                                             \
  int NAME(const T *re, const T *im, T *out, uint64_t size) {                  \
    _Pragma("omp parallel for schedule(static)")                               \
    for (uint64_t i = 0; i < size; ++i) {                                      \
      out[2 * i] = re[i];                                                      \
      out[2 * i + 1] = im[i];                                                  \
    }                                                                          \
    return 0;                                                                  \
  }

DEFINE_TO_COMPLEX(oracle_to_complex_f32, float)
DEFINE_TO_COMPLEX(oracle_to_complex_f64, double)
