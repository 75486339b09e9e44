/* sv_oracle.c -- CPU restatement of the reference's gate kernels  (TEST INFRASTRUCTURE ONLY).
 *
 * Follows apply_1q_statevector / apply_2q_statevector / expect_z_statevector of the reference
 * (src/tyxonq/libs/quantum_library/kernels/statevector.py:28-68): the same contraction
 * psi'[..a..] = sum_b G[a,b] psi[..b..] on tensor axis q == index bit n-1-q, but written as an
 * in-place loop over amplitude groups and threaded with OpenMP, so that states of 26-30 qubits
 * (beyond the reference's einsum letter limit) can be timed on the host.
 * Checked against oracle/sv_oracle.py (numpy einsum) in tests/test_oracle_c.py.
 * Used only by tests/ and by bench.py's cpu_baseline / --impl reference legs.
 */
#include <complex.h>
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex c128;
typedef float complex c64;

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Set the OpenMP thread count (torchrun exports OMP_NUM_THREADS=1); returns the count in effect. */
int orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

static inline uint64_t insert_zero(uint64_t g, int p) { return ((g >> p) << (p + 1)) | (g & ((1ull << p) - 1ull)); }

#define DEFINE_KERNELS(T, SUF)                                                                       \
  /* g: 2x2 row-major, bit = n-1-qubit */                                                           \
  void orc_apply_1q_##SUF(T *psi, int n, int bit, const double complex *g) {                        \
    const uint64_t half = 1ull << (n - 1), st = 1ull << bit;                                         \
    const T g00 = (T)g[0], g01 = (T)g[1], g10 = (T)g[2], g11 = (T)g[3];                             \
    _Pragma("omp parallel for schedule(static)")                                                    \
    for (uint64_t i = 0; i < half; ++i) {                                                            \
      const uint64_t i0 = insert_zero(i, bit), i1 = i0 | st;                                         \
      const T a = psi[i0], b = psi[i1];                                                              \
      psi[i0] = g00 * a + g01 * b;                                                                   \
      psi[i1] = g10 * a + g11 * b;                                                                   \
    }                                                                                                \
  }                                                                                                  \
  /* g: 4x4 row-major, index = 2*(bit_hi value) + (bit_lo value); bit_hi = n-1-q0, bit_lo = n-1-q1 */ \
  void orc_apply_2q_##SUF(T *psi, int n, int bit_hi, int bit_lo, const double complex *g) {         \
    const uint64_t quarter = 1ull << (n - 2);                                                        \
    const int p0 = bit_hi < bit_lo ? bit_hi : bit_lo, p1 = bit_hi < bit_lo ? bit_lo : bit_hi;        \
    const uint64_t sh = 1ull << bit_hi, sl = 1ull << bit_lo;                                         \
    T m[16];                                                                                         \
    for (int k = 0; k < 16; ++k) m[k] = (T)g[k];                                                     \
    _Pragma("omp parallel for schedule(static)")                                                    \
    for (uint64_t i = 0; i < quarter; ++i) {                                                         \
      const uint64_t b = insert_zero(insert_zero(i, p0), p1);                                        \
      const uint64_t idx[4] = {b, b | sl, b | sh, b | sh | sl};                                      \
      const T v0 = psi[idx[0]], v1 = psi[idx[1]], v2 = psi[idx[2]], v3 = psi[idx[3]];                \
      for (int r = 0; r < 4; ++r)                                                                    \
        psi[idx[r]] = m[4 * r] * v0 + m[4 * r + 1] * v1 + m[4 * r + 2] * v2 + m[4 * r + 3] * v3;     \
    }                                                                                                \
  }                                                                                                  \
  double orc_expect_z_##SUF(const T *psi, int n, int bit) {                                          \
    const uint64_t dim = 1ull << n;                                                                  \
    double acc = 0.0;                                                                                \
    _Pragma("omp parallel for schedule(static) reduction(+ : acc)")                                 \
    for (uint64_t i = 0; i < dim; ++i) {                                                             \
      const double re = creal(psi[i]), im = cimag(psi[i]);                                           \
      const double p = re * re + im * im;                                                            \
      acc += ((i >> bit) & 1) ? -p : p;                                                              \
    }                                                                                                \
    return acc;                                                                                      \
  }                                                                                                  \
  void orc_init_zero_##SUF(T *psi, int n) {                                                          \
    const uint64_t dim = 1ull << n;                                                                  \
    _Pragma("omp parallel for schedule(static)")                                                    \
    for (uint64_t i = 0; i < dim; ++i) psi[i] = 0;                                                   \
    psi[0] = 1;                                                                                      \
  }

DEFINE_KERNELS(c128, c128)
DEFINE_KERNELS(c64, c64)
