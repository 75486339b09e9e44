// tqb_reduce.cu -- read-only passes over the state: norms, <Z> expectations, matrix-free
// Pauli sums, inner products, adjoint-gradient reductions, projection, and the blocked-CDF
// sampler.  All sums are float64 and two-stage (per-CTA partials in the library workspace,
// then one CTA per batch member adds them in a fixed order), so results are deterministic.
//
// Replaces (reference paths relative to src/tyxonq/):
//   expect_z_statevector                    libs/quantum_library/kernels/statevector.py:62-68
//   pauli_string_sum_dense + expectation    libs/quantum_library/kernels/pauli.py:74-87, dynamics.py:117-126
//   apply_op (densified sparse H)           applications/chem/chem_libs/hamiltonians_chem_library/hamiltonian_builders.py:283-318
//   _project_z                              devices/simulators/statevector/engine.py:1075-1087
//   sampling (Generator.choice + bincount)  devices/simulators/statevector/engine.py:377-418
#include <cuda_pipeline.h>
#include <cuda_runtime.h>

#include <string>

#include "tqb_core.cuh"
#include "tqb_host.h"

namespace tqb {

constexpr int RT = 256;         // threads per reduction CTA
constexpr int MAX_ACC = 40;     // values reduced per CTA at most (n bits or masks per launch)
constexpr int MASKS_PER_LAUNCH = 32;

// ---- block reduction of NV doubles per thread -> sm_out[0..NV) valid in thread 0 ----------
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double *sm /* [RT/32][NV] */, double *out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double x = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) sm[warp * NV + i] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0;
    for (int w = 0; w < RT / 32; ++w) s += sm[w * NV + threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

// grid = (nbx, batch); block x handles the contiguous segment [x*seg, (x+1)*seg) of member y.
struct Seg {
  int n;
  int seg_bits;  // log2(segment length)
};

template <typename T>
__device__ __forceinline__ double prob_of(const cplx<T> a) {
  const double re = (double)a.x, im = (double)a.y;
  return __dadd_rn(__dmul_rn(re, re), __dmul_rn(im, im));
}

// a float64 array of probabilities can stand in for a state in the sampler (TQB_F64: noise-mixed distributions)
__device__ __forceinline__ double prob_of(const double p) { return p; }

// partial[(y*nbx + x)*nv + j]
template <typename E>
__global__ void __launch_bounds__(RT) norm2_kernel(const E *__restrict__ state, Seg sg, double *partial) {
  __shared__ double sm[RT / 32];
  const E *s = state + ((size_t)blockIdx.y << sg.n) + ((size_t)blockIdx.x << sg.seg_bits);
  const size_t len = (size_t)1 << sg.seg_bits;
  double v[1] = {0.0};
  for (size_t i = threadIdx.x; i < len; i += RT) v[0] += prob_of(s[i]);
  block_reduce<1>(v, sm, partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x));
}

// <Z> on every local bit.  Elements visited by one thread are i = it*RT + tid, so bits below
// log2(RT) are fixed per thread, bits in [8, seg_bits) vary with `it`, bits >= seg_bits are
// fixed per CTA.
template <typename E>
__global__ void __launch_bounds__(RT) expect_z_bits_kernel(const E *__restrict__ state, Seg sg, double *partial) {
  __shared__ double sm[(RT / 32) * MAX_ACC];
  const E *s = state + ((size_t)blockIdx.y << sg.n) + ((size_t)blockIdx.x << sg.seg_bits);
  const size_t len = (size_t)1 << sg.seg_bits;
  constexpr int LT = 8;  // log2(RT)
  constexpr int MID = 16;
  double tot = 0.0;
  double ones[MID];  // probability mass with bit (LT + j) set
#pragma unroll
  for (int j = 0; j < MID; ++j) ones[j] = 0.0;
  const int nmid = sg.seg_bits > LT ? sg.seg_bits - LT : 0;
  constexpr int U = 3;  // 2^U loads in flight per thread; the low U bits of `it` are register indices
  if (nmid >= U) {
    double acc[1 << U];
#pragma unroll
    for (int u = 0; u < (1 << U); ++u) acc[u] = 0.0;
    const size_t iters = len / RT;
    for (size_t it0 = 0; it0 < iters; it0 += (1 << U)) {
      E a[1 << U];
#pragma unroll
      for (int u = 0; u < (1 << U); ++u) a[u] = s[(it0 + u) * RT + threadIdx.x];
      double blk = 0.0;
#pragma unroll
      for (int u = 0; u < (1 << U); ++u) {
        const double p = prob_of(a[u]);
        acc[u] += p;
        blk += p;
      }
      const uint32_t hi = (uint32_t)(it0 >> U);  // uniform across the CTA: the adds below are branch-free selects
#pragma unroll
      for (int j = U; j < MID; ++j)
        if (j < nmid) ones[j] += ((hi >> (j - U)) & 1u) ? blk : 0.0;
    }
#pragma unroll
    for (int u = 0; u < (1 << U); ++u) {
      tot += acc[u];
#pragma unroll
      for (int j = 0; j < U; ++j)
        if ((u >> j) & 1) ones[j] += acc[u];
    }
  } else {
    for (size_t it = 0; it * RT + threadIdx.x < len; ++it) {
      const double p = prob_of(s[it * RT + threadIdx.x]);
      tot += p;
#pragma unroll
      for (int j = 0; j < MID; ++j)
        if (j < nmid && ((it >> j) & 1)) ones[j] += p;
    }
  }
  // combine into per-bit contributions, MAX_ACC at most
  double v[MAX_ACC];
#pragma unroll
  for (int b = 0; b < MAX_ACC; ++b) {
    double c = 0.0;
    if (b < sg.n) {
      if (b < LT && b < sg.seg_bits) c = ((threadIdx.x >> b) & 1) ? -tot : tot;
      else if (b < sg.seg_bits) c = tot - 2.0 * ones[b - LT > 0 ? (b - LT < MID ? b - LT : MID - 1) : 0];
      else c = ((blockIdx.x >> (b - sg.seg_bits)) & 1) ? -tot : tot;
    }
    v[b] = c;
  }
  block_reduce<MAX_ACC>(v, sm, partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * MAX_ACC);
}

// Sum_i p_i (-1)^popc(idx_i & mask_t) for up to 32 masks per launch.  The parities of all masks are kept as ONE
// 32-bit sign word: sw(idx) = XOR over the set bits b of idx of col[b], col[b] bit t = bit b of mask t.  A thread
// visits idx = fixed | (it << LT); stepping it -> it+1 flips bits 0..ctz(it+1) of it, i.e. one XOR with a
// prefix word.  Per amplitude and mask that leaves a predicated add (ones[t] += p where the sign bit is set) instead
// of a 64-bit popc; result = total - 2 * ones[t].
template <typename T>
__global__ void __launch_bounds__(RT, 2) expect_zmasks_kernel(const cplx<T> *__restrict__ state, Seg sg, uint64_t global_base,
                                                              const uint64_t *__restrict__ masks, int n_masks, double *partial) {
  __shared__ double sm[(RT / 32) * MASKS_PER_LAUNCH];
  __shared__ uint32_t col[64];
  __shared__ uint32_t pre[64];  // pre[k] = col[LT+U] ^ .. ^ col[LT+U+k]
  constexpr int LT = 8;         // log2(RT)
  constexpr int U = 2;          // 2^U loads in flight per thread
  const size_t off = (size_t)blockIdx.x << sg.seg_bits;
  const cplx<T> *s = state + ((size_t)blockIdx.y << sg.n) + off;
  const size_t len = (size_t)1 << sg.seg_bits;
  const bool blocked = len >= ((size_t)RT << U);
  const int first = LT + (blocked ? U : 0);
  if (threadIdx.x < 64) {
    uint32_t w = 0;
    for (int t = 0; t < n_masks; ++t) w |= (uint32_t)((masks[t] >> threadIdx.x) & 1ull) << t;
    col[threadIdx.x] = w;
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    uint32_t w = 0;
    for (int j = 0; j <= (int)threadIdx.x && first + j < 64; ++j) w ^= col[first + j];
    pre[threadIdx.x] = w;
  }
  __syncthreads();
  const uint64_t fixed = global_base | (off + threadIdx.x);
  uint32_t sw0 = 0;
  for (int b = 0; b < 64; ++b)
    if ((fixed >> b) & 1ull) sw0 ^= col[b];
  double tot = 0.0;
  double ones[MASKS_PER_LAUNCH];
#pragma unroll
  for (int t = 0; t < MASKS_PER_LAUNCH; ++t) ones[t] = 0.0;
  if (blocked) {
    uint32_t eu[1 << U];
#pragma unroll
    for (int u = 0; u < (1 << U); ++u) {
      uint32_t w = 0;
#pragma unroll
      for (int j = 0; j < U; ++j)
        if ((u >> j) & 1) w ^= col[LT + j];
      eu[u] = w;
    }
    const size_t nblk = (len / RT) >> U;
    uint32_t g = sw0;
    for (size_t blk = 0; blk < nblk; ++blk) {
      if (blk) g ^= pre[__ffsll((long long)blk) - 1];
      cplx<T> a[1 << U];
#pragma unroll
      for (int u = 0; u < (1 << U); ++u) a[u] = s[((blk << U) + u) * RT + threadIdx.x];
#pragma unroll
      for (int u = 0; u < (1 << U); ++u) {
        const double p = prob_of(a[u]);
        const uint32_t sw = g ^ eu[u];
        tot += p;
#pragma unroll
        for (int t = 0; t < MASKS_PER_LAUNCH; ++t)
          if (sw & (1u << t)) ones[t] += p;
      }
    }
  } else {
    uint32_t g = sw0;
    for (size_t it = 0; it * RT + threadIdx.x < len; ++it) {
      if (it) g ^= pre[__ffsll((long long)it) - 1];
      const double p = prob_of(s[it * RT + threadIdx.x]);
      tot += p;
#pragma unroll
      for (int t = 0; t < MASKS_PER_LAUNCH; ++t)
        if (g & (1u << t)) ones[t] += p;
    }
  }
  double v[MASKS_PER_LAUNCH];
#pragma unroll
  for (int t = 0; t < MASKS_PER_LAUNCH; ++t) v[t] = tot - 2.0 * ones[t];
  block_reduce<MASKS_PER_LAUNCH>(v, sm, partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * MASKS_PER_LAUNCH);
}

// out[y*out_stride + j] = sum_x partial[(y*nbx + x)*nv + j], fixed order.  grid = batch.
__global__ void finish_kernel(const double *partial, int nbx, int nv, int n_out, double *out, int out_stride) {
  for (int j = threadIdx.x; j < n_out; j += blockDim.x) {
    double s = 0.0;
    for (int x = 0; x < nbx; ++x) s += partial[((size_t)blockIdx.x * nbx + x) * nv + j];
    out[(size_t)blockIdx.x * out_stride + j] = s;
  }
}

// ---- Pauli sums -------------------------------------------------------------------------
// sign-sum of a group's terms at index j: sum_t coef_t (-1)^popc(j & z_t)
__device__ __forceinline__ void group_phase(const uint64_t *__restrict__ term_z, const double *__restrict__ term_coef,
                                            int t0, int t1, uint64_t j, double &re, double &im) {
  re = 0.0; im = 0.0;
  for (int t = t0; t < t1; ++t) {
    const double sgn = (__popcll(j & term_z[t]) & 1) ? -1.0 : 1.0;
    re += sgn * term_coef[2 * t];
    im += sgn * term_coef[2 * t + 1];
  }
}

// <psi|H|psi> = sum_j sum_g conj(psi_{j^x_g}) * phase_g(j) * psi_j     (P|j> = phase(j)|j^x>)
template <typename T>
__global__ void __launch_bounds__(RT) expect_pauli_kernel(const cplx<T> *__restrict__ state, Seg sg, uint64_t global_base,
                                                          const uint64_t *__restrict__ group_x, const int *__restrict__ group_ptr,
                                                          int n_groups, const uint64_t *__restrict__ term_z,
                                                          const double *__restrict__ term_coef, double *partial) {
  __shared__ double sm[(RT / 32) * 2];
  const cplx<T> *sb = state + ((size_t)blockIdx.y << sg.n);
  const size_t off = (size_t)blockIdx.x << sg.seg_bits;
  const size_t len = (size_t)1 << sg.seg_bits;
  double v[2] = {0.0, 0.0};
  for (size_t i = threadIdx.x; i < len; i += RT) {
    const uint64_t j = off + i;
    const cplx<T> a = sb[j];
    const double ar = a.x, ai = a.y;
    double hr = 0.0, hi = 0.0;  // sum_g conj(psi_{j^x}) * phase_g(j)
    for (int g = 0; g < n_groups; ++g) {
      double pr, pi;
      group_phase(term_z, term_coef, group_ptr[g], group_ptr[g + 1], global_base | j, pr, pi);
      const cplx<T> b = sb[j ^ group_x[g]];
      const double br = b.x, bi = -(double)b.y;
      hr += br * pr - bi * pi;
      hi += br * pi + bi * pr;
    }
    v[0] += hr * ar - hi * ai;
    v[1] += hr * ai + hi * ar;
  }
  block_reduce<2>(v, sm, partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2);
}

// out_j = sum_g phase_g(j ^ x_g) * psi_{j ^ x_g}
template <typename T>
__global__ void __launch_bounds__(RT) apply_pauli_kernel(const cplx<T> *__restrict__ state, cplx<T> *__restrict__ out, int n,
                                                         uint64_t global_base, const uint64_t *__restrict__ group_x,
                                                         const int *__restrict__ group_ptr, int n_groups,
                                                         const uint64_t *__restrict__ term_z,
                                                         const double *__restrict__ term_coef, int groups_per_slice) {
  // groups_per_slice > 0 (small states: too few amplitudes to fill the device): blockIdx.z takes a slice of the x-mask
  // groups and ADDS its part into `out` (zeroed by the host) -- 554 groups of a molecular Hamiltonian on 2^14 amplitudes
  // were one serial loop per thread on 64 CTAs
  const cplx<T> *sb = state + ((size_t)blockIdx.y << n);
  cplx<T> *ob = out + ((size_t)blockIdx.y << n);
  const size_t dim = (size_t)1 << n;
  const int g0 = groups_per_slice > 0 ? (int)blockIdx.z * groups_per_slice : 0;
  const int g1 = groups_per_slice > 0 ? (g0 + groups_per_slice < n_groups ? g0 + groups_per_slice : n_groups) : n_groups;
  for (size_t j = (size_t)blockIdx.x * RT + threadIdx.x; j < dim; j += (size_t)gridDim.x * RT) {
    double hr = 0.0, hi = 0.0;
    for (int g = g0; g < g1; ++g) {
      const uint64_t src = j ^ group_x[g];
      double pr, pi;
      group_phase(term_z, term_coef, group_ptr[g], group_ptr[g + 1], global_base | src, pr, pi);
      const cplx<T> b = sb[src];
      const double br = b.x, bi = b.y;
      hr += br * pr - bi * pi;
      hi += br * pi + bi * pr;
    }
    if (groups_per_slice > 0) {
      atomicAdd(&ob[j].x, (T)hr);
      atomicAdd(&ob[j].y, (T)hi);
    } else {
      ob[j] = cplx<T>{(T)hr, (T)hi};
    }
  }
}

// ---- tile-staged Pauli sums ---------------------------------------------------------------------------------
// The kernel above gathers psi_{j ^ x_g} from global memory once per group and re-reads every term's mask and
// coefficient per amplitude.  Here a CTA stages a tile of 2^m amplitudes (the L lowest index bits + the high bits hb[],
// like a gate pass) in shared memory ONCE and evaluates every group whose xmask lies inside the tile bits from there:
//   * partner amplitudes come from shared memory (xmask in tile-local coordinates, xl);
//   * a term's z mask is split into its tile-local part zl (tile-local coordinates) and the part outside the tile,
//     whose parity is constant for the tile: it is folded into the coefficient once per tile (sc[]);
//   * HERM (every group is a Hermitian operator, i.e. real Pauli coefficients): the pair (j, j ^ x) contributes
//     2 Re(conj(psi_{j^x}) phase(j) psi_j), so only the half of the tile with the lowest xmask bit clear is visited.
// The host groups the xmasks into tile layouts (pauli.py plan_layouts): one read of the state per layout.
struct PauliTile {
  int n, m, L, h;
  int g0, ng;          // groups [g0, g0 + ng) of gxl / gptr
  int layout, n_layouts;
  int8_t hb[16];
};

// c with the sign (-1)^popc(v): the parity goes straight into the sign bit
__device__ __forceinline__ float flip_sign(float c, uint32_t v) { return __int_as_float(__float_as_int(c) ^ (int)(__popc(v) << 31)); }
__device__ __forceinline__ double flip_sign(double c, uint32_t v) {
  return __hiloint2double(__double2hiint(c) ^ (int)(__popc(v) << 31), __double2loint(c));
}

template <typename R> struct Pair2 { R x, y; };

// Enumeration of the tile indices whose bits `holes` are clear: thread t visits the (t + i * RT)-th such index for
// i = 0 .. iters - 1.  The first index is computed by zero insertion, every next one by ONE add that carries across the
// holes: e' = ((e | holes) + step) & ~holes with step = the image of RT (RT is a power of two, so the thread part of the
// counter never changes).  body(e) gets the index; the trip count is uniform over the CTA.
__device__ __forceinline__ uint32_t insert_holes(uint32_t x, uint32_t holes) {
  while (holes) {
    const uint32_t p = (uint32_t)(__ffs((int)holes) - 1);
    x = ((x >> p) << (p + 1u)) | (x & ((1u << p) - 1u));
    holes &= holes - 1u;
  }
  return x;
}
template <class Body>
__device__ __forceinline__ void for_indices(uint32_t holes, uint32_t iters, Body body) {
  uint32_t e = insert_holes(threadIdx.x, holes);
  const uint32_t step = insert_holes((uint32_t)RT, holes);
#pragma unroll 4
  for (uint32_t i = 0; i < iters; ++i) {
    body(e);
    e = ((e | holes) + step) & ~holes;
  }
}

// Arithmetic type R = the state's component type (complex64 states: float products and float partial sums over the few
// iterations a thread spends on one group of one tile, added to the float64 accumulator once per group and tile).
template <typename T, bool HERM, bool REALC>   // REALC: every coefficient (i^#Y folded in) is real
__global__ void __launch_bounds__(RT) expect_pauli_tiled_kernel(const cplx<T> *__restrict__ state, PauliTile pt, uint64_t global_base,
                                                                const uint32_t *__restrict__ gxl, const int *__restrict__ gptr,
                                                                const uint32_t *__restrict__ zl, const uint64_t *__restrict__ zout,
                                                                const double *__restrict__ coef, double *partial) {
  typedef T R;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[(RT / 32) * 2];
  const int m = pt.m, L = pt.L, h = pt.h;
  const uint32_t nel = 1u << m;
  cplx<T> *tile = reinterpret_cast<cplx<T> *>(smem_raw);
  uint64_t *roff = reinterpret_cast<uint64_t *>(tile + nel);
  const int t0 = gptr[pt.g0], nt = gptr[pt.g0 + pt.ng] - t0;
  Pair2<R> *sc = reinterpret_cast<Pair2<R> *>(roff + (((1u << h) + 1u) & ~1u));   // signed coefficients of this tile
  uint32_t *szl = reinterpret_cast<uint32_t *>(sc + nt);
  uint32_t *sgx = szl + nt;               // per group: xmask (tile-local) ...
  int *sgp = reinterpret_cast<int *>(sgx + pt.ng);   // ... and term range (relative), ng + 1 entries
  // groups of more than two terms: the sign sum S(e) = sum_t c_t (-1)^popc(e & z_t) splits into a table over the low LB tile
  // bits, a table over the high ones and the few terms whose z-mask straddles the split (listed in sst)
  constexpr int LB = 7;
  int *sst = sgp + pt.ng + 2;
  // [0, 128): low table, [128, 256): high table; 16-byte aligned by OFFSET arithmetic (the address space stays known: LDS)
  const uint32_t sd_off = (uint32_t)((reinterpret_cast<unsigned char *>(sst + nt) - smem_raw) + 15) & ~15u;
  Pair2<R> *sd = reinterpret_cast<Pair2<R> *>(smem_raw + sd_off);
  __shared__ int sst_n;
  const int tid = threadIdx.x;
  for (uint32_t j = tid; j < (1u << h); j += RT) {
    uint64_t o = 0;
    for (int i = 0; i < h; ++i) o |= (uint64_t)((j >> i) & 1u) << pt.hb[i];
    roff[j] = o;
  }
  for (int t = tid; t < nt; t += RT) szl[t] = zl[t0 + t];
  for (int g = tid; g < pt.ng; g += RT) sgx[g] = gxl[pt.g0 + g];
  for (int g = tid; g <= pt.ng; g += RT) sgp[g] = gptr[pt.g0 + g] - t0;
  const cplx<T> *sb = state + ((size_t)blockIdx.y << pt.n);
  const uint64_t ntiles = 1ull << (pt.n - m);
  double acc[2] = {0.0, 0.0};
  constexpr int V = 16 / (int)sizeof(cplx<T>);
  for (uint64_t tt = blockIdx.x; tt < ntiles; tt += gridDim.x) {
    uint64_t base = tt << L;
    for (int j = 0; j < h; ++j) {
      const uint32_t p = (uint32_t)pt.hb[j];
      base = ((base >> p) << (p + 1u)) | (base & ((1ull << p) - 1ull));
    }
    __syncthreads();   // the previous tile is done with tile[] / sc[]
    // the tile: 16-byte asynchronous copies global -> shared (LDGSTS), all of a thread's copies in flight at once -- the
    // staged loop through registers kept ~4 loads per thread in flight and paid four memory round trips per tile
    for (uint32_t i = tid; i < nel / V; i += RT) {
      const uint32_t e = i * V;
      const uint64_t idx = base | roff[e >> L] | (uint64_t)(e & ((1u << L) - 1u));
      __pipeline_memcpy_async(tile + e, sb + idx, 16);
    }
    __pipeline_commit();
    const uint64_t gidx = global_base | base;
    for (int t = tid; t < nt; t += RT) {
      const bool odd = __popcll(gidx & zout[t0 + t]) & 1;
      const double cr = coef[2 * (t0 + t)], ci = coef[2 * (t0 + t) + 1];
      sc[t] = odd ? Pair2<R>{(R)-cr, (R)-ci} : Pair2<R>{(R)cr, (R)ci};
    }
    __pipeline_wait_prior(0);
    __syncthreads();
    // groups outside, the thread's amplitudes inside: a group's metadata is read once, nothing is carried from one
    // amplitude to the next but the running sums, and the inner iterations are independent (instruction-level overlap)
    for (int g = 0; g < pt.ng; ++g) {
      const uint32_t xl = sgx[g];
      const int a0 = sgp[g], a1 = sgp[g + 1];
      const bool paired = HERM && xl != 0u;
      // a Hermitian group visits each pair (e, e ^ x) once: e runs over the indices with the lowest xmask bit clear
      const uint32_t pbit = paired ? (uint32_t)(__ffs((int)xl) - 1) : 0u;
      const uint32_t phole = paired ? 1u << pbit : 0u;
      const uint32_t count = paired ? nel >> 1 : nel;
      R sr = 0, si = 0;
      // visit the `nitems` tile indices whose `holes` bits are clear (for_indices: one add per index, uniform trip count;
      // tiles smaller than 4 * RT amplitudes take the plain loop)
      auto visit = [&](uint32_t holes, uint32_t nitems, auto body) {
        if (nel >= 4u * RT) {
          for_indices(holes, nitems / RT, body);
        } else {
          for (uint32_t k = tid; k < nitems; k += RT) body(insert_holes(k, holes));
        }
      };
      // Two real terms whose z-masks differ in ONE tile bit j (XX + YY of a bond, hopping terms X Z..Z X + Y Z..Z Y): the
      // sign sum is (cn + cj (-1)^e_j) (-1)^popc(e & zc).  With |cn| = |cj| it vanishes on half of the pairs -- those are not
      // visited at all, and the other half shares one coefficient.  (Decided per tile: the signs of the outside-the-tile
      // bits are in sc[].)
      bool half_done = false;
      if (REALC && HERM && paired && a1 - a0 == 2) {
        const uint32_t pm = ~(1u << pbit);
        const uint32_t za = szl[a0] & pm, zb = szl[a0 + 1] & pm, dz = za ^ zb;
        if (__popc(dz) == 1) {
          const uint32_t jb = (uint32_t)(__ffs((int)dz) - 1);
          const bool a_has = (za >> jb) & 1u;
          const R cj = a_has ? sc[a0].x : sc[a0 + 1].x, cn = a_has ? sc[a0 + 1].x : sc[a0].x;
          const R v0 = cn + cj, v1 = cn - cj;
          if (v0 == (R)0 || v1 == (R)0) {
            const uint32_t fix = v0 == (R)0 ? 1u << jb : 0u;   // bit j of the pairs that contribute
            const R C = v0 == (R)0 ? v1 : v0;
            const uint32_t zc = za & zb;
            const uint32_t holes = phole | (1u << jb);
            if (zc == 0u) {
              visit(holes, nel >> 2, [&](uint32_t e0) {
                const uint32_t e = e0 | fix;
                const cplx<T> a = tile[e], b = tile[e ^ xl];
                sr += b.x * a.x + b.y * a.y;
              });
            } else {
              visit(holes, nel >> 2, [&](uint32_t e0) {
                const uint32_t e = e0 | fix;
                const cplx<T> a = tile[e], b = tile[e ^ xl];
                sr += flip_sign(b.x * a.x + b.y * a.y, e & zc);
              });
            }
            sr *= C;
            half_done = true;
          }
        }
      }
      if (half_done) {
      } else if (REALC && a1 - a0 <= 2) {
        // one or two real terms (XX + YY of a Heisenberg bond, a single string): everything in registers
        const R c0 = sc[a0].x, c1 = a1 - a0 == 2 ? sc[a0 + 1].x : (R)0;
        const uint32_t z0 = szl[a0], z1 = a1 - a0 == 2 ? szl[a0 + 1] : 0u;
        visit(phole, count, [&](uint32_t e) {
          const R pr = flip_sign(c0, e & z0) + flip_sign(c1, e & z1);
          const cplx<T> a = tile[e], b = tile[e ^ xl];
          sr += (b.x * a.x + b.y * a.y) * pr;
          if (!HERM) si += (b.x * a.y - b.y * a.x) * pr;
        });
      } else {
        // tables of the sign sum for this tile (the signs of the outside-the-tile bits are already in sc[])
        const uint32_t lowmask = m > LB ? (1u << LB) - 1u : nel - 1u;
        const uint32_t nhigh = m > LB ? 1u << (m - LB) : 1u;
        __syncthreads();   // (the previous group's tables are no longer read)
        for (uint32_t j = tid; j < 128u + nhigh; j += RT) {
          const bool low = j < 128u;
          const uint32_t idx = low ? j : (j - 128u) << LB;
          R pr = 0, pi = 0;
          for (int t = a0; t < a1; ++t) {
            const uint32_t z = szl[t];
            // low table: masks inside the low bits (and the constant terms); high table: masks inside the high bits
            const bool mine = low ? (z & ~lowmask) == 0u : ((z & lowmask) == 0u && z != 0u);
            if (mine) {
              const Pair2<R> c = sc[t];
              pr += flip_sign(c.x, idx & z);
              if (!REALC) pi += flip_sign(c.y, idx & z);
            }
          }
          sd[j] = Pair2<R>{pr, pi};
        }
        if (tid < 32) {   // the straddling terms, in order
          int cnt = 0;
          for (int tb = a0; tb < a1; tb += 32) {
            const int t = tb + tid;
            const bool strad = t < a1 && (szl[t < a1 ? t : a0] & lowmask) != 0u && (szl[t < a1 ? t : a0] & ~lowmask) != 0u;
            const unsigned bal = __ballot_sync(0xffffffffu, strad);
            if (strad) sst[cnt + __popc(bal & ((1u << tid) - 1u))] = t;
            cnt += __popc(bal);
          }
          if (tid == 0) sst_n = cnt;
        }
        __syncthreads();
        const int nst = sst_n;
        const bool diag = xl == 0u;
        // the first two straddling terms live in registers (a nearest-neighbour chain has one), the rest in a loop
        const Pair2<R> s0c = nst > 0 ? sc[sst[0]] : Pair2<R>{(R)0, (R)0}, s1c = nst > 1 ? sc[sst[1]] : Pair2<R>{(R)0, (R)0};
        const uint32_t s0z = nst > 0 ? szl[sst[0]] : 0u, s1z = nst > 1 ? szl[sst[1]] : 0u;
        visit(phole, count, [&](uint32_t e) {
          const Pair2<R> dl = sd[e & lowmask], dh = sd[128u + (m > LB ? e >> LB : 0u)];
          R pr = dl.x + dh.x + flip_sign(s0c.x, e & s0z) + flip_sign(s1c.x, e & s1z), pi = dl.y + dh.y;
          if (!REALC) pi += flip_sign(s0c.y, e & s0z) + flip_sign(s1c.y, e & s1z);
          for (int i = 2; i < nst; ++i) {
            const int t = sst[i];
            const Pair2<R> c = sc[t];
            const uint32_t v = e & szl[t];
            pr += flip_sign(c.x, v);
            if (!REALC) pi += flip_sign(c.y, v);
          }
          const cplx<T> a = tile[e], b = diag ? a : tile[e ^ xl];
          const R qr = b.x * a.x + b.y * a.y, qi = b.x * a.y - b.y * a.x;   // conj(b) * a
          if (REALC) {
            sr += qr * pr;
            if (!HERM) si += qi * pr;
          } else {
            sr += qr * pr - qi * pi;
            if (!HERM) si += qr * pi + qi * pr;
          }
        });
      }
      const double w = paired ? 2.0 : 1.0;
      acc[0] += w * (double)sr;
      if (!HERM) acc[1] += w * (double)si;
    }
  }
  block_reduce<2>(acc, red, partial + (((size_t)blockIdx.y * pt.n_layouts + pt.layout) * gridDim.x + blockIdx.x) * 2);
}

// ---- single-qubit transition matrices of a (bra, ket) pair -------------------------------------------------------
// T_q[a][b] = sum over the other bits of conj(bra[.., a, ..]) * ket[.., b, ..]  for up to 6 index bits q per launch: every
// single-qubit operator A on qubit q has <bra| A_q |ket> = sum_ab A[a][b] T_q[a][b], so ONE evaluation of T_q yields the adjoint
// gradient of every gate in a layer of single-qubit gates on q (generators conjugated through the later gates of the run on
// the host) -- instead of one full read of both states per parameter (grad_dense_kernel).  Tile-staged like the Pauli sums:
// both states' tiles in shared memory, the bits of interest are tile bits.
struct Trans1q {
  int n, m, L, h, nb;
  int8_t hb[16];
  int8_t tb[8];   // tile-local positions of the bits of interest
};

template <typename T, int NB>
__global__ void __launch_bounds__(RT, 2) transition_1q_kernel(const cplx<T> *__restrict__ bra, const cplx<T> *__restrict__ ket, Trans1q tq,
                                                              double *out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[RT / 32][NB * 8];
  const int m = tq.m, L = tq.L, h = tq.h;
  const uint32_t nel = 1u << m;
  cplx<T> *kt = reinterpret_cast<cplx<T> *>(smem_raw);
  cplx<T> *bt = kt + nel;
  uint64_t *roff = reinterpret_cast<uint64_t *>(bt + nel);
  const int tid = threadIdx.x;
  for (uint32_t j = tid; j < (1u << h); j += RT) {
    uint64_t o = 0;
    for (int i = 0; i < h; ++i) o |= (uint64_t)((j >> i) & 1u) << tq.hb[i];
    roff[j] = o;
  }
  uint32_t pos[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) pos[k] = (uint32_t)tq.tb[k < tq.nb ? k : 0];
  double acc[NB][8];
#pragma unroll
  for (int k = 0; k < NB; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[k][i] = 0.0;
  const uint64_t ntiles = 1ull << (tq.n - m);
  constexpr int V = 16 / (int)sizeof(cplx<T>);
  for (uint64_t tt = blockIdx.x; tt < ntiles; tt += gridDim.x) {
    uint64_t base = tt << L;
    for (int j = 0; j < h; ++j) {
      const uint32_t p = (uint32_t)tq.hb[j];
      base = ((base >> p) << (p + 1u)) | (base & ((1ull << p) - 1ull));
    }
    __syncthreads();
    for (uint32_t i = tid; i < nel / V; i += RT) {   // asynchronous 16-byte copies: all of them in flight at once
      const uint32_t e = i * V;
      const uint64_t idx = base | roff[e >> L] | (uint64_t)(e & ((1u << L) - 1u));
      __pipeline_memcpy_async(kt + e, ket + idx, 16);
      __pipeline_memcpy_async(bt + e, bra + idx, 16);
    }
    __pipeline_commit();
    __pipeline_wait_prior(0);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      if (k >= tq.nb) break;
      const uint32_t low = (1u << pos[k]) - 1u;
      for (uint32_t c = tid; c < (nel >> 1); c += RT) {
        const uint32_t e0 = ((c & ~low) << 1) | (c & low), e1 = e0 | (1u << pos[k]);
        const cplx<T> b0 = bt[e0], b1 = bt[e1], k0 = kt[e0], k1 = kt[e1];
        // conj(b) * k
        acc[k][0] += (double)b0.x * k0.x + (double)b0.y * k0.y;  acc[k][1] += (double)b0.x * k0.y - (double)b0.y * k0.x;   // T00
        acc[k][2] += (double)b0.x * k1.x + (double)b0.y * k1.y;  acc[k][3] += (double)b0.x * k1.y - (double)b0.y * k1.x;   // T01
        acc[k][4] += (double)b1.x * k0.x + (double)b1.y * k0.y;  acc[k][5] += (double)b1.x * k0.y - (double)b1.y * k0.x;   // T10
        acc[k][6] += (double)b1.x * k1.x + (double)b1.y * k1.y;  acc[k][7] += (double)b1.x * k1.y - (double)b1.y * k1.x;   // T11
      }
    }
  }
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int k = 0; k < NB; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double x = acc[k][i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (lane == 0) red[warp][k * 8 + i] = x;
    }
  __syncthreads();
  if (tid < tq.nb * 8) {
    double s2 = 0.0;
    for (int w = 0; w < RT / 32; ++w) s2 += red[w][tid];
    atomicAdd(out + tid, s2);
  }
}

template <typename T>
__global__ void __launch_bounds__(RT) inner_kernel(const cplx<T> *__restrict__ a, const cplx<T> *__restrict__ b, Seg sg,
                                                   double *partial) {
  __shared__ double sm[(RT / 32) * 2];
  const size_t base = ((size_t)blockIdx.y << sg.n) + ((size_t)blockIdx.x << sg.seg_bits);
  const size_t len = (size_t)1 << sg.seg_bits;
  double v[2] = {0.0, 0.0};
  for (size_t i = threadIdx.x; i < len; i += RT) {
    const cplx<T> x = a[base + i], y = b[base + i];
    const double xr = x.x, xi = x.y, yr = y.x, yi = y.y;
    v[0] += xr * yr + xi * yi;
    v[1] += xr * yi - xi * yr;
  }
  block_reduce<2>(v, sm, partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2);
}

// Reduced density matrix of one index bit per batch member: rho00 = sum |psi_0|^2, rho11 = sum |psi_1|^2,
// rho01 = sum psi_0 conj(psi_1), over all pairs (i, i | 1 << bit).  The segment covers PAIR indices (n - 1 bits).
template <typename T>
__global__ void __launch_bounds__(RT) reduced_1q_kernel(const cplx<T> *__restrict__ state, int n, int bit, Seg sg, double *partial) {
  __shared__ double sm[(RT / 32) * 4];
  const cplx<T> *s = state + ((size_t)blockIdx.y << n);
  const size_t p0 = (size_t)blockIdx.x << sg.seg_bits;
  const size_t len = (size_t)1 << sg.seg_bits;
  const size_t low = ((size_t)1 << bit) - 1;
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (size_t i = threadIdx.x; i < len; i += RT) {
    const size_t p = p0 + i;
    const size_t i0 = ((p & ~low) << 1) | (p & low);
    const cplx<T> a = s[i0], b = s[i0 | ((size_t)1 << bit)];
    const double ar = a.x, ai = a.y, br = b.x, bi = b.y;
    v[0] += ar * ar + ai * ai;
    v[1] += br * br + bi * bi;
    v[2] += ar * br + ai * bi;
    v[3] += ai * br - ar * bi;
  }
  block_reduce<4>(v, sm, partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4);
}

// ---- adjoint-gradient reductions ------------------------------------------------------------
struct PairGen {
  int n, k;
  int8_t sbits[TQB_MAX_GATE_BITS];
  uint64_t off_a, off_b, zmask;
};

template <typename T>
__global__ void __launch_bounds__(RT) grad_pair_kernel(const cplx<T> *__restrict__ bra, const cplx<T> *__restrict__ ket,
                                                       PairGen pg, double scale, double *out) {
  __shared__ double sm[RT / 32];
  const uint64_t ngroups = 1ull << (pg.n - pg.k);
  double v[1] = {0.0};
  for (uint64_t g = (uint64_t)blockIdx.x * RT + threadIdx.x; g < ngroups; g += (uint64_t)gridDim.x * RT) {
    const uint64_t base = insert_zeros64(g, pg.sbits, pg.k);
    const cplx<T> ba = bra[base + pg.off_a], bb = bra[base + pg.off_b];
    const cplx<T> ka = ket[base + pg.off_a], kb = ket[base + pg.off_b];
    // Re( conj(bra_A) ket_B - conj(bra_B) ket_A )
    double r = ((double)ba.x * kb.x + (double)ba.y * kb.y) - ((double)bb.x * ka.x + (double)bb.y * ka.y);
    if (__popcll(base & pg.zmask) & 1) r = -r;
    v[0] += r;
  }
  double tot[1];
  block_reduce<1>(v, sm, tot);
  if (threadIdx.x == 0) atomicAdd(out, scale * tot[0]);
}

struct DenseGen {
  int n, k;
  int8_t bits[2], sbits[2];
  double gen[32];  // 2^k x 2^k complex, row-major, (re, im)
};

template <typename T>
__global__ void __launch_bounds__(RT) grad_dense_kernel(const cplx<T> *__restrict__ bra, const cplx<T> *__restrict__ ket,
                                                        DenseGen dg, double scale, double *out) {
  __shared__ double sm[RT / 32];
  const int D = 1 << dg.k;
  const uint64_t ngroups = 1ull << (dg.n - dg.k);
  double v[1] = {0.0};
  for (uint64_t g = (uint64_t)blockIdx.x * RT + threadIdx.x; g < ngroups; g += (uint64_t)gridDim.x * RT) {
    const uint64_t base = insert_zeros64(g, dg.sbits, dg.k);
    double kr[4], ki[4];
    uint64_t off[4];
    for (int s = 0; s < D; ++s) {
      uint64_t o = 0;
      for (int j = 0; j < dg.k; ++j) o |= (uint64_t)((s >> j) & 1) << dg.bits[j];
      off[s] = o;
      const cplx<T> kk = ket[base + o];
      kr[s] = kk.x; ki[s] = kk.y;
    }
    double acc = 0.0;
    for (int r = 0; r < D; ++r) {
      double dr = 0.0, di = 0.0;  // (D ket)_r
      for (int c = 0; c < D; ++c) {
        const double gr = dg.gen[2 * (r * D + c)], gi = dg.gen[2 * (r * D + c) + 1];
        dr += gr * kr[c] - gi * ki[c];
        di += gr * ki[c] + gi * kr[c];
      }
      const cplx<T> bb = bra[base + off[r]];
      acc += (double)bb.x * dr + (double)bb.y * di;  // Re(conj(bra) * (D ket))
    }
    v[0] += acc;
  }
  double tot[1];
  block_reduce<1>(v, sm, tot);
  if (threadIdx.x == 0) atomicAdd(out, scale * tot[0]);
}

// ---- persistent PAIR sweep for small states ------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned long long *counter, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1ull);
    while (*((volatile unsigned long long *)counter) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(RT) pair_sweep_kernel(cplx<T> *ket, cplx<T> *bra, int n, const tqb_pair_step *__restrict__ steps,
                                                        int n_steps, int mode, double *out, unsigned long long *sync) {
  __shared__ double sm[RT / 32];
  __shared__ tqb_pair_step st;
  const unsigned long long nblocks = gridDim.x;
  for (int j = 0; j < n_steps; ++j) {
    if (threadIdx.x < (int)(sizeof(tqb_pair_step) / 4))
      reinterpret_cast<uint32_t *>(&st)[threadIdx.x] = reinterpret_cast<const uint32_t *>(steps + j)[threadIdx.x];
    __syncthreads();
    const int k = st.k;
    const uint64_t ngroups = 1ull << (n - k);
    const double c = st.c, s = st.s;
    double v[1] = {0.0};
    for (uint64_t g = (uint64_t)blockIdx.x * RT + threadIdx.x; g < ngroups; g += nblocks * RT) {
      const uint64_t base = insert_zeros64(g, st.sbits, k);
      const uint64_t ia = base + st.off_a, ib = base + st.off_b;
      const bool odd = __popcll(base & st.zmask) & 1;
      const double sg = odd ? -s : s;
      cplx<T> ka = ket[ia], kb = ket[ib];
      if (mode == 1) {
        cplx<T> ba = bra[ia], bb = bra[ib];
        double r = ((double)ba.x * kb.x + (double)ba.y * kb.y) - ((double)bb.x * ka.x + (double)bb.y * ka.y);
        v[0] += odd ? -r : r;
        bra[ia] = cplx<T>{(T)(c * ba.x - sg * bb.x), (T)(c * ba.y - sg * bb.y)};
        bra[ib] = cplx<T>{(T)(sg * ba.x + c * bb.x), (T)(sg * ba.y + c * bb.y)};
      }
      ket[ia] = cplx<T>{(T)(c * ka.x - sg * kb.x), (T)(c * ka.y - sg * kb.y)};
      ket[ib] = cplx<T>{(T)(sg * ka.x + c * kb.x), (T)(sg * ka.y + c * kb.y)};
    }
    if (mode == 1) {
      double tot[1];
      block_reduce<1>(v, sm, tot);
      if (threadIdx.x == 0 && tot[0] != 0.0) atomicAdd(out + st.slot, st.scale * tot[0]);
    }
    if (j + 1 < n_steps) grid_barrier(sync, (unsigned long long)(j + 1) * nblocks);
  }
}

// The same sweep for registers of <= 15 qubits in ONE CTA of 1024 threads: a step has at most 2^13 independent groups (8 per
// thread), so the barrier between two dependent steps is a __syncthreads() (tens of cycles) instead of a grid barrier
// through global atomics (~2 us), the next step's descriptor is fetched while the current one computes, and an evaluation
// occupies one SM -- many independent evaluations run side by side (ucc.energy_and_grad_batch).
constexpr int SWEEP_CTA = 1024;
template <typename T>
__global__ void __launch_bounds__(SWEEP_CTA, 1) pair_sweep_cta_kernel(cplx<T> *ket, cplx<T> *bra, int n, const tqb_pair_step *__restrict__ steps,
                                                                     int n_steps, int mode, double *out) {
  __shared__ double sm[2][SWEEP_CTA / 32];   // per-warp partial sums, double-buffered over the steps
  __shared__ tqb_pair_step st2[2];
  const int tid = threadIdx.x;
  constexpr int WORDS = (int)(sizeof(tqb_pair_step) / 4);
  if (tid < WORDS) reinterpret_cast<uint32_t *>(&st2[0])[tid] = reinterpret_cast<const uint32_t *>(steps)[tid];
  __syncthreads();
  for (int j = 0; j < n_steps; ++j) {
    const tqb_pair_step &st = st2[j & 1];
    if (j + 1 < n_steps && tid < WORDS)   // (slot (j + 1) & 1 was last read in step j - 1, before the barrier that ended it)
      reinterpret_cast<uint32_t *>(&st2[(j + 1) & 1])[tid] = reinterpret_cast<const uint32_t *>(steps + j + 1)[tid];
    const int k = st.k;
    const uint32_t ngroups = 1u << (n - k);
    const double c = st.c, s = st.s;
    double v = 0.0;
    for (uint32_t g = tid; g < ngroups; g += SWEEP_CTA) {
      const uint64_t base = insert_zeros64(g, st.sbits, k);
      const uint64_t ia = base + st.off_a, ib = base + st.off_b;
      const bool odd = __popcll(base & st.zmask) & 1;
      const double sg = odd ? -s : s;
      cplx<T> ka = ket[ia], kb = ket[ib];
      if (mode == 1) {
        cplx<T> ba = bra[ia], bb = bra[ib];
        double r = ((double)ba.x * kb.x + (double)ba.y * kb.y) - ((double)bb.x * ka.x + (double)bb.y * ka.y);
        v += odd ? -r : r;
        bra[ia] = cplx<T>{(T)(c * ba.x - sg * bb.x), (T)(c * ba.y - sg * bb.y)};
        bra[ib] = cplx<T>{(T)(sg * ba.x + c * bb.x), (T)(sg * ba.y + c * bb.y)};
      }
      ket[ia] = cplx<T>{(T)(c * ka.x - sg * kb.x), (T)(c * ka.y - sg * kb.y)};
      ket[ib] = cplx<T>{(T)(sg * ka.x + c * kb.x), (T)(sg * ka.y + c * kb.y)};
    }
    if (mode == 1) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if ((tid & 31) == 0) sm[j & 1][tid >> 5] = v;
    }
    __syncthreads();   // the step's writes (global memory and sm[]) are visible to the CTA; st2[(j + 1) & 1] is loaded
    if (mode == 1 && tid < 32) {
      double t = sm[j & 1][tid];   // (rewritten in step j + 2, after the barrier of step j + 1 that this warp takes part in)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
      if (tid == 0 && t != 0.0) out[st.slot] += st.scale * t;   // (one CTA owns `out` during the sweep)
    }
  }
}


// ---- projection / scaling / probabilities ----------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(RT) project_kernel(cplx<T> *state, Seg sg, int bit, int keep, double *partial) {
  __shared__ double sm[RT / 32];
  const size_t off = (size_t)blockIdx.x << sg.seg_bits;
  cplx<T> *s = state + ((size_t)blockIdx.y << sg.n) + off;
  const size_t len = (size_t)1 << sg.seg_bits;
  double v[1] = {0.0};
  for (size_t i = threadIdx.x; i < len; i += RT) {
    const int b = (int)(((off + i) >> bit) & 1);
    if (b == keep) v[0] += prob_of(s[i]);
    else s[i] = cplx<T>{0, 0};
  }
  block_reduce<1>(v, sm, partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x));
}

// state_b *= 1/sqrt(norm2[b]) when norm2[b] > 0
template <typename T>
__global__ void __launch_bounds__(RT) renorm_kernel(cplx<T> *state, int n, const double *norm2) {
  const double nn = norm2[blockIdx.y];
  if (!(nn > 0.0)) return;
  const double f = 1.0 / sqrt(nn);
  cplx<T> *s = state + ((size_t)blockIdx.y << n);
  const size_t dim = (size_t)1 << n;
  for (size_t i = (size_t)blockIdx.x * RT + threadIdx.x; i < dim; i += (size_t)gridDim.x * RT) {
    cplx<T> a = s[i];
    a.x = (T)((double)a.x * f);
    a.y = (T)((double)a.y * f);
    s[i] = a;
  }
}

template <typename T>
__global__ void __launch_bounds__(RT) scale_kernel(cplx<T> *state, size_t total, double f) {
  for (size_t i = (size_t)blockIdx.x * RT + threadIdx.x; i < total; i += (size_t)gridDim.x * RT) {
    cplx<T> a = state[i];
    a.x = (T)((double)a.x * f);
    a.y = (T)((double)a.y * f);
    state[i] = a;
  }
}

template <typename T>
__global__ void __launch_bounds__(RT) probs_kernel(const cplx<T> *__restrict__ state, size_t total, double *out) {
  for (size_t i = (size_t)blockIdx.x * RT + threadIdx.x; i < total; i += (size_t)gridDim.x * RT) out[i] = prob_of(state[i]);
}

// Diagonal of a density matrix stored as a 2n-bit vector (row bits high): out[b * 2^n + i] = Re rho_b[i, i].
template <typename T>
__global__ void __launch_bounds__(RT) dm_diag_kernel(const cplx<T> *__restrict__ rho, int n, size_t total, double *out) {
  const size_t dim = (size_t)1 << n;
  for (size_t j = (size_t)blockIdx.x * RT + threadIdx.x; j < total; j += (size_t)gridDim.x * RT) {
    const size_t b = j >> n, i = j & (dim - 1);
    out[j] = (double)rho[(b << (2 * n)) + i * (dim + 1)].x;
  }
}

// ---- sampler ---------------------------------------------------------------------------------
// Chunk totals in the documented order: strictly sequential float64 sum of p_i over the chunk.
// One warp owns 32 consecutive chunks; global loads are coalesced (32 consecutive amplitudes of
// one chunk per instruction) and transposed through shared memory so that lane l adds the
// elements of chunk l in index order.
constexpr int CW = 4;  // warps per CTA
template <typename E>
__global__ void __launch_bounds__(CW * 32) chunk_totals_kernel(const E *__restrict__ state, int n, int chunk_bits,
                                                               long long n_chunks_total, double *totals, double *sub) {
  // sub != nullptr (chunk_bits == 12 only): sub[chunk * 32 + b] = the chunk's sequential running sum after its first
  // 128 (b + 1) amplitudes -- the SAME partial sums the sequential in-chunk scan passes through, recorded on the way, so
  // that a sample can enter its chunk at a 128-amplitude boundary (sample_kernel) without changing one rounding.
  __shared__ double buf[CW][32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long group = (long long)blockIdx.x * CW + w;  // 32 chunks per group
  const long long c0 = group * 32;
  if (c0 >= n_chunks_total) return;
  const size_t clen = (size_t)1 << chunk_bits;
  double acc = 0.0;
  for (size_t s0 = 0; s0 < clen; s0 += 32) {
    const int width = (clen - s0) < 32 ? (int)(clen - s0) : 32;
    for (int c = 0; c < 32; ++c) {
      double p = 0.0;
      if (c0 + c < n_chunks_total && lane < width) p = prob_of(state[((size_t)(c0 + c) << chunk_bits) + s0 + lane]);
      buf[w][c][lane] = p;
    }
    __syncwarp();
    for (int e = 0; e < width; ++e) acc = __dadd_rn(acc, buf[w][lane][e]);
    if (sub && (s0 & 127) == 96 && c0 + lane < n_chunks_total) sub[(size_t)(c0 + lane) * 32 + (s0 >> 7)] = acc;
    __syncwarp();
  }
  if (c0 + lane < n_chunks_total) totals[c0 + lane] = acc;
}

// prefix[b*(nc+1) + c] = sum of totals of chunks < c (sequential), one CTA per batch member.
__global__ void __launch_bounds__(RT) chunk_prefix_kernel(const double *totals, long long nc, double *prefix) {
  __shared__ double sm[2048];
  const double *t = totals + (size_t)blockIdx.x * nc;
  double *p = prefix + (size_t)blockIdx.x * (nc + 1);
  __shared__ double carry;
  if (threadIdx.x == 0) { carry = 0.0; p[0] = 0.0; }
  __syncthreads();
  for (long long c0 = 0; c0 < nc; c0 += 2048) {
    const int cnt = (nc - c0) < 2048 ? (int)(nc - c0) : 2048;
    for (int i = threadIdx.x; i < cnt; i += RT) sm[i] = t[c0 + i];
    __syncthreads();
    if (threadIdx.x == 0) {
      double run = carry;
      for (int i = 0; i < cnt; ++i) { run = __dadd_rn(run, sm[i]); sm[i] = run; }
      carry = run;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += RT) p[c0 + i + 1] = sm[i];
    __syncthreads();
  }
}

// A state array holds chunks [c_first, c_first + c_count) of the nc chunks the prefix covers (a shard of a state
// distributed over ranks; the whole state when c_first = 0, c_count = nc).  Samples that fall into other chunks
// are written as -1; `tail_index` >= 0 is written for u beyond the last chunk (the owner of the last chunk passes
// the dimension, everyone else -1).
template <typename E>
__global__ void __launch_bounds__(128) sample_kernel(const E *__restrict__ state, int n, int chunk_bits, long long nc,
                                                     long long c_first, long long c_count, long long tail_index,
                                                     const double *__restrict__ prefix, const double *__restrict__ uniforms,
                                                     long long shots, long long *idx_out, const double *__restrict__ sub) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= shots) return;
  const long long b = blockIdx.y;
  const double *p = prefix + (size_t)b * (nc + 1);
  const double total = p[nc];
  const double u = uniforms[(size_t)b * shots + s];
  // c = #{ c : cdf_end(c) / total <= u },  cdf_end(c) = p[c+1]
  long long lo = 0, hi = nc;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (__ddiv_rn(p[mid + 1], total) <= u) lo = mid + 1; else hi = mid;
  }
  long long idx;
  if (lo >= nc) {
    idx = tail_index;
  } else if (lo < c_first || lo >= c_first + c_count) {
    idx = -1;
  } else {
    const size_t clen = (size_t)1 << chunk_bits;
    const E *sc = state + ((size_t)b << n) + ((size_t)(lo - c_first) << chunk_bits);
    const double pre = p[lo];
    double run = 0.0;
    size_t cnt = 0;
    size_t clen_eff = clen;
    if (sub) {
      // enter the chunk at the first 128-amplitude block whose LAST running sum fails the test (the test is monotone
      // along the chunk): the scan then covers at most 128 amplitudes instead of up to 4096
      const double *sr = sub + ((size_t)b * (size_t)c_count + (size_t)(lo - c_first)) * 32;
      int blo = 0, bhi = 32;
      while (blo < bhi) {
        const int mid = (blo + bhi) >> 1;
        if (__ddiv_rn(__dadd_rn(pre, sr[mid]), total) <= u) blo = mid + 1; else bhi = mid;
      }
      if (blo >= 32) {
        cnt = clen;      // every amplitude of the chunk passes (rounding at the chunk end): same result as the full scan
        clen_eff = 0;
      } else {
        cnt = (size_t)blo << 7;
        run = blo ? sr[blo - 1] : 0.0;
        clen_eff = cnt + 128;
      }
    }
    // The test is fl(cdf / total) <= u with the division correctly rounded.  A cdf safely below / above u * total
    // decides it without dividing (relative margin 2^-49 >> the 2^-53 roundings of the product and the quotient);
    // only values inside the margin take the exact division -- same indices, bit for bit, without ~4096 DDIVs a shot.
    const double ut = __dmul_rn(u, total);
    const double safe_lo = __dmul_rn(ut, 1.0 - 0x1p-49), safe_hi = __dmul_rn(ut, 1.0 + 0x1p-49);
    for (; cnt < clen_eff; ++cnt) {
      run = __dadd_rn(run, prob_of(sc[cnt]));
      const double v = __dadd_rn(pre, run);
      if (v < safe_lo) continue;
      if (v > safe_hi) break;
      if (!(__ddiv_rn(v, total) <= u)) break;
    }
    idx = (lo << chunk_bits) + (long long)cnt;
  }
  idx_out[(size_t)b * shots + s] = idx;
}

// ---- Pauli-Z products from sampled indices (postprocessing/counts_expval.py:7-20, 87-112) -------------------
// Member b (one basis-rotated copy of a state, measured `shots` times) evaluates the terms of group b % n_groups:
// <Z_S> = (shots - 2 * #{s : popc(idx_s & zmask) odd}) / shots -- integer counting, so the value is exactly what
// the reference computes from the counts dict; energy[b] = sum_t coef_t * <Z_S_t> in term order.
__global__ void __launch_bounds__(RT) expval_samples_kernel(const long long *__restrict__ idx, long long shots, int n_groups,
                                                            const int *__restrict__ term_ptr, const uint64_t *__restrict__ term_z,
                                                            const double *__restrict__ term_coef, double *energy,
                                                            double *expvals, long long expvals_stride) {
  __shared__ long long sm[RT / 32];
  __shared__ double acc;
  const long long b = blockIdx.x;
  const int g = (int)(b % n_groups);
  const long long *ib = idx + b * shots;
  if (threadIdx.x == 0) acc = 0.0;
  __syncthreads();
  for (int t = term_ptr[g]; t < term_ptr[g + 1]; ++t) {
    const uint64_t z = term_z[t];
    long long odd = 0;
    for (long long s = threadIdx.x; s < shots; s += RT) odd += __popcll((uint64_t)ib[s] & z) & 1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) odd += __shfl_down_sync(0xffffffffu, odd, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = odd;
    __syncthreads();
    if (threadIdx.x == 0) {
      long long tot = 0;
      for (int w = 0; w < RT / 32; ++w) tot += sm[w];
      const double ev = __ddiv_rn((double)(shots - 2 * tot), (double)shots);
      if (expvals) expvals[b * expvals_stride + (t - term_ptr[g])] = ev;
      acc = __dadd_rn(acc, __dmul_rn(term_coef[t], ev));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) energy[b] = acc;
}

// ---- host helpers ----------------------------------------------------------------------------
static int pick_seg(const Workspace &ws, int n, int64_t batch, int nv, int max_seg_bits, Seg *sg, int *nbx) {
  // power-of-two number of segments per batch member, ~8 CTAs per SM overall
  long long want = ((long long)ws.sm_count * 8 + batch - 1) / batch;
  int lb = 0;
  while ((1ll << lb) < want) ++lb;
  int min_seg = n < 10 ? n : 10;  // at least 1024 amplitudes per CTA (or the whole state)
  if (lb > n - min_seg) lb = n - min_seg;
  if (lb < 0) lb = 0;
  if (n - lb > max_seg_bits) lb = n - max_seg_bits;
  if ((size_t)batch * ((size_t)1 << lb) * nv * sizeof(double) > ws.bytes / 2) return fail("reduction workspace too small");
  if (batch > 65535) return fail("batch > 65535 not supported by the reduction kernels");
  sg->n = n;
  sg->seg_bits = n - lb;
  *nbx = 1 << lb;
  return 0;
}

template <typename F64, typename F32>
static int by_dtype(int dtype, F64 f64, F32 f32) {
  if (dtype == TQB_C128) { f64(); return 0; }
  if (dtype == TQB_C64) { f32(); return 0; }
  return fail("bad dtype");
}

}  // namespace tqb

using namespace tqb;

#define CD(p) reinterpret_cast<const cplx<double> *>(p)
#define CF(p) reinterpret_cast<const cplx<float> *>(p)
#define MD(p) reinterpret_cast<cplx<double> *>(p)
#define MF(p) reinterpret_cast<cplx<float> *>(p)

extern "C" {

int tqb_norm2(const void *state, int n, int64_t batch, int dtype, double *out_dev, void *stream) {
  TQB_REQUIRE(state && out_dev && n >= 0 && n < 48 && batch >= 1, "tqb_norm2: bad arguments");
  Workspace *ws = workspace();
  if (!ws) return -1;
  Seg sg; int nbx;
  if (pick_seg(*ws, n, batch, 1, 40, &sg, &nbx)) return -1;
  double *partial = (double *)ws->ptr;
  dim3 grid(nbx, (unsigned)batch);
  cudaStream_t st = as_stream(stream);
  if (dtype == TQB_F64) {
    norm2_kernel<double><<<grid, RT, 0, st>>>((const double *)state, sg, partial);
  } else
  if (by_dtype(dtype, [&] { norm2_kernel<cplx<double>><<<grid, RT, 0, st>>>(CD(state), sg, partial); },
               [&] { norm2_kernel<cplx<float>><<<grid, RT, 0, st>>>(CF(state), sg, partial); })) return -1;
  TQB_CHECK_LAUNCH("norm2_kernel");
  finish_kernel<<<(unsigned)batch, 32, 0, st>>>(partial, nbx, 1, 1, out_dev, 1);
  TQB_CHECK_LAUNCH("finish_kernel");
  return 0;
}

int tqb_expect_z_bits(const void *state, int n, int64_t batch, int dtype, double *out_dev, void *stream) {
  TQB_REQUIRE(state && out_dev && n >= 1 && n <= MAX_ACC && batch >= 1, "tqb_expect_z_bits: bad arguments (1 <= n <= 40)");
  Workspace *ws = workspace();
  if (!ws) return -1;
  Seg sg; int nbx;
  if (pick_seg(*ws, n, batch, MAX_ACC, 24, &sg, &nbx)) return -1;
  double *partial = (double *)ws->ptr;
  dim3 grid(nbx, (unsigned)batch);
  cudaStream_t st = as_stream(stream);
  if (dtype == TQB_F64) {
    expect_z_bits_kernel<double><<<grid, RT, 0, st>>>((const double *)state, sg, partial);
  } else
  if (by_dtype(dtype, [&] { expect_z_bits_kernel<cplx<double>><<<grid, RT, 0, st>>>(CD(state), sg, partial); },
               [&] { expect_z_bits_kernel<cplx<float>><<<grid, RT, 0, st>>>(CF(state), sg, partial); })) return -1;
  TQB_CHECK_LAUNCH("expect_z_bits_kernel");
  finish_kernel<<<(unsigned)batch, 64, 0, st>>>(partial, nbx, MAX_ACC, n, out_dev, n);
  TQB_CHECK_LAUNCH("finish_kernel");
  return 0;
}

int tqb_expect_zmasks(const void *state, int n, int64_t batch, int dtype, uint64_t global_base, const uint64_t *masks_dev,
                      int n_masks, double *out_dev, void *stream) {
  TQB_REQUIRE(state && out_dev && masks_dev && n >= 0 && n < 48 && batch >= 1 && n_masks >= 1, "tqb_expect_zmasks: bad arguments");
  Workspace *ws = workspace();
  if (!ws) return -1;
  Seg sg; int nbx;
  if (pick_seg(*ws, n, batch, MASKS_PER_LAUNCH, 40, &sg, &nbx)) return -1;
  double *partial = (double *)ws->ptr;
  dim3 grid(nbx, (unsigned)batch);
  cudaStream_t st = as_stream(stream);
  for (int m0 = 0; m0 < n_masks; m0 += MASKS_PER_LAUNCH) {
    const int cnt = n_masks - m0 < MASKS_PER_LAUNCH ? n_masks - m0 : MASKS_PER_LAUNCH;
    if (by_dtype(dtype,
                 [&] { expect_zmasks_kernel<double><<<grid, RT, 0, st>>>(CD(state), sg, global_base, masks_dev + m0, cnt, partial); },
                 [&] { expect_zmasks_kernel<float><<<grid, RT, 0, st>>>(CF(state), sg, global_base, masks_dev + m0, cnt, partial); }))
      return -1;
    TQB_CHECK_LAUNCH("expect_zmasks_kernel");
    finish_kernel<<<(unsigned)batch, 32, 0, st>>>(partial, nbx, MASKS_PER_LAUNCH, cnt, out_dev + m0, n_masks);
    TQB_CHECK_LAUNCH("finish_kernel");
  }
  return 0;
}

int tqb_expect_pauli_sum(const void *state, int n, int64_t batch, int dtype, uint64_t global_base, const uint64_t *group_x,
                         const int32_t *group_ptr, int n_groups, const uint64_t *term_z, const double *term_coef,
                         double *out_dev, void *stream) {
  TQB_REQUIRE(state && out_dev && group_x && group_ptr && term_z && term_coef && n >= 0 && n < 48 && batch >= 1 && n_groups >= 0,
              "tqb_expect_pauli_sum: bad arguments");
  Workspace *ws = workspace();
  if (!ws) return -1;
  Seg sg; int nbx;
  if (pick_seg(*ws, n, batch, 2, 40, &sg, &nbx)) return -1;
  double *partial = (double *)ws->ptr;
  dim3 grid(nbx, (unsigned)batch);
  cudaStream_t st = as_stream(stream);
  if (by_dtype(dtype,
               [&] { expect_pauli_kernel<double><<<grid, RT, 0, st>>>(CD(state), sg, global_base, group_x, group_ptr, n_groups, term_z, term_coef, partial); },
               [&] { expect_pauli_kernel<float><<<grid, RT, 0, st>>>(CF(state), sg, global_base, group_x, group_ptr, n_groups, term_z, term_coef, partial); }))
    return -1;
  TQB_CHECK_LAUNCH("expect_pauli_kernel");
  finish_kernel<<<(unsigned)batch, 32, 0, st>>>(partial, nbx, 2, 2, out_dev, 2);
  TQB_CHECK_LAUNCH("finish_kernel");
  return 0;
}

int tqb_expect_pauli_tiled(const void *state, int n, int64_t batch, int dtype, uint64_t global_base,
                           const tqb_pauli_layout *layouts_host, int n_layouts, const uint32_t *group_xl_dev,
                           const int32_t *group_ptr_dev, const uint32_t *term_zl_dev, const uint64_t *term_zout_dev,
                           const double *term_coef_dev, int flags, double *out_dev, void *stream) {
  TQB_REQUIRE(state && out_dev && layouts_host && n_layouts >= 1 && n_layouts <= 64 && group_xl_dev && group_ptr_dev && term_zl_dev &&
                  term_zout_dev && term_coef_dev && n >= 1 && n < 48 && batch >= 1 && batch <= 65535,
              "tqb_expect_pauli_tiled: bad arguments");
  Workspace *ws = workspace();
  if (!ws) return -1;
  const int es = dtype == TQB_C128 ? 16 : 8;
  if (dtype != TQB_C128 && dtype != TQB_C64) return fail("bad dtype");
  cudaStream_t st = as_stream(stream);
  // CTAs per batch member: about 4 per SM overall, a power of two, at most one per tile
  int nbx = 1;
  double *partial = (double *)ws->ptr;
  for (int li = 0; li < n_layouts; ++li) {
    const tqb_pauli_layout &lay = layouts_host[li];
    TQB_REQUIRE(lay.m >= 1 && lay.m <= n && lay.m <= 14 && lay.L >= 0 && lay.L <= lay.m && lay.m - lay.L <= 16 && lay.n_groups >= 1 &&
                    lay.n_terms >= 1, "tqb_expect_pauli_tiled: bad layout");
    if (li == 0) {
      const long long tiles = 1ll << (n - lay.m);
      long long want = ((long long)ws->sm_count * 4 + batch - 1) / batch;
      while (nbx < want && nbx < tiles) nbx <<= 1;
      TQB_REQUIRE((size_t)batch * n_layouts * nbx * 2 * sizeof(double) <= ws->bytes, "tqb_expect_pauli_tiled: workspace too small");
    }
    PauliTile pt;
    pt.n = n; pt.m = lay.m; pt.L = lay.L; pt.h = lay.m - lay.L; pt.g0 = lay.group_begin; pt.ng = lay.n_groups;
    pt.layout = li; pt.n_layouts = n_layouts;
    int prev = lay.L - 1;
    for (int i = 0; i < 16; ++i) {
      pt.hb[i] = i < pt.h ? lay.hb[i] : 0;
      if (i < pt.h) {
        TQB_REQUIRE(lay.hb[i] > prev && lay.hb[i] < n, "tqb_expect_pauli_tiled: hb must be ascending, >= L and < n");
        prev = lay.hb[i];
      }
    }
    const size_t smem = ((size_t)es << lay.m) + 8 * ((((size_t)1 << pt.h) + 1) & ~(size_t)1) + (size_t)lay.n_terms * 24 + (size_t)lay.n_groups * 8 +
                        256 * (size_t)es + 64;   // tile | run offsets | sc, szl, sst per term | sgx, sgp per group | sign-sum tables
    TQB_REQUIRE(smem + 1024 <= (size_t)ws->max_smem_optin, "tqb_expect_pauli_tiled: tile + terms exceed shared memory");
#define TQB_PT(T, HERM, REALC, PTR)                                                                                              \
  do {                                                                                                                            \
    auto kern = expect_pauli_tiled_kernel<T, HERM, REALC>;                                                                        \
    TQB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ws->max_smem_optin - 1024));           \
    kern<<<dim3((unsigned)nbx, (unsigned)batch), RT, smem, st>>>(PTR(state), pt, global_base, group_xl_dev, group_ptr_dev,       \
                                                                   term_zl_dev, term_zout_dev, term_coef_dev, partial);         \
  } while (0)
    const int f = flags & 3;
    if (dtype == TQB_C128) {
      if (f == 3) TQB_PT(double, true, true, CD); else if (f == 1) TQB_PT(double, true, false, CD);
      else if (f == 2) TQB_PT(double, false, true, CD); else TQB_PT(double, false, false, CD);
    } else {
      if (f == 3) TQB_PT(float, true, true, CF); else if (f == 1) TQB_PT(float, true, false, CF);
      else if (f == 2) TQB_PT(float, false, true, CF); else TQB_PT(float, false, false, CF);
    }
#undef TQB_PT
    TQB_CHECK_LAUNCH("expect_pauli_tiled_kernel");
  }
  finish_kernel<<<(unsigned)batch, 32, 0, st>>>(partial, nbx * n_layouts, 2, 2, out_dev, 2);
  TQB_CHECK_LAUNCH("finish_kernel");
  return 0;
}

int tqb_apply_pauli_sum(const void *state, void *out, int n, int64_t batch, int dtype, uint64_t global_base,
                        const uint64_t *group_x, const int32_t *group_ptr, int n_groups, const uint64_t *term_z,
                        const double *term_coef, void *stream) {
  TQB_REQUIRE(state && out && state != out && group_x && group_ptr && term_z && term_coef && n >= 0 && n < 48 && batch >= 1,
              "tqb_apply_pauli_sum: bad arguments (out must not alias state)");
  TQB_REQUIRE(batch <= 65535, "tqb_apply_pauli_sum: batch > 65535");
  Workspace *ws = workspace();
  if (!ws) return -1;
  const size_t dim = (size_t)1 << n;
  size_t bx = (dim + RT - 1) / RT;
  const size_t cap = (size_t)ws->sm_count * 16;
  if (bx > cap) bx = cap;
  cudaStream_t st = as_stream(stream);
  // small states: slices of the groups on blockIdx.z until ~4 CTAs per SM are in flight
  int slices = 1, gps = 0;
  const size_t ctas = bx * (size_t)batch, want = (size_t)ws->sm_count * 4;
  if (ctas < want && n_groups >= 8) {
    slices = (int)((want + ctas - 1) / ctas);
    if (slices > 32) slices = 32;
    if (slices > n_groups / 4) slices = n_groups / 4;
    if (slices > 1) {
      gps = (n_groups + slices - 1) / slices;
      slices = (n_groups + gps - 1) / gps;
      TQB_CHECK_CUDA(cudaMemsetAsync(out, 0, ((size_t)batch << n) * (dtype == TQB_C128 ? 16 : 8), st));
    }
  }
  dim3 grid((unsigned)bx, (unsigned)batch, (unsigned)(slices > 1 ? slices : 1));
  if (slices <= 1) gps = 0;
  if (by_dtype(dtype,
               [&] { apply_pauli_kernel<double><<<grid, RT, 0, st>>>(CD(state), MD(out), n, global_base, group_x, group_ptr, n_groups, term_z, term_coef, gps); },
               [&] { apply_pauli_kernel<float><<<grid, RT, 0, st>>>(CF(state), MF(out), n, global_base, group_x, group_ptr, n_groups, term_z, term_coef, gps); }))
    return -1;
  TQB_CHECK_LAUNCH("apply_pauli_kernel");
  return 0;
}

int tqb_transition_1q(const void *bra, const void *ket, int n, int dtype, int m, int L, const int8_t *hb, const int8_t *tile_bits,
                      int n_bits, double *out_dev, void *stream) {
  TQB_REQUIRE(bra && ket && hb && tile_bits && out_dev && n >= 1 && n < 48 && m >= 1 && m <= n && m <= 12 && L >= 0 && L <= m && m - L <= 16 &&
                  n_bits >= 1 && n_bits <= 6, "tqb_transition_1q: bad arguments (m <= 12, at most 6 bits per call)");
  TQB_REQUIRE(dtype == TQB_C128 || dtype == TQB_C64, "tqb_transition_1q: bad dtype");
  Workspace *ws = workspace();
  if (!ws) return -1;
  Trans1q tq;
  tq.n = n; tq.m = m; tq.L = L; tq.h = m - L; tq.nb = n_bits;
  int prev = L - 1;
  for (int i = 0; i < 16; ++i) {
    tq.hb[i] = i < tq.h ? hb[i] : 0;
    if (i < tq.h) {
      TQB_REQUIRE(hb[i] > prev && hb[i] < n, "tqb_transition_1q: hb must be ascending, >= L and < n");
      prev = hb[i];
    }
  }
  for (int i = 0; i < 8; ++i) {
    tq.tb[i] = i < n_bits ? tile_bits[i] : 0;
    if (i < n_bits) TQB_REQUIRE(tile_bits[i] >= 0 && tile_bits[i] < m, "tqb_transition_1q: bit of interest outside the tile");
  }
  const int es = dtype == TQB_C128 ? 16 : 8;
  const size_t smem = 2 * ((size_t)es << m) + ((size_t)8 << tq.h) + 16;
  TQB_REQUIRE(smem + 4096 <= (size_t)ws->max_smem_optin, "tqb_transition_1q: tiles exceed shared memory");
  cudaStream_t st = as_stream(stream);
  TQB_CHECK_CUDA(cudaMemsetAsync(out_dev, 0, (size_t)n_bits * 8 * sizeof(double), st));
  const unsigned long long tiles = 1ull << (n - m);
  unsigned long long grid = (unsigned long long)ws->sm_count * 2;
  if (grid > tiles) grid = tiles;
#define TQB_TR(T, PTR)                                                                                                              \
  do {                                                                                                                              \
    auto kern = transition_1q_kernel<T, 6>;                                                                                         \
    TQB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ws->max_smem_optin - 4096));             \
    kern<<<(unsigned)grid, RT, smem, st>>>(PTR(bra), PTR(ket), tq, out_dev);                                                        \
  } while (0)
  if (dtype == TQB_C128) TQB_TR(double, CD); else TQB_TR(float, CF);
#undef TQB_TR
  TQB_CHECK_LAUNCH("transition_1q_kernel");
  return 0;
}

int tqb_inner(const void *a, const void *b, int n, int64_t batch, int dtype, double *out_dev, void *stream) {
  TQB_REQUIRE(a && b && out_dev && n >= 0 && n < 48 && batch >= 1, "tqb_inner: bad arguments");
  Workspace *ws = workspace();
  if (!ws) return -1;
  Seg sg; int nbx;
  if (pick_seg(*ws, n, batch, 2, 40, &sg, &nbx)) return -1;
  double *partial = (double *)ws->ptr;
  dim3 grid(nbx, (unsigned)batch);
  cudaStream_t st = as_stream(stream);
  if (by_dtype(dtype, [&] { inner_kernel<double><<<grid, RT, 0, st>>>(CD(a), CD(b), sg, partial); },
               [&] { inner_kernel<float><<<grid, RT, 0, st>>>(CF(a), CF(b), sg, partial); })) return -1;
  TQB_CHECK_LAUNCH("inner_kernel");
  finish_kernel<<<(unsigned)batch, 32, 0, st>>>(partial, nbx, 2, 2, out_dev, 2);
  TQB_CHECK_LAUNCH("finish_kernel");
  return 0;
}

int tqb_reduced_1q(const void *state, int n, int64_t batch, int dtype, int bit, double *out_dev, void *stream) {
  TQB_REQUIRE(state && out_dev && n >= 1 && n < 48 && batch >= 1 && bit >= 0 && bit < n, "tqb_reduced_1q: bad arguments");
  Workspace *ws = workspace();
  if (!ws) return -1;
  Seg sg; int nbx;
  if (pick_seg(*ws, n - 1, batch, 4, 40, &sg, &nbx)) return -1;
  double *partial = (double *)ws->ptr;
  dim3 grid(nbx, (unsigned)batch);
  cudaStream_t st = as_stream(stream);
  if (by_dtype(dtype, [&] { reduced_1q_kernel<double><<<grid, RT, 0, st>>>(CD(state), n, bit, sg, partial); },
               [&] { reduced_1q_kernel<float><<<grid, RT, 0, st>>>(CF(state), n, bit, sg, partial); })) return -1;
  TQB_CHECK_LAUNCH("reduced_1q_kernel");
  finish_kernel<<<(unsigned)batch, 32, 0, st>>>(partial, nbx, 4, 4, out_dev, 4);
  TQB_CHECK_LAUNCH("finish_kernel");
  return 0;
}

int tqb_grad_pair(const void *bra, const void *ket, int n, int dtype, const tqb_gate *gate_host, double scale,
                  double *out_dev, int slot, void *stream) {
  TQB_REQUIRE(bra && ket && gate_host && out_dev && n >= 1 && n < 48 && slot >= 0, "tqb_grad_pair: bad arguments");
  TQB_REQUIRE(gate_host->kind == TQB_GATE_PAIR && gate_host->k >= 1 && gate_host->k <= TQB_MAX_GATE_BITS && gate_host->k <= n,
              "tqb_grad_pair: gate must be a PAIR gate with m = n");
  Workspace *ws = workspace();
  if (!ws) return -1;
  PairGen pg;
  pg.n = n; pg.k = gate_host->k; pg.off_a = gate_host->off_a; pg.off_b = gate_host->off_b; pg.zmask = gate_host->zmask;
  for (int i = 0; i < TQB_MAX_GATE_BITS; ++i) pg.sbits[i] = gate_host->sbits[i];
  const uint64_t ngroups = 1ull << (n - pg.k);
  uint64_t bx = (ngroups + RT - 1) / RT;
  const uint64_t cap = (uint64_t)ws->sm_count * 8;
  if (bx > cap) bx = cap;
  cudaStream_t st = as_stream(stream);
  if (by_dtype(dtype, [&] { grad_pair_kernel<double><<<(unsigned)bx, RT, 0, st>>>(CD(bra), CD(ket), pg, scale, out_dev + slot); },
               [&] { grad_pair_kernel<float><<<(unsigned)bx, RT, 0, st>>>(CF(bra), CF(ket), pg, scale, out_dev + slot); })) return -1;
  TQB_CHECK_LAUNCH("grad_pair_kernel");
  return 0;
}

int tqb_grad_dense(const void *bra, const void *ket, int n, int dtype, int k, const int *bits, const double *gen_host,
                   double scale, double *out_dev, int slot, void *stream) {
  TQB_REQUIRE(bra && ket && bits && gen_host && out_dev && n >= 1 && n < 48 && k >= 1 && k <= 2 && k <= n && slot >= 0,
              "tqb_grad_dense: bad arguments (k <= 2)");
  Workspace *ws = workspace();
  if (!ws) return -1;
  DenseGen dg;
  dg.n = n; dg.k = k;
  for (int j = 0; j < 2; ++j) dg.bits[j] = dg.sbits[j] = 0;
  for (int j = 0; j < k; ++j) {
    TQB_REQUIRE(bits[j] >= 0 && bits[j] < n, "tqb_grad_dense: bit out of range");
    dg.bits[j] = (int8_t)bits[j];
    dg.sbits[j] = (int8_t)bits[j];
  }
  if (k == 2) {
    TQB_REQUIRE(bits[0] != bits[1], "tqb_grad_dense: repeated bit");
    if (dg.sbits[0] > dg.sbits[1]) { int8_t t = dg.sbits[0]; dg.sbits[0] = dg.sbits[1]; dg.sbits[1] = t; }
  }
  const int D = 1 << k;
  for (int i = 0; i < 32; ++i) dg.gen[i] = i < 2 * D * D ? gen_host[i] : 0.0;
  const uint64_t ngroups = 1ull << (n - k);
  uint64_t bx = (ngroups + RT - 1) / RT;
  const uint64_t cap = (uint64_t)ws->sm_count * 8;
  if (bx > cap) bx = cap;
  cudaStream_t st = as_stream(stream);
  if (by_dtype(dtype, [&] { grad_dense_kernel<double><<<(unsigned)bx, RT, 0, st>>>(CD(bra), CD(ket), dg, scale, out_dev + slot); },
               [&] { grad_dense_kernel<float><<<(unsigned)bx, RT, 0, st>>>(CF(bra), CF(ket), dg, scale, out_dev + slot); })) return -1;
  TQB_CHECK_LAUNCH("grad_dense_kernel");
  return 0;
}

int tqb_pair_sweep(void *ket, void *bra, int n, int dtype, const tqb_pair_step *steps_dev, int n_steps, int mode,
                   double *out_dev, unsigned long long *sync_dev, void *stream) {
  TQB_REQUIRE(ket && steps_dev && sync_dev && n >= 1 && n <= 26 && n_steps >= 1 && (mode == 0 || (mode == 1 && bra && out_dev)),
              "tqb_pair_sweep: bad arguments (n <= 26: the state must stay L2-resident)");
  Workspace *ws = workspace();
  if (!ws) return -1;
  cudaStream_t st = as_stream(stream);
  if (n <= 15) {   // one CTA: barriers are __syncthreads()
    if (by_dtype(dtype,
                 [&] { pair_sweep_cta_kernel<double><<<1, SWEEP_CTA, 0, st>>>(MD(ket), MD(bra), n, steps_dev, n_steps, mode, out_dev); },
                 [&] { pair_sweep_cta_kernel<float><<<1, SWEEP_CTA, 0, st>>>(MF(ket), MF(bra), n, steps_dev, n_steps, mode, out_dev); }))
      return -1;
    TQB_CHECK_LAUNCH("pair_sweep_cta_kernel");
    return 0;
  }
  TQB_CHECK_CUDA(cudaMemsetAsync(sync_dev, 0, sizeof(unsigned long long), st));
  // all CTAs must be co-resident for the grid barrier: one CTA per SM at most
  uint64_t want = ((1ull << (n > 2 ? n - 2 : 0)) + RT - 1) / RT;
  int grid = (int)(want < (uint64_t)ws->sm_count ? want : (uint64_t)ws->sm_count);
  if (grid < 1) grid = 1;
  if (by_dtype(dtype,
               [&] { pair_sweep_kernel<double><<<grid, RT, 0, st>>>(MD(ket), MD(bra), n, steps_dev, n_steps, mode, out_dev, sync_dev); },
               [&] { pair_sweep_kernel<float><<<grid, RT, 0, st>>>(MF(ket), MF(bra), n, steps_dev, n_steps, mode, out_dev, sync_dev); }))
    return -1;
  TQB_CHECK_LAUNCH("pair_sweep_kernel");
  return 0;
}

int tqb_project_z(void *state, int n, int64_t batch, int dtype, int bit, int keep, void *stream) {
  TQB_REQUIRE(state && n >= 1 && n < 48 && batch >= 1 && bit >= 0 && bit < n, "tqb_project_z: bad arguments");
  Workspace *ws = workspace();
  if (!ws) return -1;
  Seg sg; int nbx;
  if (pick_seg(*ws, n, batch, 1, 40, &sg, &nbx)) return -1;
  double *partial = (double *)ws->ptr;
  double *norms = (double *)((char *)ws->ptr + ws->bytes / 2);  // batch doubles
  dim3 grid(nbx, (unsigned)batch);
  cudaStream_t st = as_stream(stream);
  const int kp = keep ? 1 : 0;
  if (by_dtype(dtype, [&] { project_kernel<double><<<grid, RT, 0, st>>>(MD(state), sg, bit, kp, partial); },
               [&] { project_kernel<float><<<grid, RT, 0, st>>>(MF(state), sg, bit, kp, partial); })) return -1;
  TQB_CHECK_LAUNCH("project_kernel");
  finish_kernel<<<(unsigned)batch, 32, 0, st>>>(partial, nbx, 1, 1, norms, 1);
  TQB_CHECK_LAUNCH("finish_kernel");
  const size_t dim = (size_t)1 << n;
  size_t bx = (dim + RT - 1) / RT;
  const size_t cap = (size_t)ws->sm_count * 16;
  if (bx > cap) bx = cap;
  dim3 g2((unsigned)bx, (unsigned)batch);
  if (by_dtype(dtype, [&] { renorm_kernel<double><<<g2, RT, 0, st>>>(MD(state), n, norms); },
               [&] { renorm_kernel<float><<<g2, RT, 0, st>>>(MF(state), n, norms); })) return -1;
  TQB_CHECK_LAUNCH("renorm_kernel");
  return 0;
}

int tqb_scale(void *state, int n, int64_t batch, int dtype, double factor, void *stream) {
  TQB_REQUIRE(state && n >= 0 && n < 48 && batch >= 1, "tqb_scale: bad arguments");
  Workspace *ws = workspace();
  if (!ws) return -1;
  const size_t total = (size_t)batch << n;
  size_t bx = (total + RT - 1) / RT;
  const size_t cap = (size_t)ws->sm_count * 16;
  if (bx > cap) bx = cap;
  cudaStream_t st = as_stream(stream);
  if (by_dtype(dtype, [&] { scale_kernel<double><<<(unsigned)bx, RT, 0, st>>>(MD(state), total, factor); },
               [&] { scale_kernel<float><<<(unsigned)bx, RT, 0, st>>>(MF(state), total, factor); })) return -1;
  TQB_CHECK_LAUNCH("scale_kernel");
  return 0;
}

int tqb_probabilities(const void *state, int n, int64_t batch, int dtype, double *out_dev, void *stream) {
  TQB_REQUIRE(state && out_dev && n >= 0 && n < 48 && batch >= 1, "tqb_probabilities: bad arguments");
  Workspace *ws = workspace();
  if (!ws) return -1;
  const size_t total = (size_t)batch << n;
  size_t bx = (total + RT - 1) / RT;
  const size_t cap = (size_t)ws->sm_count * 16;
  if (bx > cap) bx = cap;
  cudaStream_t st = as_stream(stream);
  if (by_dtype(dtype, [&] { probs_kernel<double><<<(unsigned)bx, RT, 0, st>>>(CD(state), total, out_dev); },
               [&] { probs_kernel<float><<<(unsigned)bx, RT, 0, st>>>(CF(state), total, out_dev); })) return -1;
  TQB_CHECK_LAUNCH("probs_kernel");
  return 0;
}

static int launch_chunk_totals(const void *state, int n, int64_t batch, int dtype, double *totals, cudaStream_t st, double *sub = nullptr) {
  const int chunk_bits = n < 12 ? n : 12;  // TQB_SCAN_BLOCK = 4096
  const long long nct = (1ll << (n - chunk_bits)) * batch;
  const long long groups = (nct + 31) / 32;
  const long long blocks = (groups + CW - 1) / CW;
  if (dtype == TQB_F64) {
    if (chunk_bits != 12) sub = nullptr;
    chunk_totals_kernel<double><<<(unsigned)blocks, CW * 32, 0, st>>>((const double *)state, n, chunk_bits, nct, totals, sub);
  } else
  if (by_dtype(dtype, [&] { chunk_totals_kernel<cplx<double>><<<(unsigned)blocks, CW * 32, 0, st>>>(CD(state), n, chunk_bits, nct, totals, chunk_bits == 12 ? sub : nullptr); },
               [&] { chunk_totals_kernel<cplx<float>><<<(unsigned)blocks, CW * 32, 0, st>>>(CF(state), n, chunk_bits, nct, totals, chunk_bits == 12 ? sub : nullptr); })) return -1;
  TQB_CHECK_LAUNCH("chunk_totals_kernel");
  return 0;
}

int tqb_dm_diag(const void *rho, int n, int64_t batch, int dtype, double *out_dev, void *stream) {
  TQB_REQUIRE(rho && out_dev && n >= 0 && n <= 20 && batch >= 1, "tqb_dm_diag: bad arguments (n <= 20 qubits)");
  Workspace *ws = workspace();
  if (!ws) return -1;
  const size_t total = (size_t)batch << n;
  size_t bx = (total + RT - 1) / RT;
  const size_t cap = (size_t)ws->sm_count * 16;
  if (bx > cap) bx = cap;
  cudaStream_t st = as_stream(stream);
  if (by_dtype(dtype, [&] { dm_diag_kernel<double><<<(unsigned)bx, RT, 0, st>>>(CD(rho), n, total, out_dev); },
               [&] { dm_diag_kernel<float><<<(unsigned)bx, RT, 0, st>>>(CF(rho), n, total, out_dev); })) return -1;
  TQB_CHECK_LAUNCH("dm_diag_kernel");
  return 0;
}

int tqb_cdf_chunks(const void *state, int n, int64_t batch, int dtype, double *chunk_prefix_dev, void *stream) {
  return tqb_cdf_chunks2(state, n, batch, dtype, chunk_prefix_dev, nullptr, stream);
}

int tqb_cdf_chunks2(const void *state, int n, int64_t batch, int dtype, double *chunk_prefix_dev, double *sub_prefix_dev,
                    void *stream) {
  TQB_REQUIRE(state && chunk_prefix_dev && n >= 0 && n < 48 && batch >= 1, "tqb_cdf_chunks: bad arguments");
  Workspace *ws = workspace();
  if (!ws) return -1;
  const int chunk_bits = n < 12 ? n : 12;
  const long long nc = 1ll << (n - chunk_bits);
  TQB_REQUIRE((size_t)(nc * batch) * sizeof(double) <= ws->bytes, "tqb_cdf_chunks: workspace too small for the chunk totals");
  double *totals = (double *)ws->ptr;
  cudaStream_t st = as_stream(stream);
  if (launch_chunk_totals(state, n, batch, dtype, totals, st, sub_prefix_dev)) return -1;
  chunk_prefix_kernel<<<(unsigned)batch, RT, 0, st>>>(totals, nc, chunk_prefix_dev);
  TQB_CHECK_LAUNCH("chunk_prefix_kernel");
  return 0;
}

int tqb_chunk_totals(const void *state, int n, int64_t batch, int dtype, double *totals_dev, void *stream) {
  TQB_REQUIRE(state && totals_dev && n >= 0 && n < 48 && batch >= 1, "tqb_chunk_totals: bad arguments");
  return launch_chunk_totals(state, n, batch, dtype, totals_dev, as_stream(stream));
}

int tqb_chunk_prefix(const double *totals_dev, int64_t n_chunks, int64_t batch, double *chunk_prefix_dev, void *stream) {
  TQB_REQUIRE(totals_dev && chunk_prefix_dev && n_chunks >= 1 && batch >= 1 && batch <= 65535, "tqb_chunk_prefix: bad arguments");
  chunk_prefix_kernel<<<(unsigned)batch, RT, 0, as_stream(stream)>>>(totals_dev, (long long)n_chunks, chunk_prefix_dev);
  TQB_CHECK_LAUNCH("chunk_prefix_kernel");
  return 0;
}

int tqb_sample(const void *state, int n, int64_t batch, int dtype, const double *chunk_prefix_dev, const double *uniforms_dev,
               int64_t shots, int64_t *idx_dev, void *stream) {
  return tqb_sample2(state, n, batch, dtype, chunk_prefix_dev, nullptr, uniforms_dev, shots, idx_dev, stream);
}

int tqb_sample2(const void *state, int n, int64_t batch, int dtype, const double *chunk_prefix_dev, const double *sub_prefix_dev,
                const double *uniforms_dev, int64_t shots, int64_t *idx_dev, void *stream) {
  TQB_REQUIRE(state && chunk_prefix_dev && uniforms_dev && idx_dev && n >= 0 && n < 48 && batch >= 1 && batch <= 65535 && shots >= 1,
              "tqb_sample: bad arguments");
  const int chunk_bits = n < 12 ? n : 12;
  const long long nc = 1ll << (n - chunk_bits);
  dim3 grid((unsigned)((shots + 127) / 128), (unsigned)batch);
  cudaStream_t st = as_stream(stream);
  const double *sub = chunk_bits == 12 ? sub_prefix_dev : nullptr;
  if (dtype == TQB_F64) {
    sample_kernel<double><<<grid, 128, 0, st>>>((const double *)state, n, chunk_bits, nc, 0, nc, 1ll << n, chunk_prefix_dev, uniforms_dev, (long long)shots, (long long *)idx_dev, sub);
  } else
  if (by_dtype(dtype,
               [&] { sample_kernel<cplx<double>><<<grid, 128, 0, st>>>(CD(state), n, chunk_bits, nc, 0, nc, 1ll << n, chunk_prefix_dev, uniforms_dev, (long long)shots, (long long *)idx_dev, sub); },
               [&] { sample_kernel<cplx<float>><<<grid, 128, 0, st>>>(CF(state), n, chunk_bits, nc, 0, nc, 1ll << n, chunk_prefix_dev, uniforms_dev, (long long)shots, (long long *)idx_dev, sub); }))
    return -1;
  TQB_CHECK_LAUNCH("sample_kernel");
  return 0;
}

int tqb_expval_from_samples(const int64_t *idx_dev, int64_t batch, int64_t shots, int n_groups, const int32_t *term_ptr_dev,
                            const uint64_t *term_z_dev, const double *term_coef_dev, double *energy_dev, double *expvals_dev,
                            int64_t expvals_stride, void *stream) {
  TQB_REQUIRE(idx_dev && term_ptr_dev && term_z_dev && term_coef_dev && energy_dev && batch >= 1 && shots >= 1 && n_groups >= 1,
              "tqb_expval_from_samples: bad arguments");
  expval_samples_kernel<<<(unsigned)batch, RT, 0, as_stream(stream)>>>((const long long *)idx_dev, (long long)shots, n_groups, term_ptr_dev,
                                                                        term_z_dev, term_coef_dev, energy_dev, expvals_dev,
                                                                        (long long)expvals_stride);
  TQB_CHECK_LAUNCH("expval_samples_kernel");
  return 0;
}

int tqb_sample_shard(const void *state, int n_local, int dtype, const double *chunk_prefix_dev, int64_t n_chunks_total,
                     int64_t chunk_first, int64_t tail_index, const double *uniforms_dev, int64_t shots, int64_t *idx_dev,
                     void *stream) {
  TQB_REQUIRE(state && chunk_prefix_dev && uniforms_dev && idx_dev && n_local >= 12 && n_local < 48 && shots >= 1 &&
                  n_chunks_total >= 1 && chunk_first >= 0,
              "tqb_sample_shard: bad arguments (shards hold whole 4096-amplitude chunks: n_local >= 12)");
  const int chunk_bits = 12;
  const long long c_count = 1ll << (n_local - chunk_bits);
  TQB_REQUIRE(chunk_first + c_count <= n_chunks_total, "tqb_sample_shard: shard chunks exceed the prefix");
  dim3 grid((unsigned)((shots + 127) / 128), 1);
  cudaStream_t st = as_stream(stream);
  if (by_dtype(dtype,
               [&] { sample_kernel<cplx<double>><<<grid, 128, 0, st>>>(CD(state), n_local, chunk_bits, (long long)n_chunks_total, (long long)chunk_first, c_count, (long long)tail_index, chunk_prefix_dev, uniforms_dev, (long long)shots, (long long *)idx_dev, nullptr); },
               [&] { sample_kernel<cplx<float>><<<grid, 128, 0, st>>>(CF(state), n_local, chunk_bits, (long long)n_chunks_total, (long long)chunk_first, c_count, (long long)tail_index, chunk_prefix_dev, uniforms_dev, (long long)shots, (long long *)idx_dev, nullptr); }))
    return -1;
  TQB_CHECK_LAUNCH("sample_kernel");
  return 0;
}

}  // extern "C"
