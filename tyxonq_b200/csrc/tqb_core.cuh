// tqb_core.cuh -- index arithmetic and per-thread tile phases of the fused gate pass.
//
// Everything here is __host__ __device__: the CUDA kernel in tqb_tile.cu calls these
// functions from every thread with __syncthreads() between phases, and the host-side
// emulator in tests/emu/ (test tooling, never shipped or loaded by the package) calls
// the very same functions from nested loops, so the index logic can be checked against
// the oracle on a box without a GPU.
//
// Reference semantics implemented (paths relative to the reference's src/tyxonq/):
//   apply_1q_statevector / apply_2q_statevector   libs/quantum_library/kernels/statevector.py:28-59
//   apply_kqubit_unitary                          libs/quantum_library/kernels/statevector.py:71-129
#pragma once
#include <stdint.h>

#include "../../include/tyxonq_b200.h"

#if defined(__CUDACC__)
#define TQB_HD __host__ __device__ __forceinline__
#else
#define TQB_HD inline
#endif

namespace tqb {

template <typename T>
struct alignas(2 * sizeof(T)) cplx {
  T x, y;
};

// V consecutive amplitudes moved as one 16-byte access (V = 1 for complex128, 2 for complex64).
template <typename T, int V>
struct alignas(sizeof(cplx<T>) * V) cvec {
  cplx<T> e[V];
};

template <typename T>
TQB_HD cplx<T> cmul(cplx<T> a, cplx<T> b) {
  return cplx<T>{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T>
TQB_HD void cmac(cplx<T> &acc, cplx<T> a, cplx<T> b) {
  acc.x += a.x * b.x;
  acc.x -= a.y * b.y;
  acc.y += a.x * b.y;
  acc.y += a.y * b.x;
}

TQB_HD int popc64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __popcll(v);
#else
  return __builtin_popcountll(v);
#endif
}

// Geometry of one pass, passed to the kernel by value.
struct TileGeom {
  int n;                 // local index bits of one batch member
  int m;                 // tile bits
  int L;                 // low contiguous bits
  int h;                 // m - L
  uint64_t global_base;  // high-order bits of a sharded state (rank << n), else 0
  int mat_begin, mat_count;  // matrices staged in shared memory (mat_count = 0: read from global)
  int8_t hb[TQB_MAX_TILE_HIGH];
};

// Insert a zero bit at each of the k ascending positions sb[].
TQB_HD uint32_t insert_zeros32(uint32_t g, const int8_t *sb, int k) {
  for (int j = 0; j < k; ++j) {
    const uint32_t p = (uint32_t)sb[j];
    g = ((g >> p) << (p + 1)) | (g & ((1u << p) - 1u));
  }
  return g;
}
TQB_HD uint64_t insert_zeros64(uint64_t g, const int8_t *sb, int k) {
  for (int j = 0; j < k; ++j) {
    const uint64_t p = (uint64_t)sb[j];
    g = ((g >> p) << (p + 1)) | (g & ((1ull << p) - 1ull));
  }
  return g;
}

// Padded tile layout (lean kernel): every contiguous run of 2^L amplitudes is followed by 16 bytes of padding, so
// that threads whose groups differ in a HIGH tile bit hit different shared-memory banks when the register-resident
// targets are LOW bits (a chain on bits 0-3 keeps 256 contiguous bytes per thread: 8-way conflicts in the plain
// layout, 2-way here).  padL = L when the tile is padded, 0 when it is not; PADSH = log2(16 / sizeof(amplitude)).
template <typename T>
TQB_HD uint32_t pidx(uint32_t e, int padL) {
  constexpr int PADSH = sizeof(cplx<T>) == 16 ? 0 : 1;
  return padL ? e + ((e >> padL) << PADSH) : e;
}

// Index (within one batch member) of element 0 of tile t: t's bits deposited into the
// index bits that are NOT tile bits.
TQB_HD uint64_t tile_base(const TileGeom &g, uint64_t t) {
  return insert_zeros64(t << g.L, g.hb, g.h);
}
// Offset of contiguous run j (j < 2^h) inside a tile.
TQB_HD uint64_t run_offset(const TileGeom &g, uint32_t j) {
  uint64_t o = 0;
  for (int i = 0; i < g.h; ++i) o |= (uint64_t)((j >> i) & 1u) << g.hb[i];
  return o;
}
// State index of tile-local element `local` (roff = table of run_offset for all j).
TQB_HD uint64_t local_to_index(const TileGeom &g, const uint64_t *roff, uint64_t base, uint32_t local) {
  return base | roff[local >> g.L] | (uint64_t)(local & ((1u << g.L) - 1u));
}

// ---- phase 1/3: stage a tile in shared memory, write it back ------------------------------
template <typename T, int V>
TQB_HD void tile_load(cplx<T> *tile, const cplx<T> *state, const TileGeom &g, const uint64_t *roff,
                      uint64_t base, int tid, int nthreads) {
  const uint32_t nvec = (1u << g.m) / V;
  for (uint32_t i = tid; i < nvec; i += nthreads) {
    const uint32_t local = i * V;
    const uint64_t idx = local_to_index(g, roff, base, local);
    *reinterpret_cast<cvec<T, V> *>(tile + local) = *reinterpret_cast<const cvec<T, V> *>(state + idx);
  }
}
template <typename T, int V>
TQB_HD void tile_store(const cplx<T> *tile, cplx<T> *state, const TileGeom &g, const uint64_t *roff,
                       uint64_t base, int tid, int nthreads) {
  const uint32_t nvec = (1u << g.m) / V;
  for (uint32_t i = tid; i < nvec; i += nthreads) {
    const uint32_t local = i * V;
    const uint64_t idx = local_to_index(g, roff, base, local);
    *reinterpret_cast<cvec<T, V> *>(state + idx) = *reinterpret_cast<const cvec<T, V> *>(tile + local);
  }
}

// ---- phase 2: one gate on the staged tile ---------------------------------------------------
// All descriptor fields are copied into registers first: the descriptor lives in shared memory
// next to the tile, so the compiler would otherwise reload it after every tile store.
template <int K>
TQB_HD uint32_t insert_zeros_k(uint32_t g, const uint32_t (&sb)[K]) {
#pragma unroll
  for (int j = 0; j < K; ++j) g = ((g >> sb[j]) << (sb[j] + 1u)) | (g & ((1u << sb[j]) - 1u));
  return g;
}

template <typename T, int K>
TQB_HD void gate_dense(cplx<T> *tile, int m, const tqb_gate &g, const cplx<T> *M, int tid, int nthreads) {
  constexpr int D = 1 << K;
  uint32_t sb[K], off[D];
#pragma unroll
  for (int j = 0; j < K; ++j) sb[j] = (uint32_t)g.sbits[j];
#pragma unroll
  for (int s = 0; s < D; ++s) {
    uint32_t o = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) o |= (uint32_t)((s >> j) & 1) << g.bits[j];
    off[s] = o;
  }
  const uint32_t ngroups = 1u << (m - K);
  if (sizeof(cplx<T>) * D * D <= 128) {  // matrix small enough to live in registers
    cplx<T> Mr[D * D];
#pragma unroll
    for (int i = 0; i < D * D; ++i) Mr[i] = M[i];
    for (uint32_t gi = tid; gi < ngroups; gi += nthreads) {
      const uint32_t base = insert_zeros_k<K>(gi, sb);
      cplx<T> v[D];
#pragma unroll
      for (int s = 0; s < D; ++s) v[s] = tile[base + off[s]];
#pragma unroll
      for (int r = 0; r < D; ++r) {
        cplx<T> acc{0, 0};
#pragma unroll
        for (int c = 0; c < D; ++c) cmac(acc, Mr[r * D + c], v[c]);
        tile[base + off[r]] = acc;
      }
    }
  } else {
    for (uint32_t gi = tid; gi < ngroups; gi += nthreads) {
      const uint32_t base = insert_zeros_k<K>(gi, sb);
      cplx<T> v[D];
#pragma unroll
      for (int s = 0; s < D; ++s) v[s] = tile[base + off[s]];
#pragma unroll 2
      for (int r = 0; r < D; ++r) {
        cplx<T> acc{0, 0};
#pragma unroll
        for (int c = 0; c < D; ++c) cmac(acc, M[r * D + c], v[c]);
        tile[base + off[r]] = acc;
      }
    }
  }
}

// PAIR: 2x2 block on (pattern A, pattern B); SWAP: exchange A and B (x, cx, swap: no arithmetic).
template <typename T, int K, bool SWAP>
TQB_HD void gate_pair_k(cplx<T> *tile, const TileGeom &geo, const uint64_t *roff, uint64_t gbase,
                        const tqb_gate &g, const cplx<T> *M, int tid, int nthreads) {
  uint32_t sb[K];
#pragma unroll
  for (int j = 0; j < K; ++j) sb[j] = (uint32_t)g.sbits[j];
  const uint32_t off_a = g.off_a, off_b = g.off_b;
  const uint64_t zmask = g.zmask;
  const uint32_t ngroups = 1u << (geo.m - K);
  if (SWAP) {
    for (uint32_t gi = tid; gi < ngroups; gi += nthreads) {
      const uint32_t base = insert_zeros_k<K>(gi, sb);
      const cplx<T> a = tile[base + off_a], b = tile[base + off_b];
      tile[base + off_a] = b;
      tile[base + off_b] = a;
    }
    return;
  }
  cplx<T> M0[4], M1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) M0[i] = M[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) M1[i] = zmask ? M[4 + i] : M[i];
  for (uint32_t gi = tid; gi < ngroups; gi += nthreads) {
    const uint32_t base = insert_zeros_k<K>(gi, sb);
    const uint32_t ia = base + off_a, ib = base + off_b;
    const cplx<T> a = tile[ia], b = tile[ib];
    bool odd = false;
    if (zmask) odd = popc64(local_to_index(geo, roff, gbase, base) & zmask) & 1;
    cplx<T> ra{0, 0}, rb{0, 0};
    cmac(ra, odd ? M1[0] : M0[0], a);
    cmac(ra, odd ? M1[1] : M0[1], b);
    cmac(rb, odd ? M1[2] : M0[2], a);
    cmac(rb, odd ? M1[3] : M0[3], b);
    tile[ia] = ra;
    tile[ib] = rb;
  }
}

template <typename T, bool SWAP>
TQB_HD void gate_pair(cplx<T> *tile, const TileGeom &geo, const uint64_t *roff, uint64_t gbase,
                      const tqb_gate &g, const cplx<T> *M, int tid, int nthreads) {
  switch (g.k) {
    case 1: gate_pair_k<T, 1, SWAP>(tile, geo, roff, gbase, g, M, tid, nthreads); break;
    case 2: gate_pair_k<T, 2, SWAP>(tile, geo, roff, gbase, g, M, tid, nthreads); break;
    case 3: gate_pair_k<T, 3, SWAP>(tile, geo, roff, gbase, g, M, tid, nthreads); break;
    case 4: gate_pair_k<T, 4, SWAP>(tile, geo, roff, gbase, g, M, tid, nthreads); break;
    default: break;
  }
}

// DIAG: bits[j] < 64 -> the table-index bit j is tile-local bit bits[j]; bits[j] >= 64 -> it is
// index bit (bits[j] - 64) outside the tile, constant for the whole tile (read from gbase).
template <typename T, int K>
TQB_HD void gate_diag_k(cplx<T> *tile, int m, uint64_t gbase, const tqb_gate &g, const cplx<T> *tab,
                        int tid, int nthreads, int padL = 0) {
  uint32_t pos[K];
  uint32_t cpart = 0, lmask = 0;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const uint32_t b = (uint32_t)(uint8_t)g.bits[j];
    if (b & 64u) {
      cpart |= (uint32_t)((gbase >> (b & 63u)) & 1ull) << j;
      pos[j] = 0;
    } else {
      pos[j] = b;
      lmask |= 1u << j;
    }
  }
  const uint32_t nel = 1u << m;
  for (uint32_t e = tid; e < nel; e += nthreads) {
    uint32_t t = cpart;
#pragma unroll
    for (int j = 0; j < K; ++j)
      if ((lmask >> j) & 1u) t |= ((e >> pos[j]) & 1u) << j;
    const uint32_t pe = pidx<T>(e, padL);
    tile[pe] = cmul(tile[pe], tab[t]);
  }
}

template <typename T>
TQB_HD void gate_diag(cplx<T> *tile, int m, uint64_t gbase, const tqb_gate &g, const cplx<T> *tab, int tid, int nthreads,
                      int padL = 0) {
  switch (g.k) {
    case 1: gate_diag_k<T, 1>(tile, m, gbase, g, tab, tid, nthreads, padL); break;
    case 2: gate_diag_k<T, 2>(tile, m, gbase, g, tab, tid, nthreads, padL); break;
    case 3: gate_diag_k<T, 3>(tile, m, gbase, g, tab, tid, nthreads, padL); break;
    case 4: gate_diag_k<T, 4>(tile, m, gbase, g, tab, tid, nthreads, padL); break;
    case 5: gate_diag_k<T, 5>(tile, m, gbase, g, tab, tid, nthreads, padL); break;
    case 6: gate_diag_k<T, 6>(tile, m, gbase, g, tab, tid, nthreads, padL); break;
    default: break;
  }
}

// MUX: dense 1-qubit gate on tile-local bit bits[0] whose 2x2 matrix is selected by ONE control
// bit (matrix U0 at mat[0..4) when the control is 0, U1 at mat[4..8) when it is 1).  This is
// what "cx followed / preceded by a 1-qubit gate on its target" fuses to on the host: the sweep
// costs the same as a plain 1-qubit gate.  bits[1] < 64: the control is tile-local bit bits[1];
// bits[1] >= 64: it is index bit bits[1] - 64 outside the tile (constant for the whole tile).
template <typename T>
TQB_HD void mux_sweep(cplx<T> *tile, uint32_t ngroups, uint32_t p0, uint32_t p1, bool two, uint32_t fixed,
                      uint32_t tstride, const cplx<T> *M, int tid, int nthreads, int padL = 0) {
  const cplx<T> m00 = M[0], m01 = M[1], m10 = M[2], m11 = M[3];
#pragma unroll 4
  for (uint32_t gi = tid; gi < ngroups; gi += nthreads) {
    uint32_t base = ((gi >> p0) << (p0 + 1u)) | (gi & ((1u << p0) - 1u));
    if (two) base = ((base >> p1) << (p1 + 1u)) | (base & ((1u << p1) - 1u));
    base |= fixed;
    const uint32_t ia = pidx<T>(base, padL), ib = pidx<T>(base + tstride, padL);
    const cplx<T> a = tile[ia], b = tile[ib];
    cplx<T> x{0, 0}, y{0, 0};
    cmac(x, m00, a); cmac(x, m01, b);
    cmac(y, m10, a); cmac(y, m11, b);
    tile[ia] = x;
    tile[ib] = y;
  }
}

template <typename T>
TQB_HD void gate_mux(cplx<T> *tile, int m, uint64_t gbase, const tqb_gate &g, const cplx<T> *M, int tid, int nthreads,
                     int padL = 0) {
  const uint32_t tb = (uint32_t)g.bits[0];
  const uint32_t cb = (uint32_t)(uint8_t)g.bits[1];
  if (cb & 64u) {  // control outside the tile: one matrix for the whole tile
    const uint32_t cv = (uint32_t)((gbase >> (cb & 63u)) & 1ull);
    mux_sweep<T>(tile, 1u << (m - 1), tb, 0, false, 0, 1u << tb, M + 4 * cv, tid, nthreads, padL);
  } else {
    const uint32_t p0 = tb < cb ? tb : cb, p1 = tb < cb ? cb : tb;
    mux_sweep<T>(tile, 1u << (m - 2), p0, p1, true, 0, 1u << tb, M, tid, nthreads, padL);
    mux_sweep<T>(tile, 1u << (m - 2), p0, p1, true, 1u << cb, 1u << tb, M + 4, tid, nthreads, padL);
  }
}

// CHAIN: R (2 or 3) 1-qubit layers applied to 2^R register-resident amplitudes per thread: one
// shared-memory round trip for R gates (the sweeps are shared-memory-bandwidth bound, so this is
// what buys the factor R).  Layer i acts on tile-local bit bits[i]; its 2x2 is selected by
//   i = 0: the outer control bit bits[R] (< 64 tile-local, 64+p outside the tile, 127 = none)
//   i > 0: the value of bit bits[i-1] AFTER layer i-1 (the cx chain of a hardware-efficient layer;
//          for independent gates the host stores the same matrix twice).
// Matrix layout: layer i -> M[8i .. 8i+4) (selector 0), M[8i+4 .. 8i+8) (selector 1).
// All register indexing is compile-time: register bit i <-> layer i.
template <typename T, int R, int I, int SEL>
TQB_HD void chain_layer(cplx<T> (&v)[1 << R], const cplx<T> *M) {
  const cplx<T> m00 = M[0], m01 = M[1], m10 = M[2], m11 = M[3];
#pragma unroll
  for (int s = 0; s < (1 << R); ++s) {
    if (s & (1 << I)) continue;
    if (I > 0 && SEL >= 0 && ((s >> (I > 0 ? I - 1 : 0)) & 1) != SEL) continue;
    const cplx<T> a = v[s], b = v[s | (1 << I)];
    cplx<T> x{0, 0}, y{0, 0};
    cmac(x, m00, a); cmac(x, m01, b);
    cmac(y, m10, a); cmac(y, m11, b);
    v[s] = x;
    v[s | (1 << I)] = y;
  }
}

// layer I > 0, both selector values in one unrolled block: 2^(R-1) independent pairs keep the FP pipe busy
template <typename T, int R, int I>
TQB_HD void chain_layer_sel(cplx<T> (&v)[1 << R], const cplx<T> *M) {
  const cplx<T> a00 = M[0], a01 = M[1], a10 = M[2], a11 = M[3];
  const cplx<T> b00 = M[4], b01 = M[5], b10 = M[6], b11 = M[7];
#pragma unroll
  for (int s = 0; s < (1 << R); ++s) {
    if (s & (1 << I)) continue;
    const bool sel = (s >> (I > 0 ? I - 1 : 0)) & 1;
    const cplx<T> a = v[s], b = v[s | (1 << I)];
    cplx<T> x{0, 0}, y{0, 0};
    cmac(x, sel ? b00 : a00, a); cmac(x, sel ? b01 : a01, b);
    cmac(y, sel ? b10 : a10, a); cmac(y, sel ? b11 : a11, b);
    v[s] = x;
    v[s | (1 << I)] = y;
  }
}

template <typename T, int R>
TQB_HD void gate_chain(cplx<T> *tile, int m, uint64_t gbase, const tqb_gate &g, const cplx<T> *M, int tid, int nthreads,
                       int padL = 0) {
  uint32_t tb[R];
#pragma unroll
  for (int i = 0; i < R; ++i) tb[i] = (uint32_t)g.bits[i];
  const uint32_t cb = (uint32_t)(uint8_t)g.bits[R];
  const bool ctrl_local = cb < 64u;
  uint32_t cv_fixed = 0;
  if (!ctrl_local && cb != 127u) cv_fixed = (uint32_t)((gbase >> (cb & 63u)) & 1ull);
  constexpr int NZ = R + 1;
  uint32_t sb[NZ];  // ascending positions where zero bits are inserted (targets [+ local control])
#pragma unroll
  for (int j = 0; j < NZ; ++j) sb[j] = (uint32_t)g.sbits[j];
  const int nz = ctrl_local ? R + 1 : R;
  const uint32_t free_bits = (uint32_t)(m - nz);
  const uint32_t ngroups = 1u << (m - R);  // control value = top bit of the group counter when local
  auto locate = [&](uint32_t gi, uint32_t &base, uint32_t &cv) {
    uint32_t lo = gi;
    cv = cv_fixed;
    if (ctrl_local) {
      cv = gi >> free_bits;
      lo = gi & ((1u << free_bits) - 1u);
    }
    base = lo;
#pragma unroll
    for (int j = 0; j < NZ; ++j)
      if (j < nz) base = ((base >> sb[j]) << (sb[j] + 1u)) | (base & ((1u << sb[j]) - 1u));
    if (ctrl_local) base |= cv << cb;
  };
  auto offset = [&](uint32_t base, int s) {
    uint32_t o = base;
#pragma unroll
    for (int i = 0; i < R; ++i)
      if (s & (1 << i)) o |= 1u << tb[i];
    return pidx<T>(o, padL);
  };
  uint32_t gi = tid;
  // two groups per iteration: two independent dependency chains keep the FP pipe fed (the sweep is latency bound
  // with the 3-4 consumer warps per scheduler that the tile buffers leave room for)
  if (sizeof(T) == 4 && ngroups % (2u * (uint32_t)nthreads) == 0) {  // complex64 only: complex128 would spill
    for (; gi < ngroups; gi += 2u * nthreads) {
      uint32_t ba, ca, bb, cb2;
      locate(gi, ba, ca);
      locate(gi + nthreads, bb, cb2);
      cplx<T> va[1 << R], vb[1 << R];
#pragma unroll
      for (int s = 0; s < (1 << R); ++s) {
        va[s] = tile[offset(ba, s)];
        vb[s] = tile[offset(bb, s)];
      }
      chain_layer<T, R, 0, -1>(va, M + 4 * ca);
      chain_layer<T, R, 0, -1>(vb, M + 4 * cb2);
      chain_layer_sel<T, R, 1>(va, M + 8);
      chain_layer_sel<T, R, 1>(vb, M + 8);
      if (R > 2) {
        chain_layer_sel<T, R, (R > 2 ? 2 : 1)>(va, M + 16);
        chain_layer_sel<T, R, (R > 2 ? 2 : 1)>(vb, M + 16);
      }
#pragma unroll
      for (int s = 0; s < (1 << R); ++s) {
        tile[offset(ba, s)] = va[s];
        tile[offset(bb, s)] = vb[s];
      }
    }
    return;
  }
  for (; gi < ngroups; gi += nthreads) {
    uint32_t base, cv;
    locate(gi, base, cv);
    cplx<T> v[1 << R];
#pragma unroll
    for (int s = 0; s < (1 << R); ++s) v[s] = tile[offset(base, s)];
    chain_layer<T, R, 0, -1>(v, M + 4 * cv);
    chain_layer_sel<T, R, 1>(v, M + 8);
    if (R > 2) chain_layer_sel<T, R, (R > 2 ? 2 : 1)>(v, M + 16);
#pragma unroll
    for (int s = 0; s < (1 << R); ++s) tile[offset(base, s)] = v[s];
  }
}

// CHAIN in rotation form (g.off_a = 4 + 2*TYPE + MUXED).  A 1-qubit unitary whose column ratio is imaginary
// (rz.rx products: every layer of a hardware-efficient or Trotter circuit) factors as U = M.D with
// M = [[a, i r], [i r, a]] (TYPE 0) and D diagonal; a real-ratio unitary (ry, h) as M = [[a, -r], [r, a]]
// (TYPE 1).  Inside a chain every D acts on a bit no earlier layer touches, so all of them commute to the FRONT
// as one table P over the 2^R register indices, and M commutes with the fused cx (X.M = M.X for TYPE 0,
// X.M = M^T.X for TYPE 1), which therefore is an input rename.  Per amplitude: one complex multiply by P and
// R real rotations = 4 + 4R multiply-adds instead of 8R, and 2^R + R matrix entries instead of 8R.
// Layers i > 0 run in SCALED form: the larger of (a_i, r_i) is divided out -- M_i = a_i [[1, i t], [i t, 1]] with
// t = r_i / a_i (mode 0, |a_i| >= |r_i|) or M_i = r_i [[c, i], [i, c]] with c = a_i / r_i (mode 1) -- and the host
// multiplies the product of those real factors into layer 0's pair (a_0, r_0): one fused multiply-add per real
// component instead of a multiply and a multiply-add, 4 + 4 + 2(R-1) instructions per amplitude instead of 4 + 4R.
// Data: M[0 .. 2^R) = P (followed by a copy with index bit 0 flipped when the gate has an outer control),
// then (S a_0, S r_0) and (t_i or c_i, mode_i) for i > 0.  MUXED: layers i > 0
// are selected by bit bits[i-1] (cx after the gate); layer 0 by the outer control when bits[R] != 127 (run-time:
// loads and P index flip with its value).
template <typename T, int R, int I, int TYPE, bool MUXED>
TQB_HD void rot_layer(cplx<T> (&v)[1 << R], const T a, const T r) {
#pragma unroll
  for (int s = 0; s < (1 << R); ++s) {
    if (s & (1 << I)) continue;
    const bool sel = MUXED && I > 0 && ((s >> (I > 0 ? I - 1 : 0)) & 1);
    const int lo = s, hi = s | (1 << I);
    const cplx<T> x0 = v[sel ? hi : lo], x1 = v[sel ? lo : hi];
    const T rr = (TYPE == 1 && sel) ? -r : r;
    cplx<T> y0, y1;
    if (TYPE == 0) {
      y0.x = a * x0.x - rr * x1.y;
      y0.y = a * x0.y + rr * x1.x;
      y1.x = a * x1.x - rr * x0.y;
      y1.y = a * x1.y + rr * x0.x;
    } else {
      y0.x = a * x0.x - rr * x1.x;
      y0.y = a * x0.y - rr * x1.y;
      y1.x = a * x1.x + rr * x0.x;
      y1.y = a * x1.y + rr * x0.y;
    }
    v[lo] = y0;
    v[hi] = y1;
  }
}

// scaled form of layer I > 0: k = t (inv false) or c (inv true); the branch is uniform over the tile
template <typename T, int R, int I, int TYPE, bool MUXED>
TQB_HD void rot_layer_scaled(cplx<T> (&v)[1 << R], const T k, const bool inv) {
  if (!inv) {
#pragma unroll
    for (int s = 0; s < (1 << R); ++s) {
      if (s & (1 << I)) continue;
      const bool sel = MUXED && ((s >> (I > 0 ? I - 1 : 0)) & 1);
      const int lo = s, hi = s | (1 << I);
      const cplx<T> x0 = v[sel ? hi : lo], x1 = v[sel ? lo : hi];
      cplx<T> y0, y1;
      if (TYPE == 0) {
        y0.x = x0.x - k * x1.y;
        y0.y = x0.y + k * x1.x;
        y1.x = x1.x - k * x0.y;
        y1.y = x1.y + k * x0.x;
      } else if (sel) {
        y0.x = x0.x + k * x1.x;
        y0.y = x0.y + k * x1.y;
        y1.x = x1.x - k * x0.x;
        y1.y = x1.y - k * x0.y;
      } else {
        y0.x = x0.x - k * x1.x;
        y0.y = x0.y - k * x1.y;
        y1.x = x1.x + k * x0.x;
        y1.y = x1.y + k * x0.y;
      }
      v[lo] = y0;
      v[hi] = y1;
    }
  } else {
#pragma unroll
    for (int s = 0; s < (1 << R); ++s) {
      if (s & (1 << I)) continue;
      const bool sel = MUXED && ((s >> (I > 0 ? I - 1 : 0)) & 1);
      const int lo = s, hi = s | (1 << I);
      const cplx<T> x0 = v[sel ? hi : lo], x1 = v[sel ? lo : hi];
      cplx<T> y0, y1;
      if (TYPE == 0) {
        y0.x = k * x0.x - x1.y;
        y0.y = k * x0.y + x1.x;
        y1.x = k * x1.x - x0.y;
        y1.y = k * x1.y + x0.x;
      } else if (sel) {
        y0.x = k * x0.x + x1.x;
        y0.y = k * x0.y + x1.y;
        y1.x = k * x1.x - x0.x;
        y1.y = k * x1.y - x0.y;
      } else {
        y0.x = k * x0.x - x1.x;
        y0.y = k * x0.y - x1.y;
        y1.x = k * x1.x + x0.x;
        y1.y = k * x1.y + x0.y;
      }
      v[lo] = y0;
      v[hi] = y1;
    }
  }
}

// index of the lowest set bit of s in 1..15 (s is a compile-time constant after unrolling)
#define TQB_LOW_BIT(s) (((s) & 1) ? 0 : ((s) & 2) ? 1 : ((s) & 4) ? 2 : 3)

// A rotation-form CHAIN descriptor decoded for the sweep (48 bytes, three 16-byte loads): the lean kernel decodes
// every gate of the pass ONCE per CTA into shared memory (tile_pass_lean_kernel), the general kernel decodes in place.
//   flags: bit 0 control is tile-local, bit 1 has a control, bit 2 table is all ones, bits 3-4 E, bit 5 MUXED,
//          bit 6 TYPE, bits 8-13 control position (tile-local bit, or index bit outside the tile), bits 16-20 number
//          of free bits (m - targets - local control), bits 21-23 padL (pidx), bits 24-28 m, bits 29-30 R - 1
//   d[i]:  byte stride of target i;  lm[j]: (1 << p_j) - 1 for the ascending positions p_j where a zero bit is
//          inserted into the group counter (targets + local control), all ones beyond the last one
struct alignas(16) RotDesc {
  int32_t d[4];
  uint32_t lm[5];
  uint32_t flags;
  uint32_t xb;   // extra table bits: xb0 | xb1 << 8 (encoding of tqb_gate.bits: < 64 tile-local, 64 + p outside) | mat_bstride << 16
  uint32_t mat;  // tqb_gate.mat_off
};

template <typename T>
TQB_HD RotDesc rot_decode(const tqb_gate &g, int m, int padL = 0) {   // padL: see pidx (0 < padL <= 7, or 0 = plain layout)
  RotDesc r;
  const int R = g.k;
  const uint32_t cb = (uint32_t)(uint8_t)g.bits[R];
  const bool ctrl_local = cb < 64u, has_ctrl = cb != 127u;
  const int nz = ctrl_local ? R + 1 : R;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t b = (uint32_t)g.bits[i];
    r.d[i] = i < R ? (int32_t)((sizeof(cplx<T>) << b) + ((padL && b >= (uint32_t)padL) ? (16u << (b - (uint32_t)padL)) : 0u)) : 0;
  }
#pragma unroll
  for (int j = 0; j < 5; ++j) r.lm[j] = j < nz ? (1u << (uint32_t)g.sbits[j]) - 1u : 0xffffffffu;
  r.flags = (ctrl_local ? 1u : 0u) | (has_ctrl ? 2u : 0u) | ((g.off_b & 128u) ? 4u : 0u) | ((g.off_b & 3u) << 3) |
            ((g.off_a & 3u) << 5) | ((cb & 63u) << 8) | ((uint32_t)(m - nz) << 16) | ((uint32_t)padL << 21) | ((uint32_t)m << 24) |
            ((uint32_t)(R - 1) << 29) | ((g.off_b & 4096u) ? 0x80000000u : 0u);   // bit 31: layer 0 in the t form (its factor is in the table)
  r.xb = (uint32_t)(uint8_t)g.bits[R + 1] | ((uint32_t)(uint8_t)g.bits[R + 2] << 8) | (g.mat_bstride << 16);  // (stride <= 2^7 + 4)
  r.mat = g.mat_off;
  return r;
}

// Sync: called by every thread exactly once per gate, right before its first access to the tile -- the lean kernel
// passes its inter-gate barrier (or the wait for the tile to land) here, so that descriptor decoding, group location
// and address arithmetic of gate g+1 overlap the other warps' tail of gate g.  NoSync: callers that synchronise outside.
struct NoSync {
  TQB_HD void operator()() const {}
};

template <typename T, int R, int TYPE, bool MUXED, class Sync = NoSync, int UNR = 1>   // UNR: groups in flight per thread
TQB_HD void chain_rot_sweep(cplx<T> *tile, uint64_t gbase, const RotDesc &rd, const cplx<T> *M, int tid, int nthreads,
                            Sync sync = Sync()) {
  constexpr int D = 1 << R;
  constexpr int NZ = R + 1;
  const uint32_t f = rd.flags;
  const bool ctrl_local = (f & 1u) != 0u, has_ctrl = (f & 2u) != 0u, unit_p = (f & 4u) != 0u;
  const uint32_t E = (f >> 3) & 3u;
  const uint32_t cb = (f >> 8) & 63u;
  const uint32_t free_bits = (f >> 16) & 31u;
  const uint32_t ngroups = 1u << (((f >> 24) & 31u) - (uint32_t)R);
  const uint32_t padL = (f >> 21) & 7u;
  uint32_t cv_fixed = 0;
  if (!ctrl_local && has_ctrl) cv_fixed = (uint32_t)((gbase >> cb) & 1ull);
  uint32_t lm[NZ];
#pragma unroll
  for (int j = 0; j < NZ; ++j) lm[j] = rd.lm[j];
  int32_t d[R];
#pragma unroll
  for (int i = 0; i < R; ++i) d[i] = rd.d[i] & ~(int32_t)(sizeof(cplx<T>) - 1);  // (a no-op that tells the compiler the accesses stay aligned)
  // table: 2^(R+E) entries, index = register index + (extras << R); a gate with a control carries a second copy
  // with register-index bit 0 flipped (control = 1: the fused cx renames the inputs of layer 0), so that the table
  // is read at compile-time offsets from a per-thread base.  The layer coefficients follow the table(s).
  const uint32_t tab = 1u << (R + E);
  const cplx<T> *coef = M + (has_ctrl ? 2u * tab : tab);
  const T a0 = coef[0].x, r0 = coef[0].y;
  T kk[R];
  bool inv[R];
#pragma unroll
  for (int i = 1; i < R; ++i) {
    const cplx<T> c = coef[i];
    kk[i] = c.x;
    inv[i] = c.y != (T)0;
  }
  kk[0] = 0;
  inv[0] = false;
  char *const tbytes = reinterpret_cast<char *>(tile);
  bool synced = false;
  // (two groups per iteration was measured slower here as well: profiles/r01_tile_sweep.md)
#pragma unroll(UNR)
  for (uint32_t gi = tid; gi < ngroups; gi += nthreads) {
    uint32_t base = gi, cv = cv_fixed;
    if (ctrl_local) {
      cv = gi >> free_bits;
      base = gi & ((1u << free_bits) - 1u);
    }
#pragma unroll
    for (int j = 0; j < NZ; ++j) base += base & ~lm[j];  // insert a zero bit at position p_j (no-op when lm is all ones)
    if (ctrl_local) base |= cv << cb;
    // register s holds input amplitude s ^ cv (cv toggles bit 0 = layer 0's bit): start at the flipped element and
    // step layer 0's stride backwards
    char *q[D];
    const uint32_t boff = base * (uint32_t)sizeof(cplx<T>) + (padL ? ((base >> padL) << 4) : 0u);  // byte offset of the group's element 0
    q[0] = tbytes + boff + (cv ? d[0] : 0);
    const int32_t d0 = (cv ? -d[0] : d[0]) & ~(int32_t)(sizeof(cplx<T>) - 1);
#pragma unroll
    for (int s = 1; s < D; ++s) q[s] = q[s & (s - 1)] + (TQB_LOW_BIT(s) == 0 ? d0 : d[TQB_LOW_BIT(s) < R ? TQB_LOW_BIT(s) : 0]);
    if (!synced) {
      sync();
      synced = true;
    }
    cplx<T> v[D];
#ifdef TQB_PROFILE_SWITCHES   // profiling only (results are wrong): flags bit 14 = no tile loads, bit 7 = no arithmetic, bit 15 = no stores
    if (f & (1u << 14)) {
#pragma unroll
      for (int s = 0; s < D; ++s) v[s] = cplx<T>{(T)(gi + s), (T)s};
    } else
#endif
    if (unit_p) {
#pragma unroll
      for (int s = 0; s < D; ++s) v[s] = *reinterpret_cast<const cplx<T> *>(q[s]);
    } else {
      uint32_t x = 0;
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if ((uint32_t)j < E) {
          const uint32_t xbj = (rd.xb >> (8 * j)) & 0xffu;
          x |= (xbj < 64u ? ((base >> xbj) & 1u) : (uint32_t)((gbase >> (xbj & 63u)) & 1ull)) << j;
        }
      const cplx<T> *P = M + ((size_t)x << R) + (cv ? tab : 0u);
#pragma unroll
      for (int s = 0; s < D; ++s) v[s] = cmul(*reinterpret_cast<const cplx<T> *>(q[s]), P[s]);
    }
#ifdef TQB_PROFILE_SWITCHES
    if (!(f & (1u << 7)))
#endif
    {
      if (f >> 31) rot_layer_scaled<T, R, 0, TYPE, false>(v, (TYPE == 1 && cv) ? -a0 : a0, false);   // coef[0] = (t0, 0)
      else rot_layer<T, R, 0, TYPE, false>(v, a0, (TYPE == 1 && cv) ? -r0 : r0);
      rot_layer_scaled<T, R, 1, TYPE, MUXED>(v, kk[1], inv[1]);
      if (R > 2) rot_layer_scaled<T, R, (R > 2 ? 2 : 1), TYPE, MUXED>(v, kk[R > 2 ? 2 : 1], inv[R > 2 ? 2 : 1]);
      if (R > 3) rot_layer_scaled<T, R, (R > 3 ? 3 : 1), TYPE, MUXED>(v, kk[R > 3 ? 3 : 1], inv[R > 3 ? 3 : 1]);
    }
    // outputs are not renamed: register s goes to element s of the group
    char *w[D];
    w[0] = tbytes + boff;
#pragma unroll
    for (int s = 1; s < D; ++s) w[s] = w[s & (s - 1)] + d[TQB_LOW_BIT(s) < R ? TQB_LOW_BIT(s) : 0];
#ifdef TQB_PROFILE_SWITCHES
    if (f & (1u << 15)) {   // keep the values alive with one store
      cplx<T> acc = v[0];
#pragma unroll
      for (int s = 1; s < D; ++s) { acc.x += v[s].x; acc.y += v[s].y; }
      if (acc.x == (T)1.2345e-300) *reinterpret_cast<cplx<T> *>(w[0]) = acc;
      continue;
    }
#endif
#pragma unroll
    for (int s = 0; s < D; ++s) *reinterpret_cast<cplx<T> *>(w[s]) = v[s];
  }
  if (!synced) sync();
}

template <typename T, class Sync, int UNR = 1>
TQB_HD void chain_rot_dispatch4(cplx<T> *tile, uint64_t gbase, const RotDesc &rd, const cplx<T> *M, int tid, int nthreads, Sync sync) {
  switch ((rd.flags >> 5) & 3u) {
    case 0: chain_rot_sweep<T, 4, 0, false, Sync, UNR>(tile, gbase, rd, M, tid, nthreads, sync); break;
    case 1: chain_rot_sweep<T, 4, 0, true, Sync, UNR>(tile, gbase, rd, M, tid, nthreads, sync); break;
    case 2: chain_rot_sweep<T, 4, 1, false, Sync, UNR>(tile, gbase, rd, M, tid, nthreads, sync); break;
    default: chain_rot_sweep<T, 4, 1, true, Sync, UNR>(tile, gbase, rd, M, tid, nthreads, sync); break;
  }
}

// dispatch on (R, TYPE, MUXED) of a decoded rotation-form chain; M = base of the matrix buffer
template <typename T, class Sync = NoSync, int MAXR = 4, int UNR = 1>   // MAXR = 3: the 4-layer bodies are compiled out (few-register variant)
TQB_HD void chain_rot_dispatch(cplx<T> *tile, uint64_t gbase, const RotDesc &rd, const cplx<T> *mats, int tid, int nthreads,
                               Sync sync = Sync(), size_t bm = 0) {   // bm: batch member of the tile (per-member matrices)
  const cplx<T> *M = mats + rd.mat + bm * (size_t)(rd.xb >> 16);
  switch ((rd.flags >> 5) & 3u | ((rd.flags >> 29) & 3u) << 2) {   // (off_a & 3) = 2 * TYPE + MUXED, R - 1
#define TQB_ROT(R) \
    case ((R - 1) << 2) | 0: chain_rot_sweep<T, R, 0, false, Sync, UNR>(tile, gbase, rd, M, tid, nthreads, sync); break; \
    case ((R - 1) << 2) | 1: chain_rot_sweep<T, R, 0, true, Sync, UNR>(tile, gbase, rd, M, tid, nthreads, sync); break;  \
    case ((R - 1) << 2) | 2: chain_rot_sweep<T, R, 1, false, Sync, UNR>(tile, gbase, rd, M, tid, nthreads, sync); break; \
    case ((R - 1) << 2) | 3: chain_rot_sweep<T, R, 1, true, Sync, UNR>(tile, gbase, rd, M, tid, nthreads, sync); break;
    TQB_ROT(2) TQB_ROT(3)
#undef TQB_ROT
    default:
      if (MAXR >= 4) chain_rot_dispatch4<T, Sync, UNR>(tile, gbase, rd, M, tid, nthreads, sync);
      else sync();
      break;
  }
}

template <typename T, int R, int TYPE, bool MUXED>
TQB_HD void gate_chain_rot(cplx<T> *tile, int m, uint64_t gbase, const tqb_gate &g, const cplx<T> *M, int tid, int nthreads) {
  RotDesc rd = rot_decode<T>(g, m);
  chain_rot_sweep<T, R, TYPE, MUXED>(tile, gbase, rd, M, tid, nthreads);
}

// R = 4 exists in rotation form only (16 amplitudes + 4 coefficient pairs fit the register budget; the general
// form would need 32 matrix registers on top)
template <typename T>
TQB_HD void gate_chain4(cplx<T> *tile, int m, uint64_t gbase, const tqb_gate &g, const cplx<T> *M, int tid, int nthreads) {
  switch (g.off_a) {
    case 4: gate_chain_rot<T, 4, 0, false>(tile, m, gbase, g, M, tid, nthreads); break;
    case 5: gate_chain_rot<T, 4, 0, true>(tile, m, gbase, g, M, tid, nthreads); break;
    case 6: gate_chain_rot<T, 4, 1, false>(tile, m, gbase, g, M, tid, nthreads); break;
    case 7: gate_chain_rot<T, 4, 1, true>(tile, m, gbase, g, M, tid, nthreads); break;
    default: break;
  }
}

template <typename T, int R>
TQB_HD void gate_chain_any(cplx<T> *tile, int m, uint64_t gbase, const tqb_gate &g, const cplx<T> *M, int tid, int nthreads) {
  switch (g.off_a) {
    case 4: gate_chain_rot<T, R, 0, false>(tile, m, gbase, g, M, tid, nthreads); break;
    case 5: gate_chain_rot<T, R, 0, true>(tile, m, gbase, g, M, tid, nthreads); break;
    case 6: gate_chain_rot<T, R, 1, false>(tile, m, gbase, g, M, tid, nthreads); break;
    case 7: gate_chain_rot<T, R, 1, true>(tile, m, gbase, g, M, tid, nthreads); break;
    default: gate_chain<T, R>(tile, m, gbase, g, M, tid, nthreads); break;
  }
}

// gbase = global_base | tile base: the state index of tile element 0, high shard bits included.
// MAXK bounds the dense gate size this instantiation can execute (2 = light, few registers;
// 4 = heavy): the host picks the variant per pass.  mats = base of the matrix buffer (staged in
// shared memory when the pass says so), bm = batch member of the tile.
template <typename T, int MAXK>
TQB_HD void tile_apply_gate(cplx<T> *tile, const TileGeom &geo, const uint64_t *roff, uint64_t gbase,
                            const tqb_gate &g, const cplx<T> *mats, size_t bm, int tid, int nthreads) {
  const cplx<T> *mat = mats + g.mat_off + bm * g.mat_bstride;
  switch (g.kind) {
    case TQB_GATE_DENSE:
      switch (g.k) {
        case 1: mux_sweep<T>(tile, 1u << (geo.m - 1), (uint32_t)g.bits[0], 0, false, 0, 1u << g.bits[0], mat, tid, nthreads); break;
        case 2: gate_dense<T, 2>(tile, geo.m, g, mat, tid, nthreads); break;
        case 3: if (MAXK >= 3) gate_dense<T, 3>(tile, geo.m, g, mat, tid, nthreads); break;
        case 4: if (MAXK >= 4) gate_dense<T, 4>(tile, geo.m, g, mat, tid, nthreads); break;
        default: break;
      }
      break;
    case TQB_GATE_DIAG: gate_diag<T>(tile, geo.m, gbase, g, mat, tid, nthreads); break;
    case TQB_GATE_PAIR: gate_pair<T, false>(tile, geo, roff, gbase, g, mat, tid, nthreads); break;
    case TQB_GATE_SWAP: gate_pair<T, true>(tile, geo, roff, gbase, g, mat, tid, nthreads); break;
    case TQB_GATE_MUX: gate_mux<T>(tile, geo.m, gbase, g, mat, tid, nthreads); break;
    case TQB_GATE_CHAIN:
      if (g.k == 2) gate_chain_any<T, 2>(tile, geo.m, gbase, g, mat, tid, nthreads);
      else if (g.k == 3) gate_chain_any<T, 3>(tile, geo.m, gbase, g, mat, tid, nthreads);
      else if (g.k == 4) gate_chain4<T>(tile, geo.m, gbase, g, mat, tid, nthreads);
      break;
    default: break;
  }
}

// Lean dispatch: passes that hold only 1-qubit-layer gates (DENSE k = 1, DIAG, MUX, rotation-form CHAIN) -- every pass
// of a hardware-efficient / QAOA / Trotter circuit after fusion.  The kernel built on it carries none of the dense
// k >= 2 / PAIR / SWAP / general-CHAIN code, which is what keeps its loop state in registers.
template <typename T, bool WITH_ROT = true>   // WITH_ROT = false: the caller handles rotation-form chains itself
TQB_HD void tile_apply_gate_lean(cplx<T> *tile, int m, uint64_t gbase, const tqb_gate &g, const cplx<T> *mats, int tid,
                                 int nthreads, int padL = 0, size_t bm = 0) {
  const cplx<T> *mat = mats + g.mat_off + bm * g.mat_bstride;
  switch (g.kind) {
    case TQB_GATE_DENSE:
      mux_sweep<T>(tile, 1u << (m - 1), (uint32_t)g.bits[0], 0, false, 0, 1u << g.bits[0], mat, tid, nthreads, padL);
      break;
    case TQB_GATE_DIAG: gate_diag<T>(tile, m, gbase, g, mat, tid, nthreads, padL); break;
    case TQB_GATE_MUX: gate_mux<T>(tile, m, gbase, g, mat, tid, nthreads, padL); break;
    case TQB_GATE_CHAIN:
      if (g.off_a >= 4u) {
        if (WITH_ROT) {
          const RotDesc rd = rot_decode<T>(g, m, padL);
          chain_rot_dispatch<T>(tile, gbase, rd, mats, tid, nthreads, NoSync(), bm);
        }
      } else if (g.k == 2) {
        gate_chain<T, 2>(tile, m, gbase, g, mat, tid, nthreads, padL);
      } else if (g.k == 3) {
        gate_chain<T, 3>(tile, m, gbase, g, mat, tid, nthreads, padL);
      }
      break;
    default: break;
  }
}

}  // namespace tqb
