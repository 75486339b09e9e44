// tqb_emu.cpp -- TEST TOOLING ONLY (never loaded by the tyxonq_b200 package).
//
// Host emulation of tile_pass_kernel (tyxonq_b200/csrc/tqb_tile.cu): it runs the very same
// __host__ __device__ phase functions of tqb_core.cuh from nested loops (tile -> phase ->
// thread), so the tile/index logic and the planner output can be checked against the numpy
// oracle in the CPU-only test tier.  Built by tests/emu/build_emu.py with g++.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../tyxonq_b200/csrc/tqb_core.cuh"

using namespace tqb;

template <typename T, int V>
static void run_pass(cplx<T> *state, const TileGeom &geo, long long batch, const tqb_gate *gates, int n_gates,
                     const cplx<T> *mats, int nthreads, int max_dense_k) {
  std::vector<cplx<T>> tile((size_t)1 << geo.m);
  std::vector<uint64_t> roff((size_t)1 << geo.h);
  for (uint32_t j = 0; j < (1u << geo.h); ++j) roff[j] = run_offset(geo, j);
  const int tb = geo.n - geo.m;
  const unsigned long long total = (unsigned long long)batch << tb;
  for (unsigned long long tt = 0; tt < total; ++tt) {
    const unsigned long long b = tt >> tb;
    const uint64_t t = tt & ((1ull << tb) - 1ull);
    const uint64_t base = tile_base(geo, t);
    cplx<T> *sb = state + (b << geo.n);
    for (int tid = 0; tid < nthreads; ++tid) tile_load<T, V>(tile.data(), sb, geo, roff.data(), base, tid, nthreads);
    for (int gi = 0; gi < n_gates; ++gi) {
      for (int tid = 0; tid < nthreads; ++tid) {
        if (max_dense_k > 2)
          tile_apply_gate<T, 4>(tile.data(), geo, roff.data(), geo.global_base | base, gates[gi], mats, (size_t)b, tid, nthreads);
        else
          tile_apply_gate<T, 2>(tile.data(), geo, roff.data(), geo.global_base | base, gates[gi], mats, (size_t)b, tid, nthreads);
      }
    }
    for (int tid = 0; tid < nthreads; ++tid) tile_store<T, V>(tile.data(), sb, geo, roff.data(), base, tid, nthreads);
  }
}

// Lean path (tile_pass_lean_kernel): padded tile layout, lean dispatch, chains decoded through RotDesc.
template <typename T>
static void run_pass_lean(cplx<T> *state, const TileGeom &geo, long long batch, const tqb_gate *gates, int n_gates,
                          const cplx<T> *mats, int nthreads, bool pad) {
  const int padL = (pad && geo.L >= 1 && geo.L <= 7 && geo.h > 0) ? geo.L : 0;
  const uint32_t nel = 1u << geo.m;
  std::vector<cplx<T>> tile((size_t)pidx<T>(nel - 1, padL) + 1);
  std::vector<uint64_t> roff((size_t)1 << geo.h);
  for (uint32_t j = 0; j < (1u << geo.h); ++j) roff[j] = run_offset(geo, j);
  const int tb = geo.n - geo.m;
  const unsigned long long total = (unsigned long long)batch << tb;
  for (unsigned long long tt = 0; tt < total; ++tt) {
    const unsigned long long b = tt >> tb;
    const uint64_t base = tile_base(geo, tt & ((1ull << tb) - 1ull));
    cplx<T> *sb = state + (b << geo.n);
    for (uint32_t e = 0; e < nel; ++e) tile[pidx<T>(e, padL)] = sb[local_to_index(geo, roff.data(), base, e)];
    for (int gi = 0; gi < n_gates; ++gi)
      for (int tid = 0; tid < nthreads; ++tid)
        tile_apply_gate_lean<T, true>(tile.data(), geo.m, geo.global_base | base, gates[gi], mats, tid, nthreads, padL, (size_t)b);
    for (uint32_t e = 0; e < nel; ++e) sb[local_to_index(geo, roff.data(), base, e)] = tile[pidx<T>(e, padL)];
  }
}

extern "C" int tqb_emu_run_passes(void *state, int n, long long batch, int dtype, unsigned long long global_base,
                                  const tqb_pass *passes, int n_passes, const tqb_gate *gates, const void *mats,
                                  int threads) {
  for (int p = 0; p < n_passes; ++p) {
    const tqb_pass &ps = passes[p];
    TileGeom geo;
    geo.n = n; geo.m = ps.m; geo.L = ps.L; geo.h = ps.m - ps.L; geo.global_base = global_base;
    geo.mat_begin = 0; geo.mat_count = 0;
    if (geo.m > n || geo.L > geo.m || geo.h > TQB_MAX_TILE_HIGH) return -1;
    int prev = ps.L - 1;
    for (int i = 0; i < TQB_MAX_TILE_HIGH; ++i) {
      geo.hb[i] = i < geo.h ? ps.hb[i] : 0;
      if (i < geo.h) {
        if (!(ps.hb[i] > prev && ps.hb[i] < n)) return -2;
        prev = ps.hb[i];
      }
    }
    const tqb_gate *g = gates + ps.gate_begin;
    if (ps.max_dense_k < 0) {   // lean-eligible pass (the planner's flag), as tqb_run_passes would run it
      if (dtype == TQB_C128) run_pass_lean<double>((cplx<double> *)state, geo, batch, g, ps.n_gates, (const cplx<double> *)mats, threads > 128 ? 128 : threads, ps.max_dense_k == -2);
      else run_pass_lean<float>((cplx<float> *)state, geo, batch, g, ps.n_gates, (const cplx<float> *)mats, threads > 128 ? 128 : threads, ps.max_dense_k == -2);
      continue;
    }
    if (dtype == TQB_C128)
      run_pass<double, 1>((cplx<double> *)state, geo, batch, g, ps.n_gates, (const cplx<double> *)mats, threads, ps.max_dense_k);
    else if (ps.L >= 1)
      run_pass<float, 2>((cplx<float> *)state, geo, batch, g, ps.n_gates, (const cplx<float> *)mats, threads, ps.max_dense_k);
    else
      run_pass<float, 1>((cplx<float> *)state, geo, batch, g, ps.n_gates, (const cplx<float> *)mats, threads, ps.max_dense_k);
  }
  return 0;
}
