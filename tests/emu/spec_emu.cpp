// spec_emu.cpp -- TEST TOOLING ONLY (never loaded by the tyxonq_b200 package).
//
// Host emulation of one SPECIALISED pass kernel (tyxonq_b200/csrc/tqb_spec.cuh): the generated constants of the pass
// (-include <header>) plus the very same per-thread gate code the NVRTC kernel runs, driven from loops.  The loops
// follow the synchronisation the generator chose, in the most adversarial legal order: gates separated by "no sync"
// run thread by thread, gates separated by __syncwarp() warp by warp, and only a consumer barrier (or the start of the
// tile) makes all threads finish a gate before the next one starts -- so a wrong warp-privacy claim of the generator
// shows up as a wrong state.  Built per pass shape by tests/emu/spec_emu.py with g++.
#define TQB_SPEC_EMU
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../tyxonq_b200/csrc/tqb_spec.cuh"

using namespace tqbs;

typedef void (*gate_fn)(char *, const amp *, unsigned, unsigned);
struct NoSync {
  void operator()() const {}
};
static amp g_v[D];   // a thread's register block: lives across gates that are fused in registers (sync == -1)
template <int GI>
static void run_gate(char *tile, const amp *sm, unsigned tid, unsigned extv) {
  if (!fuse_in(GI))
    for (int s = 0; s < D; ++s) g_v[s] = amp{(T)1e300, (T)-1e300};   // poison: a gate that is not fused must load
  if constexpr (LOOPED != 0) apply_gate_looped<GI>(tile, sm, tid, extv, NoSync());
  else apply_gate<GI>(tile, sm, tid, extv, g_v, NoSync());
}

extern "C" int spec_emu_run(void *state_v, int n, long long batch, unsigned long long global_base, const signed char *hb,
                            const signed char *ext, const void *mats_v) {
  amp *state = reinterpret_cast<amp *>(state_v);
  const amp *mats_all = reinterpret_cast<const amp *>(mats_v);
  std::vector<amp> staged(MAT_COUNT + 1);
  gate_fn fn[NG];
  static_for<NG>([&](auto gc) {
    constexpr int GI = decltype(gc)::value;
    fn[GI] = &run_gate<GI>;
  });
  const unsigned nel = 1u << M;
  std::vector<unsigned char> tile((size_t)pbyte(nel - 1) + ES + 64);
  const int tbits = n - M;
  const unsigned long long total = (unsigned long long)batch << tbits;
  for (unsigned long long tt = 0; tt < total; ++tt) {
    const unsigned long long bm = tt >> tbits;
    unsigned long long x = (tt & ((1ull << tbits) - 1ull)) << L;
    for (int j = 0; j < H; ++j) {
      const unsigned p = (unsigned)hb[j];
      x = ((x >> p) << (p + 1u)) | (x & ((1ull << p) - 1ull));
    }
    amp *sb = state + (bm << n);
    auto gidx = [&](unsigned e) -> unsigned long long {
      unsigned long long o = x | (e & ((1u << L) - 1u));
      for (int j = 0; j < H; ++j) o |= (unsigned long long)((e >> (L + j)) & 1u) << hb[j];
      return o;
    };
    for (unsigned e = 0; e < nel; ++e) memcpy(tile.data() + pbyte(e), &sb[gidx(e)], sizeof(amp));
    // the kernel's staged copy of the matrices: gate by gate, the data of batch member bm
    for (int gi = 0; gi < NG; ++gi)
      for (int i = 0; i < G[gi].mlen; ++i) staged[G[gi].mat + i] = mats_all[G[gi].msrc + bm * (unsigned long long)G[gi].mbs + i];
    const amp *mats = staged.data();
    unsigned extv = 0;
    for (int j = 0; j < NEXT; ++j) extv |= (unsigned)(((global_base | x) >> ext[j]) & 1ull) << j;
    // segments: [lo, hi) with G[lo].sync in {0, 2}; runs inside a segment start at sync == 1
    for (int lo = 0; lo < NG;) {
      int hi = lo + 1;
      while (hi < NG && G[hi].sync != 2 && G[hi].sync != 0) ++hi;
      for (unsigned w = 0; w < CT / 32; ++w) {
        for (int rlo = lo; rlo < hi;) {
          int rhi = rlo + 1;
          while (rhi < hi && G[rhi].sync == -1) ++rhi;
          for (unsigned lane = 0; lane < 32; ++lane)
            for (int gi = rlo; gi < rhi; ++gi) fn[gi](reinterpret_cast<char *>(tile.data()), mats, w * 32 + lane, extv);
          rlo = rhi;
        }
      }
      lo = hi;
    }
    for (unsigned e = 0; e < nel; ++e) memcpy(&sb[gidx(e)], tile.data() + pbyte(e), sizeof(amp));
  }
  return 0;
}
