// tqb_jit.cu -- per-pass-shape specialisation of the fused gate pass.
//
// tqb_run_passes2() hands every lean-eligible pass (only 1-qubit-layer gates, matrices staged) to spec_try_launch():
//   1. spec_plan()   turns the pass + its gate descriptors into a compile-time description: for every gate the
//                    register bits (targets + fillers), the thread-bit map, the kind of synchronisation in front of
//                    it, the tile layout (plain / padded) -- chosen by exhaustive search over the few candidates with
//                    an exact shared-memory bank-conflict count;
//   2. spec_header() prints that description as C++ constants (the cache key);
//   3. the constants + the hand-written kernel text (tqb_spec.cuh, embedded at build time) are compiled by NVRTC for
//      sm_100a, the cubin is loaded with cudaLibraryLoadData and cached in memory and on disk;
//   4. the kernel is launched like tile_pass_lean_kernel (persistent grid of resident CTAs).
// Outside-the-tile bit positions, the high tile bits hb[] and all matrix VALUES are run-time parameters, so e.g. the
// four middle passes of every layer period of a hardware-efficient ansatz share one kernel.
//
// Modes (tqb_set_jit): 0 = off (generic kernels only), 1 = asynchronous (default: shapes compile on background
// threads while the generic lean kernel runs them; tqb_jit_wait() drains the queue), 2 = synchronous (compile on first
// use, errors are returned to the caller).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <chrono>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "tqb_host.h"

namespace tqb {

static const char *const kSpecTemplate =
#include "tqb_spec_src.inc"
    ;

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
enum { K_DENSE1 = 0, K_DIAG = 1, K_MUX = 4, K_CHAIN = 5, K_ROT = 6 };

struct SpecGate {
  int kind = 0, R = 0, type = 0, muxed = 0, unit_p = 0, E = 0, ctrl = -1, mat = 0, sync = 0;
  int msrc = 0, mlen = 0, mbs = 0;   // matrices: offset of member 0's data from the pass's first matrix, length, per-member stride
  int inv = -1;           // rotation-form chains: bit i-1 = layer i runs in the c form, bit 3 = layer 0 runs in the t form with its factor
                          // in the table (tqb_gate.off_b bits 8..10, 12); -1 = decided at run time (bit 11 clear)
  int xb[2] = {-1, -1};
  int dbits[6] = {-1, -1, -1, -1, -1, -1};
  int rb[5] = {-1, -1, -1, -1, -1};
  int tb[7] = {-1, -1, -1, -1, -1, -1, -1};
  uint32_t targets = 0;   // mask of target tile bits
  uint32_t reserved = 0;  // tile bits that must stay thread bits (control, extra table bits)
};

struct SpecPlan {
  int dtype = 0, m = 0, L = 0, padL = 0, next = 0, mat_count = 0, rbits = 0;
  int batched = 0;  // some gate carries one matrix set per batch member (tqb_gate.mat_bstride)
  int looped = 0;   // code shape: 0 = one unrolled copy of the code per gate, 1 = one body per gate class (tqb_spec.cuh)
  int ext[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<SpecGate> g;
};

static inline int popc(uint32_t v) { return __builtin_popcount(v); }

// exact bank-conflict count of the gate's tile accesses for warp 0: sum over the sampled register indices and the
// phases of a warp-wide access of the largest number of lanes that need DIFFERENT 128-byte rows of the same bank group
// unit: bytes of one access -- es, or 16 for complex64 gates that keep tile bit 0 in registers (tqb_spec.cuh pair_reg)
static int conflict_cost(const SpecGate &s, int rbits, int es, int padL, int unit) {
  auto pbyte = [&](uint32_t e) -> uint32_t { return e * (uint32_t)es + (padL ? ((e >> padL) << 4) : 0u); };
  const int lanes_per_phase = 128 / unit;
  const int D = 1 << rbits;
  int cost = 0;
  const int samples[3] = {0, D - 1, D / 2 - 1 > 0 ? D / 2 - 1 : 0};
  for (int si = 0; si < 3; ++si) {
    uint32_t so = 0;
    for (int i = 0; i < rbits; ++i)
      if ((samples[si] >> i) & 1) so |= 1u << s.rb[i];
    if (unit > es) so &= ~1u;   // the pair's first element
    for (int ph = 0; ph < 32 / lanes_per_phase; ++ph) {
      uint32_t rows[16][16];
      int cnt[16];
      for (int c = 0; c < 16; ++c) cnt[c] = 0;
      for (int l = 0; l < lanes_per_phase; ++l) {
        const uint32_t tid = (uint32_t)(ph * lanes_per_phase + l);
        uint32_t e = so;
        for (int j = 0; j < 5; ++j) e |= ((tid >> j) & 1u) << s.tb[j];
        const uint32_t a = pbyte(e);
        const int chunk = (int)((a / (uint32_t)unit) % (uint32_t)lanes_per_phase);
        const uint32_t row = a / 128u;
        bool seen = false;
        for (int k = 0; k < cnt[chunk]; ++k) seen = seen || rows[chunk][k] == row;
        if (!seen) rows[chunk][cnt[chunk]++] = row;
      }
      int deg = 1;
      for (int c = 0; c < lanes_per_phase; ++c) deg = std::max(deg, cnt[c]);
      cost += deg;
    }
  }
  return cost;
}

// best register / thread mapping of one gate for warp bits W; returns its cost (conflicts first, then preferences)
static int map_gate(SpecGate &s, int m, int rbits, int es, int padL, uint32_t W) {
  const uint32_t all = (1u << m) - 1u;
  const uint32_t free_bits = all & ~s.targets & ~W;
  const int nf = rbits - popc(s.targets);
  const uint32_t fill_cand = free_bits & ~s.reserved;
  // complex64: tile bit 0 in registers makes every access of the gate a 16-byte one (a forced filler unless a layer
  // targets it); impossible when bit 0 is a warp bit or carries a control / table bit of the gate
  const bool pair_forced = es == 8 && !(s.targets & 1u) && (fill_cand & 1u) && nf >= 1;
  const bool pair = es == 8 && ((s.targets & 1u) || pair_forced);
  const int unit = pair ? 16 : es;
  std::vector<int> cand;
  for (int b = 0; b < m; ++b)
    if ((fill_cand >> b) & 1u) cand.push_back(b);
  int best = 1 << 30;
  SpecGate best_s = s;
  const int nc = (int)cand.size();
  // enumerate filler subsets (nf of nc candidates) by bitmask, at most 24 of them (high bits first)
  int tried = 0;
  for (uint32_t sub = (1u << nc); sub-- > 0;) {
    if (popc(sub) != nf) continue;
    uint32_t fillers = 0;
    for (int i = 0; i < nc; ++i)
      if ((sub >> i) & 1u) fillers |= 1u << cand[i];
    if (pair_forced && !(fillers & 1u)) continue;
    if (++tried > 24) break;
    SpecGate t = s;
    {
      int r = s.R;
      if (s.kind == K_DIAG) r = 0;
      for (int b = m - 1; b >= 0; --b)
        if ((fillers >> b) & 1u) t.rb[r++] = b;
    }
    const uint32_t lane_bits = free_bits & ~fillers;
    std::vector<int> lb;
    for (int b = 0; b < m; ++b)
      if ((lane_bits >> b) & 1u) lb.push_back(b);
    if ((int)lb.size() != 5) continue;
    // which lane bits take the low lane positions: enumerate orderings by choosing the subset for lanes 0..nl-1
    const int nl = unit == 16 ? 3 : 4;
    for (uint32_t lo = 0; lo < 32u; ++lo) {
      if (popc(lo) != nl) continue;
      int j = 0;
      for (int i = 0; i < 5; ++i)
        if ((lo >> i) & 1u) t.tb[j++] = lb[i];
      for (int i = 0; i < 5; ++i)
        if (!((lo >> i) & 1u)) t.tb[j++] = lb[i];
      int w = 5;
      for (int b = 0; b < m; ++b)
        if ((W >> b) & 1u) t.tb[w++] = b;
      // (per byte moved: an 8-byte access has half the phases of a 16-byte one)
      int c = conflict_cost(t, rbits, es, padL, unit) * (unit > es || es == 16 ? 1000 : 2000);
      // preferences: control / extra bits outside the low lanes (table loads stay one broadcast per phase),
      // fillers on high bits
      for (int i = 0; i < nl; ++i)
        if ((s.reserved >> t.tb[i]) & 1u) c += 10;
      for (int b = 0; b < m; ++b)
        if ((fillers >> b) & 1u) c += (m - 1 - b);
      if (c < best) {
        best = c;
        best_s = t;
      }
    }
  }
  s = best_s;
  return best;
}

// code shape of the generated kernels: 0 = unrolled (one copy of the code per gate), 1 = looped (one body per gate class),
// 2 = by dtype: complex64 looped (its unrolled kernels are instruction-fetch bound), complex128 unrolled
static std::atomic<int> g_spec_loop{2};

// cheap half: the pass in normalised form (outside-the-tile bits as slots, matrix offsets relative to the pass)
static bool spec_parse(const tqb_pass &ps, const tqb_gate *gh, int dtype, SpecPlan &P, std::string &why) {
  const int m = ps.m, L = ps.L, h = ps.m - ps.L;
  const int es = dtype == TQB_C128 ? 16 : 8;
  P = SpecPlan();
  P.dtype = dtype;
  P.m = m;
  P.L = L;
  P.rbits = m - 7;
  P.mat_count = 0;   // the staged (per-member) length: filled in gate by gate below
  {
    const int lm = g_spec_loop.load();
    P.looped = lm == 2 ? (dtype == TQB_C128 ? 0 : 1) : (lm ? 1 : 0);
  }
  if (dtype == TQB_C128 ? P.rbits != 4 : (P.rbits != 4 && P.rbits != 5)) return why = "tile size", false;
  if (ps.max_dense_k >= 0) return why = "not a lean pass", false;
  if (((size_t)es << L) < 128 || h < 1 || h > 6) return why = "run length", false;
  if (ps.n_gates < 1 || ps.n_gates > 64) return why = "gate count", false;
  auto code = [&](int8_t b, bool &ok) -> int {
    const unsigned u = (uint8_t)b;
    if (u == 127u) return -1;
    if (u >= 64u) {
      const int pos = (int)(u - 64u);
      for (int j = 0; j < P.next; ++j)
        if (P.ext[j] == pos) return 64 + j;
      if (P.next >= 8) {
        ok = false;
        return -1;
      }
      P.ext[P.next] = pos;
      return 64 + P.next++;
    }
    if ((int)u >= m) ok = false;
    return (int)u;
  };
  for (int gi = 0; gi < ps.n_gates; ++gi) {
    const tqb_gate &q = gh[ps.gate_begin + gi];
    SpecGate s;
    bool ok = true;
    s.msrc = (int)q.mat_off - ps.mat_begin;
    s.mbs = (int)q.mat_bstride;
    s.mat = P.mat_count;   // where the gate's (member's) data sit in the staged copy
    if (s.msrc < 0) return why = "matrix range", false;
    if (s.mbs != 0) P.batched = 1;
    auto target = [&](int i, int8_t b) {
      if (b < 0 || b >= m || ((s.targets >> b) & 1u)) ok = false;
      else {
        s.rb[i] = b;
        s.targets |= 1u << b;
      }
    };
    auto reserve = [&](int c) {
      if (c >= 0 && c < 64) {
        if ((s.targets >> c) & 1u) ok = false;
        s.reserved |= 1u << c;
      }
    };
    switch (q.kind) {
      case TQB_GATE_DENSE:
        if (q.k != 1) return why = "dense k > 1", false;
        s.kind = K_DENSE1;
        s.R = 1;
        target(0, q.bits[0]);
        break;
      case TQB_GATE_MUX:
        s.kind = K_MUX;
        s.R = 1;
        target(0, q.bits[0]);
        s.ctrl = code(q.bits[1], ok);
        if (s.ctrl < 0) ok = false;
        reserve(s.ctrl);
        break;
      case TQB_GATE_CHAIN: {
        s.R = q.k;
        if (q.k < 2 || q.k > 4 || q.k > P.rbits) return why = "chain length", false;
        for (int i = 0; i < q.k; ++i) target(i, q.bits[i]);
        s.ctrl = code(q.bits[q.k], ok);
        reserve(s.ctrl);
        if (q.off_a >= 4u) {
          s.kind = K_ROT;
          s.type = (int)((q.off_a >> 1) & 1u);
          s.muxed = (int)(q.off_a & 1u);
          s.E = (int)(q.off_b & 3u);
          s.unit_p = (q.off_b & 128u) ? 1 : 0;
          s.inv = (q.off_b & 2048u) ? (int)(((q.off_b >> 8) & 7u) | ((q.off_b & 4096u) ? 8u : 0u)) : -1;
          if ((q.off_b & 4096u) && !(q.off_b & 2048u)) return why = "chain form bits", false;
          if (s.E > 2 || q.k + 1 + s.E > TQB_MAX_GATE_BITS) return why = "extra bits", false;
          for (int j = 0; j < s.E; ++j) {
            s.xb[j] = code(q.bits[q.k + 1 + j], ok);
            if (s.xb[j] < 0) ok = false;
            reserve(s.xb[j]);
          }
        } else if (q.off_a == 0u && q.k <= 3) {
          s.kind = K_CHAIN;
        } else {
          return why = "chain form", false;
        }
        break;
      }
      case TQB_GATE_DIAG:
        s.kind = K_DIAG;
        s.R = q.k;
        if (q.k < 1 || q.k > 6) return why = "diag size", false;
        for (int j = 0; j < q.k; ++j) {
          s.dbits[j] = code(q.bits[j], ok);
          if (s.dbits[j] < 0) ok = false;
        }
        break;
      default:
        return why = "gate kind", false;
    }
    if (!ok) return why = "bad gate bits", false;
    switch (s.kind) {   // length of the gate's data (complex elements)
      case K_DENSE1: s.mlen = 4; break;
      case K_MUX: s.mlen = 8; break;
      case K_CHAIN: s.mlen = 8 * s.R; break;
      case K_DIAG: s.mlen = 1 << s.R; break;
      default: s.mlen = ((1 << (s.R + s.E)) << (s.ctrl >= 0 ? 1 : 0)) + s.R; break;   // K_ROT: table(s) + one coefficient per layer
    }
    if (s.mbs != 0 && s.mbs != s.mlen) return why = "per-member stride", false;
    P.mat_count += s.mlen;
    P.g.push_back(s);
  }
  if (P.mat_count <= 0 || P.mat_count > 2048) return why = "matrices not staged", false;
  return true;
}

// The shape of a parsed pass: everything the generated constants depend on (NOT the outside-the-tile bit positions).
static std::string shape_key(const SpecPlan &P) {
  std::string k;
  auto put = [&](int v) { k.append(reinterpret_cast<const char *>(&v), sizeof v); };
  put(P.dtype); put(P.m); put(P.L); put(P.mat_count); put(P.next); put((int)P.g.size()); put(P.looped);
  for (const SpecGate &s : P.g) {
    put(s.kind); put(s.R); put(s.type); put(s.muxed); put(s.unit_p); put(s.E); put(s.ctrl); put(s.mat); put(s.inv); put(s.msrc); put(s.mbs);
    put(s.xb[0]); put(s.xb[1]);
    for (int j = 0; j < 6; ++j) put(s.dbits[j]);
    for (int j = 0; j < 5; ++j) put(s.kind == K_DIAG ? -1 : (j < s.R ? s.rb[j] : -1));
  }
  return k;
}

// expensive half: segments, warp bits, register / lane maps, tile layout (a search with exact bank-conflict counts)
static bool spec_search(SpecPlan &P, std::string &why) {
  const int m = P.m, L = P.L, h = P.m - P.L;
  const int es = P.dtype == TQB_C128 ? 16 : 8;
  const int ng = (int)P.g.size();
  const uint32_t all = (1u << m) - 1u;

  // Segments and warp bits.  A segment is a run of gates that leaves >= 2 tile bits untargeted: those become the warp
  // bits W of the segment (gates inside it are separated by __syncwarp() only), segments are separated by the
  // consumers' barrier.  cost[g][W] = bank conflicts of gate g's best mapping under W (memoised: map_gate is a search);
  // a dynamic program over the segment boundaries minimises conflicts + BARRIER per boundary, for the plain and the
  // padded tile layout.
  const int BARRIER = 4000;
  std::vector<uint32_t> pairs;
  for (int a = 0; a < m; ++a)
    for (int b = a + 1; b < m; ++b) pairs.push_back((1u << a) | (1u << b));
  const int np = (int)pairs.size();
  int best_total = 1 << 30;
  std::vector<SpecGate> best_g;
  int best_pad = 0;
  const bool can_pad = L >= 1 && L <= 7 && h > 0;
  for (int pad = 0; pad < (can_pad ? 2 : 1); ++pad) {
    const int padL = pad ? L : 0;
    // memo over (targets, reserved, kind-independent): gates of the same shape share the search
    std::unordered_map<uint64_t, std::pair<int, SpecGate>> memo;
    auto gate_cost = [&](int gi, int pi) -> std::pair<int, SpecGate> {
      const SpecGate &s0 = P.g[gi];
      const uint32_t W = pairs[pi];
      const uint64_t key = ((uint64_t)s0.targets << 40) ^ ((uint64_t)s0.reserved << 16) ^ (uint64_t)pi ^ ((uint64_t)s0.R << 60);
      auto it = memo.find(key);
      if (it != memo.end()) {
        // same shape: copy the mapping, keep the gate's own fields
        SpecGate t = s0;
        const SpecGate &src = it->second.second;
        // (targets are listed in the gate's own layer order; fillers follow)
        for (int k = s0.kind == K_DIAG ? 0 : s0.R; k < 5; ++k) t.rb[k] = src.rb[k];
        for (int k = 0; k < 7; ++k) t.tb[k] = src.tb[k];
        return {it->second.first, t};
      }
      SpecGate t = s0;
      int c;
      if (popc(all & ~t.targets & ~W & ~t.reserved) < P.rbits - popc(t.targets)) c = 1 << 28;
      else c = map_gate(t, m, P.rbits, es, padL, W);
      memo[key] = {c, t};
      return {c, t};
    };
    // dp[i] = best cost of gates i.. ; choice[i] = (j, pair)
    std::vector<int> dp(ng + 1, 1 << 30), nxt(ng + 1, -1), pick(ng + 1, -1);
    dp[ng] = 0;
    for (int i = ng - 1; i >= 0; --i) {
      uint32_t C = all;
      for (int j = i + 1; j <= ng; ++j) {
        C &= ~P.g[j - 1].targets;
        if (popc(C) < 2) break;
        if (dp[j] >= (1 << 28)) continue;
        for (int pi = 0; pi < np; ++pi) {
          if ((pairs[pi] & C) != pairs[pi]) continue;
          int c = 0;
          for (int g = i; g < j && c < (1 << 28); ++g)
            if (P.g[g].kind != K_DIAG) c += gate_cost(g, pi).first;
          if (c >= (1 << 28)) continue;
          c += dp[j] + (j < ng ? BARRIER : 0);
          // tie: prefer longer segments, then higher warp bits
          if (c < dp[i] || (c == dp[i] && (j > nxt[i] || (j == nxt[i] && pairs[pi] > pairs[pick[i]])))) {
            dp[i] = c;
            nxt[i] = j;
            pick[i] = pi;
          }
        }
      }
    }
    if (dp[0] >= (1 << 28)) continue;
    std::vector<SpecGate> cur = P.g;
    for (int i = 0; i < ng;) {
      const int j = nxt[i], pi = pick[i];
      for (int g = i; g < j; ++g)
        if (P.g[g].kind != K_DIAG) cur[g] = gate_cost(g, pi).second;
        else {   // remember the segment's warp bits on diagonal gates too (used when the whole segment is diagonal)
          int w = 5;
          for (int b = 0; b < m; ++b)
            if ((pairs[pi] >> b) & 1u) cur[g].tb[w++] = b;
        }
      i = j;
    }
    if (dp[0] < best_total) {
      best_total = dp[0];
      best_g = cur;
      best_pad = padL;
    }
  }
  if (best_total == (1 << 30)) return why = "no feasible warp bits", false;
  P.g = best_g;
  P.padL = best_pad;
  // DIAG gates ride on a neighbour's mapping (same thread, same amplitudes: no synchronisation in between)
  for (int i = 0; i < ng; ++i) {
    if (P.g[i].kind != K_DIAG) continue;
    int src = -1;
    for (int j = i - 1; j >= 0 && src < 0; --j)
      if (P.g[j].kind != K_DIAG) src = j;
    for (int j = i + 1; j < ng && src < 0; ++j)
      if (P.g[j].kind != K_DIAG) src = j;
    if (src >= 0) {
      for (int k = 0; k < 5; ++k) P.g[i].rb[k] = P.g[src].rb[k];
      for (int k = 0; k < 7; ++k) P.g[i].tb[k] = P.g[src].tb[k];
    } else {  // a pass of diagonal gates only: registers on the top bits, lanes on the low bits
      for (int k = 0; k < P.rbits; ++k) P.g[i].rb[k] = m - 1 - k;
      for (int k = 0; k < 7; ++k) P.g[i].tb[k] = k;
    }
  }
  for (int i = 0; i < ng; ++i) {
    SpecGate &s = P.g[i];
    if (i == 0) {
      s.sync = 0;
      continue;
    }
    const SpecGate &p = P.g[i - 1];
    bool same = true;
    uint32_t ra = 0, rb = 0;
    for (int k = 0; k < P.rbits; ++k) {
      ra |= 1u << s.rb[k];
      rb |= 1u << p.rb[k];
    }
    same = ra == rb;
    for (int k = 0; k < 7; ++k) same = same && s.tb[k] == p.tb[k];
    const bool same_w = s.tb[5] == p.tb[5] && s.tb[6] == p.tb[6];
    s.sync = same ? -1 : (same_w ? 1 : 2);
  }
  // sanity: register bits and thread bits partition the tile bits
  for (const SpecGate &s : P.g) {
    uint32_t seen = 0;
    for (int k = 0; k < P.rbits; ++k) {
      if (s.rb[k] < 0 || s.rb[k] >= m || ((seen >> s.rb[k]) & 1u)) return why = "internal: register bits", false;
      seen |= 1u << s.rb[k];
    }
    for (int k = 0; k < 7; ++k) {
      if (s.tb[k] < 0 || s.tb[k] >= m || ((seen >> s.tb[k]) & 1u)) return why = "internal: thread bits", false;
      seen |= 1u << s.tb[k];
    }
    if (seen != all) return why = "internal: partition", false;
    for (int k = 0; k < P.rbits; ++k)
      if ((s.reserved >> s.rb[k]) & 1u) return why = "internal: reserved bit in registers", false;
  }
  return true;
}

static bool spec_plan(const tqb_pass &ps, const tqb_gate *gh, int dtype, SpecPlan &P, std::string &why) {
  return spec_parse(ps, gh, dtype, P, why) && spec_search(P, why);
}

static std::string spec_header(const SpecPlan &P) {
  std::string o;
  char buf[512];
  o += "namespace tqbs {\n";
  o += P.dtype == TQB_C128 ? "typedef double T;\n" : "typedef float T;\n";
  snprintf(buf, sizeof buf, "constexpr int M = %d, L = %d, PADL = %d, NG = %d, NEXT = %d, MAT_COUNT = %d, RBITS = %d, LOOPED = %d, BATCHED = %d;\n",
           P.m, P.L, P.padL, (int)P.g.size(), P.next, P.mat_count, P.rbits, P.looped, P.batched);
  o += buf;
  o += "struct GateC { int kind, R, type, muxed, unit_p, E, ctrl, mat, sync; int xb[2]; int dbits[6]; int rb[5]; int tb[7]; int inv, msrc, mlen, mbs; };\n";
  o += "constexpr GateC G[NG] = {\n";
  for (const SpecGate &s : P.g) {
    snprintf(buf, sizeof buf,
             "  {%d, %d, %d, %d, %d, %d, %d, %d, %d, {%d, %d}, {%d, %d, %d, %d, %d, %d}, {%d, %d, %d, %d, %d}, {%d, %d, %d, %d, %d, %d, %d}, %d, %d, %d, %d},\n",
             s.kind, s.R, s.type, s.muxed, s.unit_p, s.E, s.ctrl, s.mat, s.sync, s.xb[0], s.xb[1], s.dbits[0], s.dbits[1],
             s.dbits[2], s.dbits[3], s.dbits[4], s.dbits[5], s.rb[0], s.rb[1], s.rb[2], s.rb[3], s.rb[4], s.tb[0], s.tb[1],
             s.tb[2], s.tb[3], s.tb[4], s.tb[5], s.tb[6], s.inv, s.msrc, s.mlen, s.mbs);
    o += buf;
  }
  o += "};\n}\n";
  return o;
}

// ------------------------------------------------------------------------------------------
// NVRTC (loaded with dlopen: the library itself links only libcudart)
// ------------------------------------------------------------------------------------------
static void jit_shutdown_hook();
struct Nvrtc {
  void *h = nullptr;
  int (*CreateProgram)(void **, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
  int (*CompileProgram)(void *, int, const char *const *) = nullptr;
  int (*GetCUBINSize)(void *, size_t *) = nullptr;
  int (*GetCUBIN)(void *, char *) = nullptr;
  int (*GetProgramLogSize)(void *, size_t *) = nullptr;
  int (*GetProgramLog)(void *, char *) = nullptr;
  int (*DestroyProgram)(void **) = nullptr;
  int (*Version)(int *, int *) = nullptr;
  std::string err;
};

static Nvrtc *nvrtc() {
  static Nvrtc N;
  static std::once_flag once;
  std::call_once(once, [] {
    std::vector<std::string> cands;
    if (const char *e = getenv("TQB_NVRTC")) cands.push_back(e);
    cands.push_back("libnvrtc.so.12");
    cands.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
    cands.push_back("libnvrtc.so");
    for (const std::string &c : cands) {
      N.h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
      if (N.h) break;
    }
    if (!N.h) {
      N.err = "libnvrtc.so.12 not found (set TQB_NVRTC)";
      return;
    }
    std::atexit(jit_shutdown_hook);   // (after the dlopen: runs before NVRTC's own exit-time destructors)
#define TQB_SYM(field, name)                                             \
  N.field = reinterpret_cast<decltype(N.field)>(dlsym(N.h, name));       \
  if (!N.field) N.err = std::string("nvrtc symbol missing: ") + name;
    TQB_SYM(CreateProgram, "nvrtcCreateProgram")
    TQB_SYM(CompileProgram, "nvrtcCompileProgram")
    TQB_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    TQB_SYM(GetCUBIN, "nvrtcGetCUBIN")
    TQB_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    TQB_SYM(GetProgramLog, "nvrtcGetProgramLog")
    TQB_SYM(DestroyProgram, "nvrtcDestroyProgram")
    TQB_SYM(Version, "nvrtcVersion")
#undef TQB_SYM
  });
  return &N;
}

// header + template -> cubin (sm_100a).  Thread safe.
static bool spec_compile(const std::string &header, std::vector<char> &cubin, std::string &log) {
  Nvrtc *N = nvrtc();
  if (!N->h || !N->err.empty()) {
    log = N->err;
    return false;
  }
  const std::string src = header + kSpecTemplate;
  void *prog = nullptr;
  if (N->CreateProgram(&prog, src.c_str(), "tqb_spec_pass.cu", 0, nullptr, nullptr) != 0) {
    log = "nvrtcCreateProgram failed";
    return false;
  }
  const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device"};
  const int rc = N->CompileProgram(prog, 4, opts);
  size_t ls = 0;
  N->GetProgramLogSize(prog, &ls);
  if (ls > 1) {
    log.resize(ls);
    N->GetProgramLog(prog, &log[0]);
  }
  bool ok = rc == 0;
  if (ok) {
    size_t cs = 0;
    ok = N->GetCUBINSize(prog, &cs) == 0 && cs > 0;
    if (ok) {
      cubin.resize(cs);
      ok = N->GetCUBIN(prog, cubin.data()) == 0;
    }
    if (!ok) log += " (no cubin)";
  }
  N->DestroyProgram(&prog);
  return ok;
}

// ------------------------------------------------------------------------------------------
// cache
// ------------------------------------------------------------------------------------------
struct SpecParams {  // must match tqbs::SpecParams in tqb_spec.cuh
  void *state;
  const void *mats;
  uint64_t global_base;
  long long batch;
  int n;
  int dbg;
  signed char hb[16];
  signed char ext[8];
  int use_tensor;
  unsigned seg_mask[5];
  signed char seg_start[8];
};

static std::atomic<int> g_tensor_tma{1};   // tqb_set_jit(512 + v): tensor-map staging on (1, default) / off (0)

// Describe the state as a rank-5 tensor whose dimensions are the maximal runs of tile / non-tile index bits (runs of tile
// bits longer than 7 are split: a box dimension holds at most 256 elements, two per amplitude) and encode the CUtensorMap whose box is one
// tile.  Returns false when the tile does not fit the scheme (more than 5 runs, batch not a power of two, padded layout).
static bool make_tensor_map(void *state, int n, int64_t batch, int dtype, const tqb_pass &ps, CUtensorMap *tm, SpecParams &prm) {
  static PFN_cuTensorMapEncodeTiled encode = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    else
      cudaGetLastError();
  });
  if (!encode || batch < 1 || (batch & (batch - 1))) return false;
  int lb = 0;
  while ((1ll << lb) < batch) ++lb;
  const int N = n + lb;   // index bits of the whole array
  const int h = ps.m - ps.L;
  uint64_t tile_bits = (1ull << ps.L) - 1ull;
  for (int j = 0; j < h; ++j) tile_bits |= 1ull << ps.hb[j];
  struct Seg { int start, len; bool tile; };
  Seg segs[8];
  int ns = 0;
  for (int b = 0; b < N;) {
    const bool t = (tile_bits >> b) & 1ull;
    int e = b;
    while (e < N && (((tile_bits >> e) & 1ull) != 0) == t && (!t || e - b < 7) && (t || e - b < 31)) ++e;
    if (ns >= 5) return false;
    segs[ns++] = {b, e - b, t};
    b = e;
  }
  if (!segs[0].tile) return false;
  const int es = dtype == TQB_C128 ? 16 : 8;
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t box[5], estr[5];
  for (int k = 0; k < 5; ++k) {
    estr[k] = 1;
    if (k < ns) {
      gdim[k] = 1ull << segs[k].len;
      box[k] = segs[k].tile ? (cuuint32_t)(1u << segs[k].len) : 1u;
      if (k > 0) gstride[k - 1] = ((cuuint64_t)es) << segs[k].start;
      prm.seg_start[k] = (signed char)segs[k].start;
      prm.seg_mask[k] = segs[k].tile ? 0u : (unsigned)((1ull << segs[k].len) - 1ull);
    } else {
      gdim[k] = 1;
      box[k] = 1;
      gstride[k - 1] = ((cuuint64_t)es) << N;
      prm.seg_start[k] = 0;
      prm.seg_mask[k] = 0;
    }
  }
  // dimension 0 in units of the amplitude's real components (no 16-byte element type exists)
  gdim[0] *= 2;
  box[0] *= 2;
  if (box[0] > 256) return false;
  const CUresult r = encode(tm, dtype == TQB_C128 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, state, gdim,
                            gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

struct SpecKernel {
  enum State { PENDING, READY, FAILED };
  std::atomic<int> state{PENDING};
  SpecPlan plan;        // parsed by the launching thread, searched by build_kernel
  std::string header;
  std::vector<char> cubin;
  std::string log;
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kern = nullptr;
  bool loaded = false;
  bool configured = false;
  bool rejected = false;   // the search found no mapping: the generic kernel runs the pass (not an error)
  int resident = 0;
  size_t smem = 0;
  int64_t uses = 0;
};

// The synchronisation objects, the cache and the queue are never destroyed: worker threads are detached and may be
// blocked on the condition variable when the process exits (destroying a condition variable with waiters hangs).
static std::mutex &g_jit_mu = *new std::mutex;
static std::condition_variable &g_jit_cv = *new std::condition_variable;
static std::unordered_map<std::string, std::shared_ptr<SpecKernel>> &g_cache = *new std::unordered_map<std::string, std::shared_ptr<SpecKernel>>;   // key: shape_key() of the parsed pass
static std::deque<std::shared_ptr<SpecKernel>> &g_queue = *new std::deque<std::shared_ptr<SpecKernel>>;
static std::vector<std::thread> &g_workers = *new std::vector<std::thread>;
static int g_pending = 0;
static std::atomic<int> g_jit_mode{1};
static std::atomic<int> g_jit_dbg{0};
static std::string &g_cache_dir = *new std::string;
static std::atomic<int64_t> g_spec_launches{0}, g_spec_compiles{0}, g_spec_disk_hits{0};

static uint64_t fnv1a(const std::string &s, uint64_t h = 1469598103934665603ull) {
  for (unsigned char c : s) {
    h ^= c;
    h *= 1099511628211ull;
  }
  return h;
}

static std::string disk_path(const std::string &header) {
  if (g_cache_dir.empty()) return std::string();
  static const uint64_t th = fnv1a(kSpecTemplate);
  char name[64];
  snprintf(name, sizeof name, "/spec_%016llx.cubin", (unsigned long long)fnv1a(header, th));
  return g_cache_dir + name;
}

static void build_kernel(SpecKernel &k) {
  if (k.header.empty()) {
    std::string why;
    if (!spec_search(k.plan, why)) {
      k.log = "not eligible: " + why;
      k.rejected = true;
      k.state.store(SpecKernel::FAILED);
      return;
    }
    k.header = spec_header(k.plan);
  }
  const std::string path = disk_path(k.header);
  if (!path.empty()) {
    if (FILE *f = fopen(path.c_str(), "rb")) {
      fseek(f, 0, SEEK_END);
      const long sz = ftell(f);
      fseek(f, 0, SEEK_SET);
      if (sz > 0) {
        k.cubin.resize((size_t)sz);
        if (fread(k.cubin.data(), 1, (size_t)sz, f) == (size_t)sz) {
          fclose(f);
          g_spec_disk_hits.fetch_add(1);
          k.state.store(SpecKernel::READY);
          return;
        }
      }
      fclose(f);
      k.cubin.clear();
    }
  }
  if (spec_compile(k.header, k.cubin, k.log)) {
    g_spec_compiles.fetch_add(1);
    if (!path.empty()) {
      const std::string tmp = path + ".tmp" + std::to_string((long long)getpid()) + "_" + std::to_string((unsigned long long)fnv1a(k.header) & 0xffff);
      if (FILE *f = fopen(tmp.c_str(), "wb")) {
        const bool w = fwrite(k.cubin.data(), 1, k.cubin.size(), f) == k.cubin.size();
        fclose(f);
        if (w) rename(tmp.c_str(), path.c_str());
        else remove(tmp.c_str());
      }
    }
    k.state.store(SpecKernel::READY);
  } else {
    k.state.store(SpecKernel::FAILED);
  }
}

static bool g_stop = false;      // set at shutdown: workers take no more jobs
static int g_building = 0;       // compilations in flight

static void worker_main() {
  for (;;) {
    std::shared_ptr<SpecKernel> k;
    {
      std::unique_lock<std::mutex> lk(g_jit_mu);
      g_jit_cv.wait(lk, [] { return g_stop || !g_queue.empty(); });
      if (g_stop) return;
      k = g_queue.front();
      g_queue.pop_front();
      ++g_building;
    }
    build_kernel(*k);
    {
      std::lock_guard<std::mutex> lk(g_jit_mu);
      --g_pending;
      --g_building;
    }
    g_jit_cv.notify_all();
  }
}

// Process exit must not pull NVRTC's own static state from under a compilation in flight: stop taking jobs, drop the
// queue and wait (bounded) for the running ones.  Registered with atexit once NVRTC is loaded; also exported.
static void jit_shutdown() {
  std::unique_lock<std::mutex> lk(g_jit_mu);
  g_stop = true;
  g_pending -= (int)g_queue.size();
  g_queue.clear();
  g_jit_cv.notify_all();
  g_jit_cv.wait_for(lk, std::chrono::seconds(60), [] { return g_building == 0; });
}

static void jit_shutdown_hook() { jit_shutdown(); }

// Try to run pass `ps` with its specialised kernel.  *used = false: the caller runs the generic kernel.
int spec_try_launch(void *state, int n, int64_t batch, int dtype, uint64_t global_base, const tqb_pass &ps,
                    const tqb_gate *gates_host, const void *mats_dev, const Workspace &ws, cudaStream_t st, bool *used) {
  *used = false;
  const int mode = g_jit_mode.load();
  if (mode == 0 || !gates_host || n <= ps.m) return 0;
  std::shared_ptr<SpecKernel> k;
  SpecPlan parsed;
  {
    std::string why;
    if (!spec_parse(ps, gates_host, dtype, parsed, why)) return 0;
  }
  const std::string key = shape_key(parsed);
  bool fresh = false;
  {
    std::lock_guard<std::mutex> lk(g_jit_mu);
    auto it = g_cache.find(key);
    if (it == g_cache.end()) {
      k = std::make_shared<SpecKernel>();
      k->plan = parsed;
      g_cache[key] = k;
      fresh = true;
      if (mode == 1 && !g_stop) {
        if (g_workers.empty()) {
          unsigned nt = std::thread::hardware_concurrency();
          nt = nt < 2 ? 1 : (nt > 8 ? 8 : nt / 2 + 1);
          for (unsigned i = 0; i < nt; ++i) {
            g_workers.emplace_back(worker_main);
            g_workers.back().detach();
          }
        }
        g_queue.push_back(k);
        ++g_pending;
      }
    } else {
      k = it->second;
    }
    ++k->uses;
  }
  if (fresh && mode == 1) {
    g_jit_cv.notify_all();
    return 0;
  }
  if (fresh && mode >= 2) build_kernel(*k);
  if (mode >= 2 && k->state.load() == SpecKernel::PENDING) {   // another thread is compiling it
    std::unique_lock<std::mutex> lk(g_jit_mu);
    g_jit_cv.wait_for(lk, std::chrono::seconds(120), [&] { return k->state.load() != SpecKernel::PENDING; });
  }
  const int stt = k->state.load();
  const SpecPlan &plan = k->plan;   // (searched: padL is final; the outside-the-tile positions come from `parsed`)
  if (stt == SpecKernel::FAILED) {
    if (mode >= 2 && !k->rejected) return fail("tqb_run_passes: NVRTC failed for a specialised pass: " + k->log);
    return 0;
  }
  if (stt != SpecKernel::READY) return 0;
  {
    std::lock_guard<std::mutex> lk(g_jit_mu);
    if (!k->loaded) {
      cudaError_t e = cudaLibraryLoadData(&k->lib, k->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
      if (e == cudaSuccess) e = cudaLibraryGetKernel(&k->kern, k->lib, "tqb_spec_pass");
      if (e != cudaSuccess) {
        k->state.store(SpecKernel::FAILED);
        k->log = std::string("cudaLibraryLoadData / GetKernel: ") + cudaGetErrorString(e);
        cudaGetLastError();
        if (mode >= 2) return fail("tqb_run_passes: " + k->log);
        return 0;
      }
      k->loaded = true;
    }
  }
  const int es = dtype == TQB_C128 ? 16 : 8;
  const int h = ps.m - ps.L;
  const size_t run_stride = ((size_t)es << ps.L) + (plan.padL ? 16 : 0);
  const size_t smem = 2 * (run_stride << h) + 64 + (size_t)((parsed.mat_count + 1) & ~1) * es + ((size_t)8 << h);
  if (smem > (size_t)ws.max_smem_optin) return 0;
  const void *fn = reinterpret_cast<const void *>(k->kern);
  if (!k->configured) {
    TQB_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, ws.max_smem_optin));
    TQB_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int resident = 0;
    TQB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, fn, 160, smem));
    if (resident < 1) {
      k->state.store(SpecKernel::FAILED);
      k->log = "specialised kernel does not fit on an SM";
      return 0;
    }
    k->resident = resident;
    k->smem = smem;
    k->configured = true;
  }
  SpecParams prm;
  memset(&prm, 0, sizeof prm);
  prm.state = state;
  prm.mats = static_cast<const char *>(mats_dev) + (size_t)ps.mat_begin * es;
  prm.global_base = global_base;
  prm.batch = (long long)batch;
  prm.n = n;
  prm.dbg = g_jit_dbg.load();
  for (int i = 0; i < h && i < 16; ++i) prm.hb[i] = ps.hb[i];
  for (int j = 0; j < parsed.next && j < 8; ++j) prm.ext[j] = (signed char)parsed.ext[j];
  const unsigned long long total = (unsigned long long)batch << (n - ps.m);
  unsigned long long grid = (unsigned long long)ws.sm_count * k->resident;
  if (grid > total) grid = total;
  alignas(64) CUtensorMap tmap;
  memset(&tmap, 0, sizeof tmap);
  prm.use_tensor = 0;
  if (plan.padL == 0 && g_tensor_tma.load() && !(prm.dbg & 6) && make_tensor_map(state, n, batch, dtype, ps, &tmap, prm)) prm.use_tensor = 1;
  void *args[] = {&prm, &tmap};
  cudaError_t e = cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(160), args, smem, st);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return fail(std::string("tqb_spec_pass launch failed: ") + cudaGetErrorString(e));
  g_spec_launches.fetch_add(1);
  *used = true;
  return 0;
}

}  // namespace tqb

using namespace tqb;

extern "C" {

int tqb_set_jit(int mode) {
  if (mode >= 1024) {   // 1024 + v: code shape of the generated kernels (see g_spec_loop)
    const int v = mode - 1024;
    g_spec_loop.store(v < 0 || v > 2 ? 2 : v);
    return g_jit_mode.load();
  }
  if (mode >= 512) {   // 512 + v: tensor-map staging (cp.async.bulk.tensor) on / off
    g_tensor_tma.store(mode - 512 ? 1 : 0);
    return g_jit_mode.load();
  }
  if (mode >= 256) {   // 256 + flags: profiling switches of the specialised kernels (results are WRONG with any flag set)
    g_jit_dbg.store(mode - 256);
    return g_jit_mode.load();
  }
  return g_jit_mode.exchange(mode < 0 ? 0 : (mode > 2 ? 2 : mode));
}

int tqb_set_jit_cache(const char *dir) {
  std::lock_guard<std::mutex> lk(g_jit_mu);
  g_cache_dir = dir ? dir : "";
  if (!g_cache_dir.empty()) mkdir(g_cache_dir.c_str(), 0755);
  return 0;
}

int tqb_jit_shutdown(void) {
  jit_shutdown();
  return 0;
}

int tqb_jit_wait(void) {
  std::unique_lock<std::mutex> lk(g_jit_mu);
  g_jit_cv.wait(lk, [] { return g_pending <= 0 || g_stop; });
  return 0;
}

int tqb_jit_stats(int64_t *out4) {
  std::lock_guard<std::mutex> lk(g_jit_mu);
  out4[0] = g_spec_launches.load();
  out4[1] = g_spec_compiles.load();
  out4[2] = g_spec_disk_hits.load();
  out4[3] = (int64_t)g_cache.size();
  return 0;
}

// The generated constants + the kernel template of one pass (what NVRTC would compile); returns the length needed
// (including the terminating 0), or a negative value when the pass is not eligible (tqb_last_error() says why).
int64_t tqb_spec_source(const tqb_pass *pass, const tqb_gate *gates_host, int dtype, int with_template, char *buf,
                        int64_t cap) {
  if (!pass || !gates_host) return fail("tqb_spec_source: null argument");
  SpecPlan plan;
  std::string why;
  if (!spec_plan(*pass, gates_host, dtype, plan, why)) return fail("pass not eligible for specialisation: " + why);
  std::string s = spec_header(plan);
  if (with_template) s += kSpecTemplate;
  if (buf && cap > 0) {
    const size_t n = std::min((size_t)cap - 1, s.size());
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return (int64_t)s.size() + 1;
}

// NVRTC-compile one pass for sm_100a without launching anything (works without a GPU); fills the disk cache.
int tqb_spec_compile(const tqb_pass *pass, const tqb_gate *gates_host, int dtype) {
  if (!pass || !gates_host) return fail("tqb_spec_compile: null argument");
  SpecPlan plan;
  std::string why;
  if (!spec_plan(*pass, gates_host, dtype, plan, why)) return fail("pass not eligible for specialisation: " + why);
  SpecKernel k;
  k.header = spec_header(plan);
  build_kernel(k);
  if (k.state.load() != SpecKernel::READY) return fail("NVRTC: " + k.log);
  return (int)k.cubin.size();
}

}  // extern "C"
