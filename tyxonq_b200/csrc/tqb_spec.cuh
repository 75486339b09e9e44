// tqb_spec.cuh -- the fused gate pass, SPECIALISED per pass shape.
//
// This file is the hand-written source of every specialised pass kernel.  It is never compiled on its own: the
// library (tqb_jit.cu) prepends a small generated block of compile-time constants -- the pass description
// `tqbs::G[]`, one entry per gate, with every bit position, chain type and thread mapping of THAT pass -- and
// hands the text to NVRTC (sm_100a).  With the description known at compile time the per-gate work of the generic
// kernel (descriptor decode, group location, 2 x 16 address computations, dispatch on chain type: 71 % of its
// executed instructions, profiles/r01_lean_pass_ncu_full.txt) folds into immediates: a gate is 16 LDS.128 with
// immediate offsets, the arithmetic, 16 STS.128.
//
// Structure of a pass kernel (same producer / consumer split as tile_pass_lean_kernel in tqb_tile.cu):
//   * persistent grid, CTA = 128 consumer threads + 1 producer warp, two tile buffers of 2^M amplitudes;
//   * the producer moves tiles with cp.async.bulk (one bulk copy per contiguous run) and mbarriers;
//   * every consumer thread owns 2^RBITS amplitudes of the tile per gate (RBITS = M - 7): the gate's target bits plus
//     "filler" bits are REGISTER bits, the other 7 tile bits are THREAD bits (tb[j] = tile bit of thread-id bit j);
//   * thread-id bits 5,6 (the warp index) sit on two tile bits that no gate of a SEGMENT of the pass targets, so inside
//     a segment a warp only ever touches its own quarter of the tile and gates are separated by __syncwarp() instead
//     of a CTA barrier (sync = 1); segments are separated by the consumers' named barrier (sync = 2);
//   * lane bits 0-2 are chosen by the generator so that the 8 lanes of a quarter warp hit 8 different 16-byte bank
//     groups (padded tile layout when a gate keeps index bits 0-2 in registers).
//
// The same text compiles with g++ (-DTQB_SPEC_EMU) into the CPU-side emulator used by the tests (tests/emu/).
//
// Replaces: apply_1q_statevector / apply_2q_statevector / apply_kqubit_unitary
// (reference libs/quantum_library/kernels/statevector.py:28-129) and the op loop of StatevectorEngine.run/state
// (devices/simulators/statevector/engine.py:52-374, 914-1038) for passes made of 1-qubit-layer gates.

#ifdef TQB_SPEC_EMU
#define SPEC_DEV inline
#define SPEC_HD inline constexpr
#else
#define SPEC_DEV __device__ __forceinline__
#define SPEC_HD __host__ __device__ constexpr
#endif

namespace tqbs {

// ---- generated before this point ------------------------------------------------------------------
//   typedef double T;                       amplitude component type
//   constexpr int M, L, PADL, NG, NEXT, MAT_COUNT, RBITS, LOOPED, BATCHED;
//   struct GateC;  constexpr GateC G[NG];
// (struct GateC is declared by the generated block so that both sides agree on the field order.)

constexpr int TBITS = 7;            // thread bits: 128 consumer threads
constexpr int CT = 1 << TBITS;
constexpr int D = 1 << RBITS;       // amplitudes per thread per gate
constexpr int H = M - L;
constexpr int ES = 2 * (int)sizeof(T);   // bytes per amplitude

constexpr int K_DENSE1 = 0, K_DIAG = 1, K_MUX = 4, K_CHAIN = 5, K_ROT = 6;

template <typename S>
struct alignas(2 * sizeof(S)) cplx {
  S x, y;
};
typedef cplx<T> amp;

template <int V>
struct IC {
  static constexpr int value = V;
};
template <int I, int N, class F>
SPEC_DEV void static_for_impl(F &f) {
  if constexpr (I < N) {
    f(IC<I>{});
    static_for_impl<I + 1, N>(f);
  }
}
template <int N, class F>
SPEC_DEV void static_for(F f) {
  static_for_impl<0, N>(f);
}

// byte offset of tile element e (padded layout: 16 bytes after every run of 2^PADL amplitudes); additive over
// disjoint bit sets, so offset(base | s) = offset(base) + offset(s)
SPEC_HD unsigned pbyte(unsigned e) { return e * (unsigned)ES + (PADL ? ((e >> PADL) << 4) : 0u); }

// element offset of register index s of gate GI
template <int GI>
SPEC_HD unsigned soff(int s) {
  unsigned o = 0;
  for (int i = 0; i < RBITS; ++i)
    if ((s >> i) & 1) o |= 1u << G[GI].rb[i];
  return o;
}
template <int GI>
SPEC_HD bool is_reg_bit(int p) {
  for (int i = 0; i < RBITS; ++i)
    if (G[GI].rb[i] == p) return true;
  return false;
}
template <int GI>
SPEC_HD int reg_index(int p) {
  for (int i = 0; i < RBITS; ++i)
    if (G[GI].rb[i] == p) return i;
  return -1;
}

// complex64: when tile bit 0 is a register bit of the gate (the generator arranges that: a filler when no layer targets
// it), the two amplitudes that differ in it are one aligned 16-byte shared-memory access -- 8 lanes per wavefront, the
// same bank picture as complex128 -- instead of two 8-byte ones.  pair_reg: that register bit's index, or -1.
template <int GI>
SPEC_HD int pair_reg() {
  return ES == 8 ? reg_index<GI>(0) : -1;
}
struct alignas(4 * sizeof(T)) amp2 {
  cplx<T> a, b;
};
// v[s] <- element at addr(s) for every register index s; J0 >= 0: s and s | (1 << J0) are adjacent (one access)
template <int N, int J0, class AddrF>
SPEC_DEV void load_amps(amp (&v)[N], AddrF addr) {
  static_for<N>([&](auto sc) {
    constexpr int s = decltype(sc)::value;
    if constexpr (J0 < 0) {
      v[s] = *reinterpret_cast<const amp *>(addr(sc));
    } else if constexpr (!((s >> (J0 < 0 ? 0 : J0)) & 1)) {
      const amp2 p = *reinterpret_cast<const amp2 *>(addr(sc));
      v[s] = p.a;
      v[s | (1 << (J0 < 0 ? 0 : J0))] = p.b;
    }
  });
}
template <int N, int J0, class AddrF>
SPEC_DEV void store_amps(const amp (&v)[N], AddrF addr) {
  static_for<N>([&](auto sc) {
    constexpr int s = decltype(sc)::value;
    if constexpr (J0 < 0) {
      *reinterpret_cast<amp *>(addr(sc)) = v[s];
    } else if constexpr (!((s >> (J0 < 0 ? 0 : J0)) & 1)) {
      amp2 p;
      p.a = v[s];
      p.b = v[s | (1 << (J0 < 0 ? 0 : J0))];
      *reinterpret_cast<amp2 *>(addr(sc)) = p;
    }
  });
}

// DIAG: the part of the table index that comes from register bits, for register index s
template <int GI>
SPEC_HD unsigned diag_tvar(int s) {
  unsigned t = 0;
  for (int j = 0; j < G[GI].R; ++j) {
    const int code = G[GI].dbits[j];
    if (code >= 0 && code < 64 && is_reg_bit<GI>(code)) t |= (unsigned)((s >> reg_index<GI>(code)) & 1) << j;
  }
  return t;
}

// Register fusion: a gate whose mapping equals its predecessor's (sync == -1: same thread, same amplitudes) takes its
// block straight from the predecessor's registers instead of a shared-memory round trip -- unless it is a rotation
// chain with a run-time control, whose loads rename the inputs.
SPEC_HD bool fuse_in(int gi) {
  return gi > 0 && gi < NG && G[gi].sync == -1 && !(G[gi].kind == K_ROT && G[gi].ctrl >= 0);
}
// register index of gate GI that holds tile element offset o (o is a union of GI's register bits)
template <int GI>
SPEC_HD int reg_of_offset(unsigned o) {
  int s = 0;
  for (int i = 0; i < RBITS; ++i)
    if ((o >> G[GI].rb[i]) & 1u) s |= 1 << i;
  return s;
}

// tile element index of the thread's block (register bits zero)
template <int GI>
SPEC_DEV unsigned thread_base(unsigned tid) {
  unsigned e = 0;
  static_for<TBITS>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    constexpr int p = G[GI].tb[j];
    e |= ((tid >> j) & 1u) << p;
  });
  return e;
}

// value of a (non-register) bit given as: tile-local position p < 64 (a thread bit of gate GI), or 64 + j = value of
// outside-the-tile bit slot j (constant for the tile: bit j of extv)
template <int GI, int CODE>
SPEC_DEV unsigned bit_value(unsigned base, unsigned extv) {
  if constexpr (CODE < 0) return 0u;
  else if constexpr (CODE >= 64) return (extv >> (CODE - 64)) & 1u;
  else return (base >> CODE) & 1u;
}

// ---- arithmetic on the register block ------------------------------------------------------------------
// dense 2x2 on register bit I; SELBIT >= 0: pairs whose register bit SELBIT is 1 use (b..) instead of (a..)
template <int I, int SELBIT, int N>
SPEC_DEV void layer_dense(amp (&v)[N], const amp *Ma, const amp *Mb) {
  const amp a00 = Ma[0], a01 = Ma[1], a10 = Ma[2], a11 = Ma[3];
  amp b00 = a00, b01 = a01, b10 = a10, b11 = a11;
  if constexpr (SELBIT >= 0) {
    b00 = Mb[0]; b01 = Mb[1]; b10 = Mb[2]; b11 = Mb[3];
  }
#pragma unroll
  for (int s = 0; s < N; ++s) {
    if (s & (1 << I)) continue;
    const bool sel = SELBIT >= 0 && ((s >> (SELBIT >= 0 ? SELBIT : 0)) & 1);
    const amp m00 = sel ? b00 : a00, m01 = sel ? b01 : a01, m10 = sel ? b10 : a10, m11 = sel ? b11 : a11;
    const amp a = v[s], b = v[s | (1 << I)];
    amp x, y;
    // same operation order as cmac() of the generic kernel (tqb_core.cuh): acc = 0; acc += m*a; acc += m*b
    x.x = (T)0 + m00.x * a.x; x.x -= m00.y * a.y; x.y = (T)0 + m00.x * a.y; x.y += m00.y * a.x;
    x.x += m01.x * b.x; x.x -= m01.y * b.y; x.y += m01.x * b.y; x.y += m01.y * b.x;
    y.x = (T)0 + m10.x * a.x; y.x -= m10.y * a.y; y.y = (T)0 + m10.x * a.y; y.y += m10.y * a.x;
    y.x += m11.x * b.x; y.x -= m11.y * b.y; y.y += m11.x * b.y; y.y += m11.y * b.x;
    v[s] = x;
    v[s | (1 << I)] = y;
  }
}

// rotation-form layers (tqb_core.cuh rot_layer / rot_layer_scaled): M = [[a, i r], [i r, a]] (TYPE 0) or
// [[a, -r], [r, a]] (TYPE 1); MUXED: the pair's inputs are renamed when register bit I-1 is 1 (the fused cx)
template <int I, int TYPE, bool MUXED, int N>
SPEC_DEV void rot_layer(amp (&v)[N], const T a, const T r) {
#pragma unroll
  for (int s = 0; s < N; ++s) {
    if (s & (1 << I)) continue;
    const bool sel = MUXED && I > 0 && ((s >> (I > 0 ? I - 1 : 0)) & 1);
    const int lo = s, hi = s | (1 << I);
    const amp x0 = v[sel ? hi : lo], x1 = v[sel ? lo : hi];
    const T rr = (TYPE == 1 && sel) ? -r : r;
    amp y0, y1;
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

// INVC: 0 / 1 = the form is known when the kernel is generated (tqb_gate.off_b bits 8..11: no run-time branch, one copy of the
// layer's code), -1 = read from the coefficient at run time
template <int I, int TYPE, bool MUXED, int INVC, int N>
SPEC_DEV void rot_layer_scaled(amp (&v)[N], const T k, const bool inv_rt) {
  const bool inv = INVC < 0 ? inv_rt : (INVC != 0);
  if (!inv) {
#pragma unroll
    for (int s = 0; s < N; ++s) {
      if (s & (1 << I)) continue;
      const bool sel = MUXED && ((s >> (I > 0 ? I - 1 : 0)) & 1);
      const int lo = s, hi = s | (1 << I);
      const amp x0 = v[sel ? hi : lo], x1 = v[sel ? lo : hi];
      amp y0, y1;
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
    for (int s = 0; s < N; ++s) {
      if (s & (1 << I)) continue;
      const bool sel = MUXED && ((s >> (I > 0 ? I - 1 : 0)) & 1);
      const int lo = s, hi = s | (1 << I);
      const amp x0 = v[sel ? hi : lo], x1 = v[sel ? lo : hi];
      amp y0, y1;
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

SPEC_DEV amp cmul(const amp a, const amp b) { return amp{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }

// ---- one gate on the staged tile ------------------------------------------------------------------
// tile: byte address of the tile buffer; sm: the pass's staged matrices; extv: values of the outside-the-tile bits;
// sync(): called once, right before the first tile access (after the address arithmetic)
template <int GI, class Sync>
SPEC_DEV void apply_gate(char *tile, const amp *sm, unsigned tid, unsigned extv, amp (&v)[D], Sync sync) {
  constexpr int KIND = G[GI].kind;
  constexpr int R = G[GI].R;
  constexpr bool FIN = fuse_in(GI), FOUT = fuse_in(GI + 1);
  const unsigned base = thread_base<GI>(tid);
  char *const tb0 = tile + pbyte(base);
  constexpr int MAT = G[GI].mat;
  const amp *const Mg = sm + MAT;
  if constexpr (FIN) {   // rename the predecessor's registers into this gate's register order (no instructions)
    amp t[D];
    static_for<D>([&](auto sc) {
      constexpr int s = decltype(sc)::value;
      constexpr int sp = reg_of_offset<(GI > 0 ? GI - 1 : 0)>(soff<GI>(s));
      t[s] = v[sp];
    });
    static_for<D>([&](auto sc) {
      constexpr int s = decltype(sc)::value;
      v[s] = t[s];
    });
  }
  constexpr int J0 = pair_reg<GI>();
  auto elem = [&](auto sc) { return tb0 + pbyte(soff<GI>(decltype(sc)::value)); };
  // plain load of the block (element s -> register s)
  auto load_block = [&]() {
    if constexpr (!FIN) load_amps<D, J0>(v, elem);
  };

  if constexpr (KIND == K_ROT) {
    constexpr int TYPE = G[GI].type;
    constexpr bool MUXED = G[GI].muxed != 0;
    constexpr int E = G[GI].E;
    constexpr bool HAS_CTRL = G[GI].ctrl >= 0;
    constexpr unsigned TAB = 1u << (R + E);
    const unsigned cv = bit_value<GI, G[GI].ctrl>(base, extv);
    // register s holds INPUT element s ^ cv (cv toggles register bit 0 = layer 0's bit: the fused cx on layer 0)
    constexpr unsigned d0 = pbyte(1u << G[GI].rb[0]);
    const char *const pa = tb0 + (cv ? d0 : 0u);   // inputs of the registers with bit 0 clear
    const char *const pb = tb0 + (cv ? 0u : d0);   // ... with bit 0 set
    unsigned x = 0;
    if constexpr (E > 0) x |= bit_value<GI, G[GI].xb[0]>(base, extv);
    if constexpr (E > 1) x |= bit_value<GI, G[GI].xb[1]>(base, extv) << 1;
    // table entry of register s is P[(s & mask) ^ cv] (the gate's second table copy is the first with index bit 0
    // flipped): even registers read at +cv, odd ones at -cv, so lanes with different control values read ADJACENT
    // entries -- different banks, one wavefront -- instead of two copies 2^(R+E) entries apart
    const amp *const P = Mg + (x << R);
    const amp *const Pe = P + cv;
    const amp *const Po = P - cv;
    const amp *const coef = Mg + (HAS_CTRL ? 2u * TAB : TAB);
    const T a0 = coef[0].x;
    T r0 = coef[0].y;
    if (TYPE == 1 && cv) r0 = -r0;
    T kk[4];
    bool inv[4];
#pragma unroll
    for (int i = 1; i < 4; ++i) {
      kk[i] = (T)0;
      inv[i] = false;
      if (i < R) {
        const amp c = coef[i];
        kk[i] = c.x;
        if constexpr (G[GI].inv < 0) inv[i] = c.y != (T)0;
      }
    }
    sync();
    if constexpr (!FIN) {
      // a run-time control swaps the two halves of a pair that sits on layer 0's bit: single accesses then
      constexpr int JL = (HAS_CTRL && J0 == 0) ? -1 : J0;
      load_amps<D, JL>(v, [&](auto sc) {
        constexpr int s = decltype(sc)::value;
        return ((s & 1) ? pb : pa) + pbyte(soff<GI>(s & ~1));
      });
    }
    if constexpr (G[GI].unit_p == 0) {
      static_for<D>([&](auto sc) {
        constexpr int s = decltype(sc)::value;
        v[s] = cmul(v[s], ((s & 1) ? Po : Pe)[s & ((1 << R) - 1)]);
      });
    }
    constexpr int IV = G[GI].inv;
    if constexpr (IV >= 0 && (IV & 8) != 0) rot_layer_scaled<0, TYPE, false, 0>(v, (TYPE == 1 && cv) ? -a0 : a0, false);   // coef[0] = (t0, 0)
    else rot_layer<0, TYPE, false>(v, a0, r0);
    if constexpr (R > 1) rot_layer_scaled<(R > 1 ? 1 : 0), TYPE, MUXED, (IV < 0 ? -1 : (IV & 1))>(v, kk[1], inv[1]);
    if constexpr (R > 2) rot_layer_scaled<(R > 2 ? 2 : 0), TYPE, MUXED, (IV < 0 ? -1 : ((IV >> 1) & 1))>(v, kk[2], inv[2]);
    if constexpr (R > 3) rot_layer_scaled<(R > 3 ? 3 : 0), TYPE, MUXED, (IV < 0 ? -1 : ((IV >> 2) & 1))>(v, kk[3], inv[3]);
  } else if constexpr (KIND == K_CHAIN) {
    // general chain: layer i -> Mg[8i .. 8i+4) (selector 0), Mg[8i+4 .. 8i+8) (selector 1); layer 0 is selected by the
    // control, layer i > 0 by register bit i-1 after layer i-1
    const unsigned cv = bit_value<GI, G[GI].ctrl>(base, extv);
    sync();
    load_block();
    layer_dense<0, -1>(v, Mg + 4u * cv, Mg);
    if constexpr (R > 1) layer_dense<(R > 1 ? 1 : 0), 0>(v, Mg + 8, Mg + 12);
    if constexpr (R > 2) layer_dense<(R > 2 ? 2 : 0), 1>(v, Mg + 16, Mg + 20);
  } else if constexpr (KIND == K_MUX || KIND == K_DENSE1) {
    unsigned cv = 0;
    if constexpr (KIND == K_MUX) cv = bit_value<GI, G[GI].ctrl>(base, extv);
    sync();
    load_block();
    layer_dense<0, -1>(v, Mg + 4u * cv, Mg);
  } else {  // K_DIAG: table index bit j = bit dbits[j] (register bit, thread bit or outside bit)
    unsigned tfix = 0;
    static_for<6>([&](auto jc) {
      constexpr int j = decltype(jc)::value;
      if constexpr (j < R) {
        constexpr int code = G[GI].dbits[j];
        if constexpr (code >= 64 || !is_reg_bit<GI>(code)) tfix |= bit_value<GI, code>(base, extv) << j;
      }
    });
    const amp *const tab = Mg + tfix;
    sync();
    load_block();
    static_for<D>([&](auto sc) {
      constexpr int s = decltype(sc)::value;
      constexpr unsigned tvar = diag_tvar<GI>(s);
      v[s] = cmul(v[s], tab[tvar]);
    });
  }
  if constexpr (!FOUT) store_amps<D, J0>(v, elem);
}

// ---- the looped variant (LOOPED != 0) ---------------------------------------------------------------------------
// The unrolled form above spends ~600 instructions per complex64 gate and thread; a pass of 16 gates is 160 KB of
// straight-line code that every warp streams through once per tile, far beyond the instruction caches, and the kernel
// becomes instruction-fetch bound (profiles/r02_trotter_pass4_ncu.txt: stall no_instruction on top, issue slots 38 % busy).
// Here gates of the same CLASS (kind, chain length and type, table / control present, forms of the scaled layers) share
// ONE body, a real (not inlined) function: the call site of a gate still folds the gate's bit positions into immediates
// -- thread base, control / table bits, matrix offsets, the synchronisation -- and passes the byte strides of the
// register bits as arguments; the body forms the thread's 2^RBITS addresses with one add each.  A pass then is a few
// bodies of 5-10 KB plus ~40 instructions per gate and runs from the instruction cache.  No register fusion in this form: a gate marked sync == -1 round-trips through shared memory
// (same thread, same amplitudes: no barrier).
#ifdef TQB_SPEC_EMU
#define SPEC_BODY static
static char *g_emu_tile = nullptr;   // the bodies address the tile by 32-bit offsets from here
SPEC_DEV char *smem_at(unsigned off) { return g_emu_tile + off; }
#else
#define SPEC_BODY __device__ __noinline__
extern __shared__ __align__(128) unsigned char smem_raw[];
// 32-bit offsets into the CTA's shared memory: address arithmetic in the bodies stays 32-bit (one IADD per amplitude)
SPEC_DEV char *smem_at(unsigned off) { return reinterpret_cast<char *>(smem_raw) + off; }
#endif

SPEC_HD int pair_reg_of(int gi) {
  if (ES != 8) return -1;
  for (int i = 0; i < RBITS; ++i)
    if (G[gi].rb[i] == 0) return i;
  return -1;
}
SPEC_HD bool same_class(int i, int j) {
  if (G[i].kind != G[j].kind || pair_reg_of(i) != pair_reg_of(j)) return false;
  if (G[i].kind == K_DIAG) return true;
  if (G[i].R != G[j].R) return false;
  if (G[i].kind == K_ROT)
    return G[i].type == G[j].type && G[i].muxed == G[j].muxed && G[i].unit_p == G[j].unit_p &&
           (G[i].ctrl >= 0) == (G[j].ctrl >= 0) && G[i].inv == G[j].inv;
  return true;
}
SPEC_HD int class_of(int gi) {
  for (int j = 0; j < gi; ++j)
    if (same_class(j, gi)) return j;
  return gi;
}
SPEC_HD int ctz_c(int s) {
  int k = 0;
  while (!((s >> k) & 1)) ++k;
  return k;
}
struct Strides {
  unsigned d[5];   // byte offset (padded layout) of register bit j
};
SPEC_DEV void block_addresses(const Strides &st, unsigned a0, unsigned (&a)[D]) {
  a[0] = a0;
  static_for<D>([&](auto sc) {
    constexpr int s = decltype(sc)::value;
    if constexpr (s > 0) a[s] = a[s & (s - 1)] + st.d[ctz_c(s)];
  });
}

// rotation-form chain of class REP; tb: the thread's block (register bits zero), P: table (already offset by the extra
// bits), coef: the layer coefficients, cv: value of the control bit
template <int REP>
SPEC_BODY void rot_body(unsigned tb, const amp *P, const amp *coef, unsigned cv, Strides st) {
  constexpr int R = G[REP].R;
  constexpr int TYPE = G[REP].type;
  constexpr bool MUXED = G[REP].muxed != 0;
  constexpr bool HAS_CTRL = G[REP].ctrl >= 0;
  constexpr int J0 = pair_reg_of(REP);
  const amp *const Pe = P + cv;
  const amp *const Po = P - cv;
  const T a0 = coef[0].x;
  T r0 = coef[0].y;
  if (TYPE == 1 && cv) r0 = -r0;
  T kk[4];
  bool inv[4];
#pragma unroll
  for (int i = 1; i < 4; ++i) {
    kk[i] = (T)0;
    inv[i] = false;
    if (i < R) {
      const amp c = coef[i];
      kk[i] = c.x;
      if constexpr (G[REP].inv < 0) inv[i] = c.y != (T)0;
    }
  }
  const unsigned dcv = cv ? st.d[0] : 0u, dncv = cv ? 0u : st.d[0];
  unsigned a[D];
  block_addresses(st, tb, a);
  amp v[D];
  // register s holds INPUT element s ^ cv (the fused cx on layer 0)
  constexpr int JL = (HAS_CTRL && J0 == 0) ? -1 : J0;
  load_amps<D, JL>(v, [&](auto sc) {
    constexpr int s = decltype(sc)::value;
    if constexpr (HAS_CTRL) return smem_at(a[s & ~1] + ((s & 1) ? dncv : dcv));
    else return smem_at(a[s]);
  });
  if constexpr (G[REP].unit_p == 0) {
    static_for<D>([&](auto sc) {
      constexpr int s = decltype(sc)::value;
      v[s] = cmul(v[s], ((s & 1) ? Po : Pe)[s & ((1 << R) - 1)]);
    });
  }
  constexpr int IV = G[REP].inv;
  if constexpr (IV >= 0 && (IV & 8) != 0) rot_layer_scaled<0, TYPE, false, 0>(v, (TYPE == 1 && cv) ? -a0 : a0, false);   // coef[0] = (t0, 0)
  else rot_layer<0, TYPE, false>(v, a0, r0);
  if constexpr (R > 1) rot_layer_scaled<(R > 1 ? 1 : 0), TYPE, MUXED, (IV < 0 ? -1 : (IV & 1))>(v, kk[1], inv[1]);
  if constexpr (R > 2) rot_layer_scaled<(R > 2 ? 2 : 0), TYPE, MUXED, (IV < 0 ? -1 : ((IV >> 1) & 1))>(v, kk[2], inv[2]);
  if constexpr (R > 3) rot_layer_scaled<(R > 3 ? 3 : 0), TYPE, MUXED, (IV < 0 ? -1 : ((IV >> 2) & 1))>(v, kk[3], inv[3]);
  store_amps<D, J0>(v, [&](auto sc) { return smem_at(a[decltype(sc)::value]); });
}

// diagonal gate of class REP: tab = table + the index bits that are not register bits; dw.d[j] = table-index weight of
// register bit j
template <int REP>
SPEC_BODY void diag_body(unsigned tb, const amp *tab, Strides st, Strides dw) {
  constexpr int J0 = pair_reg_of(REP);
  unsigned a[D], t[D];
  block_addresses(st, tb, a);
  block_addresses(dw, 0u, t);
  amp v[D];
  load_amps<D, J0>(v, [&](auto sc) { return smem_at(a[decltype(sc)::value]); });
  static_for<D>([&](auto sc) {
    constexpr int s = decltype(sc)::value;
    v[s] = cmul(v[s], tab[t[s]]);
  });
  store_amps<D, J0>(v, [&](auto sc) { return smem_at(a[decltype(sc)::value]); });
}

// K_CHAIN (general), K_MUX, K_DENSE1 of class REP: dense 2x2 layers; M0: layer 0's matrix (control applied), Mg: the gate's
template <int REP>
SPEC_BODY void dense_body(unsigned tb, const amp *M0, const amp *Mg, Strides st) {
  constexpr int KIND = G[REP].kind;
  constexpr int R = G[REP].R;
  constexpr int J0 = pair_reg_of(REP);
  unsigned a[D];
  block_addresses(st, tb, a);
  amp v[D];
  load_amps<D, J0>(v, [&](auto sc) { return smem_at(a[decltype(sc)::value]); });
  layer_dense<0, -1>(v, M0, Mg);
  if constexpr (KIND == K_CHAIN && R > 1) layer_dense<(R > 1 ? 1 : 0), 0>(v, Mg + 8, Mg + 12);
  if constexpr (KIND == K_CHAIN && R > 2) layer_dense<(R > 2 ? 2 : 0), 1>(v, Mg + 16, Mg + 20);
  store_amps<D, J0>(v, [&](auto sc) { return smem_at(a[decltype(sc)::value]); });
}

// the call site of gate GI: everything about the gate is an immediate here; sync() as in apply_gate
template <int GI, class Sync>
SPEC_DEV void apply_gate_looped(char *tile, const amp *sm, unsigned tid, unsigned extv, Sync sync) {
  constexpr int KIND = G[GI].kind;
  constexpr int R = G[GI].R;
  constexpr int REP = class_of(GI);
  const unsigned base = thread_base<GI>(tid);
#ifdef TQB_SPEC_EMU
  g_emu_tile = tile;
  const unsigned tb = pbyte(base);
#else
  const unsigned tb = (unsigned)(tile - reinterpret_cast<char *>(smem_raw)) + pbyte(base);
#endif
  const amp *const Mg = sm + G[GI].mat;
  Strides st;
  static_for<5>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    constexpr unsigned dj = j < RBITS ? pbyte(1u << G[GI].rb[j < RBITS ? j : 0]) : 0u;
    st.d[j] = dj;
  });
  if constexpr (KIND == K_ROT) {
    constexpr int E = G[GI].E;
    constexpr unsigned TAB = 1u << (R + E);
    const unsigned cv = bit_value<GI, G[GI].ctrl>(base, extv);
    unsigned x = 0;
    if constexpr (E > 0) x |= bit_value<GI, G[GI].xb[0]>(base, extv);
    if constexpr (E > 1) x |= bit_value<GI, G[GI].xb[1]>(base, extv) << 1;
    sync();
    rot_body<REP>(tb, Mg + (x << R), Mg + (G[GI].ctrl >= 0 ? 2u * TAB : TAB), cv, st);
  } else if constexpr (KIND == K_DIAG) {
    unsigned tfix = 0;
    Strides dw;
#pragma unroll
    for (int j = 0; j < 5; ++j) dw.d[j] = 0u;
    static_for<6>([&](auto jc) {
      constexpr int j = decltype(jc)::value;
      if constexpr (j < R) {
        constexpr int code = G[GI].dbits[j];
        if constexpr (code >= 64 || !is_reg_bit<GI>(code)) tfix |= bit_value<GI, code>(base, extv) << j;
        else dw.d[reg_index<GI>(code)] = 1u << j;
      }
    });
    sync();
    diag_body<REP>(tb, Mg + tfix, st, dw);
  } else {
    unsigned cv = 0;
    if constexpr (KIND != K_DENSE1) cv = bit_value<GI, G[GI].ctrl>(base, extv);
    sync();
    dense_body<REP>(tb, Mg + 4u * cv, Mg, st);
  }
}

#ifndef TQB_SPEC_EMU
// ---- the kernel ------------------------------------------------------------------------------------
typedef unsigned long long u64;
typedef unsigned int u32;

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64 *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
  const u32 addr = smem_u32(bar);
  u32 done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// the producer warp's wait: sleeps between polls, so that its polling does not take issue slots from the consumer warps
// that share its scheduler (compute-bound passes: the polls were 20 % of all executed instructions)
__device__ __forceinline__ void mbar_wait_relaxed(u64 *bar, u32 parity) {
  const u32 addr = smem_u32(bar);
  u32 done;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(64);
  }
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, u32 bytes, u64 *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst_gmem, const void *src_smem, u32 bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory"); }

struct SpecParams {
  void *state;
  const void *mats;          // the pass's first matrix (device memory, the state's dtype)
  u64 global_base;
  long long batch;
  int n;                     // index bits of one batch member
  int dbg;                   // profiling only: 1 = skip the gates, 2 = skip bulk loads, 4 = skip bulk stores
  signed char hb[16];        // index bit of high tile bit j (ascending)
  signed char ext[8];        // index bit of outside-the-tile slot j
  // tensor-map staging (use_tensor != 0): the state is a rank-5 tensor whose dimensions are the maximal runs of
  // tile / non-tile index bits; coordinate k of a tile = (element index >> seg_start[k]) & seg_mask[k] (0 for tile runs)
  int use_tensor;
  unsigned seg_mask[5];
  signed char seg_start[8];
};

struct alignas(64) TensorMap {   // CUtensorMap (cuda.h): 128 opaque bytes written by cuTensorMapEncodeTiled on the host
  unsigned long long opaque[16];
};

// One instruction moves a whole tile: the TMA engine walks the 2^H runs itself (SASS UTMALDG / UTMASTG) instead of the
// producer warp issuing one bulk copy per run (UBLKCP, ~70 cycles of issue each: 128 per tile).
__device__ __forceinline__ void tensor_load(void *dst_smem, const TensorMap *tm, const int (&c)[5], u64 *bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(dst_smem)),
      "l"(reinterpret_cast<u64>(tm)), "r"(smem_u32(bar)), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])
      : "memory");
}
__device__ __forceinline__ void tensor_store(const TensorMap *tm, const int (&c)[5], const void *src_smem) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                   reinterpret_cast<u64>(tm)),
               "r"(smem_u32(src_smem)), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])
               : "memory");
}

constexpr int NB = 2;
constexpr u32 NRUNS = 1u << H;
constexpr u32 RUN_BYTES = (u32)ES << L;
constexpr u32 RUN_STRIDE = RUN_BYTES + (PADL ? 16u : 0u);     // PADL is either 0 or L
constexpr u32 TILE_BYTES = (u32)ES << M;
constexpr u32 TILE_STRIDE = NRUNS * RUN_STRIDE;
constexpr u32 SMEM_BARS = NB * TILE_STRIDE;
constexpr u32 SMEM_MATS = SMEM_BARS + 64;
constexpr u32 SMEM_ROFF = SMEM_MATS + (u32)(((MAT_COUNT + 1) & ~1) * ES);
constexpr u32 SMEM_TOTAL = SMEM_ROFF + 8u * NRUNS;

extern "C" __global__ void __launch_bounds__(CT + 32, 3) tqb_spec_pass(const SpecParams prm,
                                                                        const __grid_constant__ TensorMap tmap) {
  u64 *full = reinterpret_cast<u64 *>(smem_raw + SMEM_BARS);
  u64 *done = full + 4;
  amp *smats = reinterpret_cast<amp *>(smem_raw + SMEM_MATS);
  u64 *roff = reinterpret_cast<u64 *>(smem_raw + SMEM_ROFF);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const bool producer = tid >= CT;
  for (u32 j = tid; j < NRUNS; j += CT + 32) {
    u64 o = 0;
#pragma unroll
    for (int i = 0; i < H; ++i) o |= (u64)((j >> i) & 1u) << prm.hb[i];
    roff[j] = o;
  }
  const int tbits = prm.n - M;
  const u64 total = (u64)prm.batch << tbits;
  // tiles of this CTA: first + it * stride, it < count.  Interleaved over the grid, except when gates carry one matrix
  // set per batch member (BATCHED): then every CTA takes a contiguous range of tiles, so that it stays with one member
  // for hundreds of tiles and re-stages the matrices only when the member changes.
  u64 first = blockIdx.x, stride = gridDim.x;
  u64 count = first < total ? (total - first + stride - 1) / stride : 0;
  if constexpr (BATCHED != 0) {
    const u64 per = (total + gridDim.x - 1) / gridDim.x;
    first = (u64)blockIdx.x * per;
    stride = 1;
    count = first < total ? (total - first < per ? total - first : per) : 0;
  }
  // the pass's matrices, gate by gate (a gate's data of batch member bm start at msrc + bm * mbs); consumers only
  const amp *const mats = reinterpret_cast<const amp *>(prm.mats);
  auto stage_mats = [&](u64 bm, bool all) {
    static_for<NG>([&](auto gc) {
      constexpr int GI = decltype(gc)::value;
      if (all || G[GI].mbs != 0) {
        const amp *src = mats + G[GI].msrc + bm * (u64)G[GI].mbs;
        for (int i = tid; i < G[GI].mlen; i += CT) smats[G[GI].mat + i] = src[i];
      }
    });
  };
  if (!producer && count > 0) stage_mats(first >> tbits, true);
  if (tid == 0) {
    for (int i = 0; i < NB; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&done[i], (u32)CT);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto tile_index = [&](u64 tt) -> u64 {   // element index (within the batch member) of tile tt's first amplitude
    u64 x = (tt & ((1ull << tbits) - 1ull)) << L;
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const u32 p = (u32)prm.hb[j];
      x = ((x >> p) << (p + 1u)) | (x & ((1ull << p) - 1ull));
    }
    return x;
  };
  amp *const state = reinterpret_cast<amp *>(prm.state);

  if (producer) {
    auto tile_ptr = [&](u64 it) -> amp * {
      const u64 tt = first + it * stride;
      return state + ((tt >> tbits) << prm.n) + tile_index(tt);
    };
    auto issue_load = [&](u64 it) {
      const int b = (int)(it % NB);
      unsigned char *dst = smem_raw + b * TILE_STRIDE;
      const amp *src = tile_ptr(it);
      if (prm.dbg & 2) {
        if (lane == 0) mbar_arrive(&full[b]);
        return;
      }
      if (lane == 0) mbar_expect_tx(&full[b], TILE_BYTES);
      __syncwarp();
      for (u32 j = lane; j < NRUNS; j += 32) bulk_load(dst + j * RUN_STRIDE, src + roff[j], RUN_BYTES, &full[b]);
    };
    if (prm.use_tensor) {
      // tensor-map staging: one elected lane, one instruction per tile and direction
      if (lane == 0) {
        auto coords = [&](u64 it, int (&c)[5]) {
          const u64 tt = first + it * stride;
          const u64 idx = ((tt >> tbits) << prm.n) + tile_index(tt);
#pragma unroll
          for (int k = 0; k < 5; ++k) c[k] = (int)((idx >> prm.seg_start[k]) & (u64)prm.seg_mask[k]);
        };
        int c[5];
        for (u64 it = 0; it < (u64)NB && it < count; ++it) {
          const int b = (int)(it % NB);
          coords(it, c);
          mbar_expect_tx(&full[b], TILE_BYTES);
          tensor_load(smem_raw + b * TILE_STRIDE, &tmap, c, &full[b]);
        }
        for (u64 it = 0; it < count; ++it) {
          const int b = (int)(it % NB);
          mbar_wait_relaxed(&done[b], (u32)((it / NB) & 1));
          coords(it, c);
          tensor_store(&tmap, c, smem_raw + b * TILE_STRIDE);
          bulk_commit();
          if (it + NB < count) {
            bulk_wait_read<0>();   // the store has read the buffer: refill it
            coords(it + NB, c);
            mbar_expect_tx(&full[b], TILE_BYTES);
            tensor_load(smem_raw + b * TILE_STRIDE, &tmap, c, &full[b]);
          }
        }
        bulk_wait_all0();
      }
      return;
    }
    // every buffer cycles compute -> store drain -> refill; the refill of a run is issued right behind its own store
    // (a lane waits only until ITS earlier store has been read out of shared memory)
    for (u64 it = 0; it < (u64)NB && it < count; ++it) issue_load(it);
    for (u64 it = 0; it < count; ++it) {
      const int b = (int)(it % NB);
      mbar_wait_relaxed(&done[b], (u32)((it / NB) & 1));
      const u64 nxt = it + NB;
      const bool refill = nxt < count;
      amp *dstg = tile_ptr(it);
      unsigned char *buf = smem_raw + b * TILE_STRIDE;
      if (NRUNS <= 64 && !(prm.dbg & 6)) {
        const amp *srcn = refill ? tile_ptr(nxt) : nullptr;
        if (refill && lane == 0) mbar_expect_tx(&full[b], TILE_BYTES);
        __syncwarp();
        const u32 j0 = (u32)lane, j1 = (u32)lane + 32u;
        unsigned char *s0 = buf + j0 * RUN_STRIDE, *s1 = buf + j1 * RUN_STRIDE;
        if (j0 < NRUNS) bulk_store(dstg + roff[j0], s0, RUN_BYTES);
        bulk_commit();
        if (j1 < NRUNS) bulk_store(dstg + roff[j1], s1, RUN_BYTES);
        bulk_commit();
        if (refill) {
          bulk_wait_read<1>();
          if (j0 < NRUNS) bulk_load(s0, srcn + roff[j0], RUN_BYTES, &full[b]);
          bulk_wait_read<0>();
          if (j1 < NRUNS) bulk_load(s1, srcn + roff[j1], RUN_BYTES, &full[b]);
        }
      } else {
        if (!(prm.dbg & 4))
          for (u32 j = lane; j < NRUNS; j += 32) bulk_store(dstg + roff[j], buf + j * RUN_STRIDE, RUN_BYTES);
        bulk_commit();
        if (refill) {
          bulk_wait_read<0>();
          __syncwarp();
          issue_load(nxt);
        }
      }
    }
    bulk_wait_all0();
    return;
  }

  u64 member = first >> tbits;
#pragma unroll 1
  for (u64 it = 0; it < count; ++it) {
    const int b = (int)(it % NB);
    const u32 parity = (u32)((it / NB) & 1);
    char *tile = reinterpret_cast<char *>(smem_raw + b * TILE_STRIDE);
    if constexpr (BATCHED != 0) {
      const u64 bm = (first + it * stride) >> tbits;
      if (bm != member) {   // next batch member: its matrices replace the staged ones (all consumers are between tiles)
        member = bm;
        consumer_sync();
        stage_mats(bm, false);
        consumer_sync();
      }
    }
    u32 extv = 0;
    if constexpr (NEXT > 0) {
      const u64 gidx = prm.global_base | tile_index(first + it * stride);
#pragma unroll
      for (int j = 0; j < NEXT; ++j) extv |= (u32)((gidx >> prm.ext[j]) & 1ull) << j;
    }
    if (prm.dbg & 1) {
      mbar_wait(&full[b], parity);
    } else if constexpr (LOOPED != 0) {
      static_for<NG>([&](auto gc) {
        constexpr int GI = decltype(gc)::value;
        apply_gate_looped<GI>(tile, smats, (unsigned)tid, extv, [&]() {
          constexpr int S = G[GI].sync;
          if constexpr (S == 0) mbar_wait(&full[b], parity);
          else if constexpr (S == 1) __syncwarp();
          else if constexpr (S == 2) consumer_sync();
        });
      });
    } else {
      amp v[D];
      static_for<NG>([&](auto gc) {
        constexpr int GI = decltype(gc)::value;
        apply_gate<GI>(tile, smats, (unsigned)tid, extv, v, [&]() {
          constexpr int S = G[GI].sync;
          if constexpr (S == 0) mbar_wait(&full[b], parity);
          else if constexpr (S == 1) __syncwarp();
          else if constexpr (S == 2) consumer_sync();
        });
      });
    }
    fence_proxy_async();
    mbar_arrive(&done[b]);
  }
}
#endif  // !TQB_SPEC_EMU

}  // namespace tqbs
