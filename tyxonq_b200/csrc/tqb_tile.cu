// tqb_tile.cu -- the fused gate pass: every CTA stages 2^m-amplitude tiles of the state in
// shared memory (contiguous runs of 2^L amplitudes, 16-byte vector accesses), applies every
// gate of the pass to the staged tile, and writes the tile back.  One pass = one read + one
// write of the state, however many gates it carries (2 * 2^n * B bytes).
//
// Replaces: apply_1q_statevector / apply_2q_statevector / apply_kqubit_unitary
// (reference libs/quantum_library/kernels/statevector.py:28-129) and the op loop of
// StatevectorEngine.run/state (devices/simulators/statevector/engine.py:52-374, 914-1038).
#include <cuda_runtime.h>

#include <mutex>
#include <string>

#include "tqb_core.cuh"
#include "tqb_host.h"

namespace tqb {

// ------------------------------------------------------------------------------------------
// library state
// ------------------------------------------------------------------------------------------
static thread_local std::string t_err;
std::atomic<int64_t> g_launches{0};
static Workspace g_ws[64];
static std::mutex g_ws_mu;

void set_error(const std::string &msg) { t_err = msg; }
int fail(const std::string &msg) {
  t_err = msg;
  return -1;
}
// The reduction kernels write their partial sums into the device's workspace; calls that run CONCURRENTLY on different
// streams (independent evaluations, e.g. the replicas of ucc.energy_and_grad_batch) must not share it.  A host thread
// selects slot k of K equal slices (tqb_workspace_slot); everything it enqueues afterwards -- CUDA-graph captures
// included, the pointers are baked in -- uses that slice.
static thread_local int t_ws_slot = 0, t_ws_slots = 1;
Workspace *workspace() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || g_ws[dev].ptr == nullptr) {
    set_error("tqb_init(device) has not been called for the current device");
    return nullptr;
  }
  if (t_ws_slots <= 1) return &g_ws[dev];
  static thread_local Workspace view;
  view = g_ws[dev];
  const size_t slice = (view.bytes / (size_t)t_ws_slots) & ~(size_t)255;
  view.ptr = static_cast<char *>(view.ptr) + slice * (size_t)t_ws_slot;
  view.bytes = slice;
  return &view;
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
template <typename T>
__host__ __device__ constexpr size_t tile_region_bytes(int m) {
  return (sizeof(cplx<T>) << m) < 16 ? 16 : (sizeof(cplx<T>) << m);
}

// MAXK = 2: light variant (<= 512 threads, <= 64 registers); MAXK = 4: heavy variant
// (<= 256 threads, <= 128 registers) for passes that carry dense 3- and 4-qubit blocks.
template <typename T, int V, int MAXK>
__global__ void __launch_bounds__(256, MAXK <= 2 ? 3 : 2)
tile_pass_kernel(cplx<T> *__restrict__ state, const TileGeom geo, const long long batch,
                 const tqb_gate *__restrict__ gates, const int n_gates, const cplx<T> *__restrict__ mats) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cplx<T> *tile = reinterpret_cast<cplx<T> *>(smem_raw);
  // layout: tile | staged matrices | run-offset table | descriptors (each region keeps 16/8-byte alignment)
  cplx<T> *smats = reinterpret_cast<cplx<T> *>(smem_raw + tile_region_bytes<T>(geo.m));
  uint64_t *roff = reinterpret_cast<uint64_t *>(smats + geo.mat_count);
  tqb_gate *sg = reinterpret_cast<tqb_gate *>(roff + (1u << geo.h));

  const int tid = threadIdx.x, nthreads = blockDim.x;
  for (uint32_t j = tid; j < (1u << geo.h); j += nthreads) roff[j] = run_offset(geo, j);
  for (int i = tid; i < geo.mat_count; i += nthreads) smats[i] = mats[geo.mat_begin + i];
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(gates);
    uint32_t *dst = reinterpret_cast<uint32_t *>(sg);
    const int nw = n_gates * (int)(sizeof(tqb_gate) / 4);
    for (int i = tid; i < nw; i += nthreads) dst[i] = src[i];
  }
  __syncthreads();
  const cplx<T> *mat_base = geo.mat_count > 0 ? smats - geo.mat_begin : mats;

  const int tb = geo.n - geo.m;  // tile-index bits per batch member
  const unsigned long long total = (unsigned long long)batch << tb;
  for (unsigned long long tt = blockIdx.x; tt < total; tt += gridDim.x) {
    const unsigned long long b = tt >> tb;
    const uint64_t t = tt & ((1ull << tb) - 1ull);
    const uint64_t base = tile_base(geo, t);
    cplx<T> *sb = state + (b << geo.n);
    tile_load<T, V>(tile, sb, geo, roff, base, tid, nthreads);
    __syncthreads();
    for (int gi = 0; gi < n_gates; ++gi) {
      tile_apply_gate<T, MAXK>(tile, geo, roff, geo.global_base | base, sg[gi], mat_base, (size_t)b, tid, nthreads);
      __syncthreads();
    }
    tile_store<T, V>(tile, sb, geo, roff, base, tid, nthreads);
    __syncthreads();
  }
}

// ---- TMA variant -------------------------------------------------------------------------------
// Same pass, but tiles are staged by the TMA engine (cp.async.bulk, one bulk copy per contiguous
// run) into a double buffer: while all warps apply the gates to tile i, the runs of tile i+1 are
// already in flight and the runs of tile i-1 are being written back by bulk stores.  The SM issues
// no load/store instructions for the state at all; completion is tracked by one mbarrier per buffer.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
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
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Warp-specialised: blockDim = C consumer threads + one producer warp.  The producer warp issues
// every bulk load and bulk store (a UBLKCP per contiguous run, tens of cycles each) so that the
// consumers never wait for TMA issue; the two sides meet only through mbarriers:
//   full[b]  (1 arrival + tx bytes)  producer -> consumers: tile landed in buffer b
//   done[b]  (C arrivals)            consumers -> producer: gates applied, tile may be stored
// Consumers synchronise among themselves with named barrier 1.  NB = ring depth (2 or 3).
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync(int count) { asm volatile("bar.sync 1, %0;" ::"r"(count) : "memory"); }

template <typename T, int MAXK, int NB>
__global__ void __launch_bounds__(160, 3)
tile_pass_tma_kernel(cplx<T> *__restrict__ state, const TileGeom geo, const long long batch,
                     const tqb_gate *__restrict__ gates, const int n_gates, const cplx<T> *__restrict__ mats,
                     const int dbg) {  // dbg (profiling only): 1 = skip gate arithmetic, 2 = skip bulk loads, 4 = skip bulk stores
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: NB tiles | 8 mbarrier slots (full[NB], done[NB]) | staged matrices | run-offset table | descriptors
  const size_t tile_bytes = sizeof(cplx<T>) << geo.m;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + NB * tile_bytes);
  uint64_t *done = full + 4;
  cplx<T> *smats = reinterpret_cast<cplx<T> *>(full + 8);
  uint64_t *roff = reinterpret_cast<uint64_t *>(smats + geo.mat_count);
  tqb_gate *sg = reinterpret_cast<tqb_gate *>(roff + (1u << geo.h));

  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int ncons = nthreads - 32;  // consumer threads; the last warp is the producer
  const int lane = tid & 31;
  const bool producer = tid >= ncons;
  for (uint32_t j = tid; j < (1u << geo.h); j += nthreads) roff[j] = run_offset(geo, j);
  for (int i = tid; i < geo.mat_count; i += nthreads) smats[i] = mats[geo.mat_begin + i];
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(gates);
    uint32_t *dst = reinterpret_cast<uint32_t *>(sg);
    const int nw = n_gates * (int)(sizeof(tqb_gate) / 4);
    for (int i = tid; i < nw; i += nthreads) dst[i] = src[i];
  }
  if (tid == 0) {
    for (int i = 0; i < NB; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&done[i], (uint32_t)ncons);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int tb = geo.n - geo.m;
  const unsigned long long total = (unsigned long long)batch << tb;
  const unsigned long long first = blockIdx.x, stride = gridDim.x;
  const unsigned long long count = first < total ? (total - first + stride - 1) / stride : 0;
  const uint32_t nruns = 1u << geo.h;
  const uint32_t run_elems = 1u << geo.L;
  const uint32_t run_bytes = (uint32_t)(sizeof(cplx<T>) << geo.L);

  if (producer) {
    auto tile_ptr = [&](unsigned long long it) -> cplx<T> * {
      const unsigned long long tt = first + it * stride;
      return state + ((tt >> tb) << geo.n) + tile_base(geo, tt & ((1ull << tb) - 1ull));
    };
    auto issue_load = [&](unsigned long long it) {
      const int b = (int)(it % NB);
      cplx<T> *dst = reinterpret_cast<cplx<T> *>(smem_raw + b * tile_bytes);
      const cplx<T> *src = tile_ptr(it);
      if (dbg & 2) {
        if (lane == 0) mbar_arrive(&full[b]);
        return;
      }
      if (lane == 0) mbar_expect_tx(&full[b], (uint32_t)tile_bytes);
      __syncwarp();
      for (uint32_t j = lane; j < nruns; j += 32) bulk_load(dst + (size_t)j * run_elems, src + roff[j], run_bytes, &full[b]);
    };
    for (unsigned long long it = 0; it < (unsigned long long)(NB - 1) && it < count; ++it) issue_load(it);
    for (unsigned long long it = 0; it < count; ++it) {
      const int b = (int)(it % NB);
      if (it + NB - 1 < count) {
        // the buffer to refill last held tile it-1: this lane's bulk stores of it have read their source
        bulk_wait_read<0>();
        __syncwarp();
        issue_load(it + NB - 1);
      }
      mbar_wait(&done[b], (uint32_t)((it / NB) & 1));  // consumers finished tile it (their fence.proxy.async precedes the arrive)
      cplx<T> *dstg = tile_ptr(it);
      const cplx<T> *srcs = reinterpret_cast<const cplx<T> *>(smem_raw + b * tile_bytes);
      if (!(dbg & 4))
        for (uint32_t j = lane; j < nruns; j += 32) bulk_store(dstg + roff[j], srcs + (size_t)j * run_elems, run_bytes);
      bulk_commit();
    }
    bulk_wait_all0();
    return;
  }

  for (unsigned long long it = 0; it < count; ++it) {
    const int b = (int)(it % NB);
    mbar_wait(&full[b], (uint32_t)((it / NB) & 1));
    const unsigned long long tt = first + it * stride;
    const unsigned long long bm = tt >> tb;
    const uint64_t base = tile_base(geo, tt & ((1ull << tb) - 1ull));
    cplx<T> *tile = reinterpret_cast<cplx<T> *>(smem_raw + b * tile_bytes);
    if (dbg & 1) {
    } else if (geo.mat_count > 0) {  // matrices staged in shared memory: keep the pointer's address space known (LDS, not LD)
      const cplx<T> *sm = smats - geo.mat_begin;
      for (int gi = 0; gi < n_gates; ++gi) {
        tile_apply_gate<T, MAXK>(tile, geo, roff, geo.global_base | base, sg[gi], sm, (size_t)0, tid, ncons);
        if (gi + 1 < n_gates) consumer_sync(ncons);
      }
    } else {
      for (int gi = 0; gi < n_gates; ++gi) {
        tile_apply_gate<T, MAXK>(tile, geo, roff, geo.global_base | base, sg[gi], mats, (size_t)bm, tid, ncons);
        if (gi + 1 < n_gates) consumer_sync(ncons);
      }
    }
    fence_proxy_async();  // this thread's generic-proxy writes of the tile -> visible to the bulk-store engine
    mbar_arrive(&done[b]);
  }
}

// ---- lean variant ---------------------------------------------------------------------------------
// Same producer / consumer structure as tile_pass_tma_kernel for passes that hold only 1-qubit-layer gates (DENSE
// k = 1, DIAG, MUX, rotation-form CHAIN; matrices staged in shared memory): every pass of a hardware-efficient, QAOA
// or Trotter circuit.  The general kernel inlines all gate kinds twice (staged / global matrices), needs its full
// register budget for that and keeps the tile / gate loop state in LOCAL memory (ncu, profiles/r01_tile_sweep.md:
// 27 % of the stall samples were long-scoreboard waits on those reloads); here the body is small enough for the
// loop state to stay in registers, descriptors are read from a 16-byte aligned region and the tile base is
// computed from a register-packed copy of hb[].
// layout: NB (padded) tiles | 8 mbarrier slots | descriptors | decoded chains | staged matrices | run-offset table
// CT = consumer threads per CTA (128: <= 136 registers, 4-layer complex128 chains; a 256-thread / 72-register variant
// with twice the resident warps was measured slower and is not instantiated: profiles/r01_tile_sweep.md)
// GM = true: matrices are read from global memory (per-batch-member matrices, tqb_gate.mat_bstride) instead of the
// staged copy -- a separate instantiation so that the staged variant keeps shared-address-space loads
// PAD = true: padded tile layout (the planner asks for it when a gate of the pass keeps LOW index bits in registers)
template <typename T, int NB, int CT, bool GM, bool PAD>
__global__ void __launch_bounds__(CT + 32, 3)
tile_pass_lean_kernel(cplx<T> *__restrict__ state, const TileGeom geo, const long long batch,
                      const tqb_gate *__restrict__ gates, const int n_gates, const cplx<T> *__restrict__ mats,
                      const int dbg) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const size_t tile_bytes = sizeof(cplx<T>) << geo.m;
  // padded tile layout (tqb_core.cuh pidx): 16 bytes after every run
  const int padL = (PAD && geo.L >= 1 && geo.L <= 7 && geo.h > 0) ? geo.L : 0;
  const uint32_t pad_elems = padL ? (uint32_t)(16 / sizeof(cplx<T>)) : 0u;
  const size_t tile_stride = tile_bytes + (padL ? ((size_t)16 << geo.h) : 0);  // bytes between the NB tile buffers
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + NB * tile_stride);
  uint64_t *done = full + 4;
  tqb_gate *sg = reinterpret_cast<tqb_gate *>(full + 8);
  RotDesc *srd = reinterpret_cast<RotDesc *>(sg + n_gates);  // decoded rotation-form chains (both 48 bytes per gate)
  cplx<T> *smats = reinterpret_cast<cplx<T> *>(srd + n_gates);
  uint64_t *roff = reinterpret_cast<uint64_t *>(smats + ((geo.mat_count + 1) & ~1));

  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int ncons = nthreads - 32;
  const int lane = tid & 31;
  const bool producer = tid >= ncons;
  for (uint32_t j = tid; j < (1u << geo.h); j += nthreads) roff[j] = run_offset(geo, j);
  for (int i = tid; i < geo.mat_count; i += nthreads) smats[i] = mats[geo.mat_begin + i];
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(gates);
    uint32_t *dst = reinterpret_cast<uint32_t *>(sg);
    const int nw = n_gates * (int)(sizeof(tqb_gate) / 4);
    for (int i = tid; i < nw; i += nthreads) dst[i] = src[i];
  }
  for (int i = tid; i < n_gates; i += nthreads)
    if (gates[i].kind == TQB_GATE_CHAIN && gates[i].off_a >= 4u) {
      srd[i] = rot_decode<T>(gates[i], geo.m, padL);
#ifdef TQB_PROFILE_SWITCHES   // dbg bits 8 / 16 / 32: chain sweeps without arithmetic / tile loads / tile stores
      srd[i].flags |= ((dbg & 8) ? 1u << 7 : 0u) | ((dbg & 16) ? 1u << 14 : 0u) | ((dbg & 32) ? 1u << 15 : 0u);
#endif
    }
  if (tid == 0) {
    for (int i = 0; i < NB; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&done[i], (uint32_t)ncons);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int tb = geo.n - geo.m;
  const unsigned long long total = (unsigned long long)batch << tb;
  const unsigned long long first = blockIdx.x, stride = gridDim.x;
  const unsigned long long count = first < total ? (total - first + stride - 1) / stride : 0;
  const uint32_t nruns = 1u << geo.h;
  const uint32_t run_elems = 1u << geo.L;
  const uint32_t run_bytes = (uint32_t)(sizeof(cplx<T>) << geo.L);
  // hb[] packed into two registers (a run-time index into the by-value parameter would send it to local memory)
  uint64_t hb_lo = 0, hb_hi = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    hb_lo |= (uint64_t)(uint8_t)geo.hb[j] << (8 * j);
    hb_hi |= (uint64_t)(uint8_t)geo.hb[8 + j] << (8 * j);
  }
  auto tile_index = [&](unsigned long long tt) -> uint64_t {  // element index of tile tt's first amplitude (batch included)
    uint64_t x = (tt & ((1ull << tb) - 1ull)) << geo.L;
    for (int j = 0; j < geo.h; ++j) {
      const uint32_t p = (uint32_t)((j < 8 ? hb_lo >> (8 * j) : hb_hi >> (8 * (j - 8))) & 0xffu);
      x = ((x >> p) << (p + 1u)) | (x & ((1ull << p) - 1ull));
    }
    return x;
  };

  if (producer) {
    auto tile_ptr = [&](unsigned long long it) -> cplx<T> * {
      const unsigned long long tt = first + it * stride;
      return state + ((tt >> tb) << geo.n) + tile_index(tt);
    };
    auto issue_load = [&](unsigned long long it) {
      const int b = (int)(it % NB);
      cplx<T> *dst = reinterpret_cast<cplx<T> *>(smem_raw + b * tile_stride);
      const cplx<T> *src = tile_ptr(it);
      if (dbg & 2) {
        if (lane == 0) mbar_arrive(&full[b]);
        return;
      }
      if (lane == 0) mbar_expect_tx(&full[b], (uint32_t)tile_bytes);
      __syncwarp();
      for (uint32_t j = lane; j < nruns; j += 32) bulk_load(dst + (size_t)j * (run_elems + pad_elems), src + roff[j], run_bytes, &full[b]);
    };
    // Every buffer cycles compute -> store drain -> refill.  The refill of a buffer is issued right behind its own
    // stores, run by run: a lane waits only until ITS earlier store has been read out of shared memory
    // (cp.async.bulk.wait_group.read counts per thread) and then loads the same run of the tile NB steps ahead, so the
    // drain of one half of the tile overlaps the refill of the other (per-tile period (Tc + S + Ld) / NB with the drain
    // S and the load latency Ld in sequence before; measured with 3-gate passes: 8.2 ms against 5.5 ms of transfer).
    for (unsigned long long it = 0; it < (unsigned long long)NB && it < count; ++it) issue_load(it);
    for (unsigned long long it = 0; it < count; ++it) {
      const int b = (int)(it % NB);
      mbar_wait(&done[b], (uint32_t)((it / NB) & 1));  // consumers finished tile it (their fence.proxy.async precedes the arrive)
      const unsigned long long nxt = it + NB;
      const bool refill = nxt < count;
      cplx<T> *dstg = tile_ptr(it);
      cplx<T> *buf = reinterpret_cast<cplx<T> *>(smem_raw + b * tile_stride);
      if (nruns <= 64 && !(dbg & 6)) {
        const cplx<T> *srcn = refill ? tile_ptr(nxt) : nullptr;
        if (refill && lane == 0) mbar_expect_tx(&full[b], (uint32_t)tile_bytes);
        __syncwarp();
        const uint32_t j0 = (uint32_t)lane, j1 = (uint32_t)lane + 32u;
        cplx<T> *s0 = buf + (size_t)j0 * (run_elems + pad_elems), *s1 = buf + (size_t)j1 * (run_elems + pad_elems);
        if (j0 < nruns) bulk_store(dstg + roff[j0], s0, run_bytes);
        bulk_commit();
        if (j1 < nruns) bulk_store(dstg + roff[j1], s1, run_bytes);
        bulk_commit();
        if (refill) {
          bulk_wait_read<1>();
          if (j0 < nruns) bulk_load(s0, srcn + roff[j0], run_bytes, &full[b]);
          bulk_wait_read<0>();
          if (j1 < nruns) bulk_load(s1, srcn + roff[j1], run_bytes, &full[b]);
        }
      } else {
        if (!(dbg & 4))
          for (uint32_t j = lane; j < nruns; j += 32) bulk_store(dstg + roff[j], buf + (size_t)j * (run_elems + pad_elems), run_bytes);
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

  const cplx<T> *sm = GM ? mats : smats - geo.mat_begin;
  const int m = geo.m;
  // every gate synchronises right before its first tile access: the first gate of a tile waits for the tile to
  // land, the others for the previous gate's stores (named barrier) -- decoding and addressing run ahead of both
  struct GateSync {
    uint64_t *bar;    // != nullptr: first gate of the tile, wait on this mbarrier
    uint32_t parity;
    int ncons;
    __device__ __forceinline__ void operator()() const {
      if (bar) mbar_wait(bar, parity);
      else consumer_sync(ncons);
    }
  };
#pragma unroll 1
  for (unsigned long long it = 0; it < count; ++it) {
    const int b = (int)(it % NB);
    const uint64_t gbase = geo.global_base | tile_index(first + it * stride);
    const size_t bm = GM ? (size_t)((first + it * stride) >> tb) : 0;
    cplx<T> *tile = reinterpret_cast<cplx<T> *>(smem_raw + b * tile_stride);
    if (dbg & 1) {
      mbar_wait(&full[b], (uint32_t)((it / NB) & 1));
    } else {
#pragma unroll 1
      for (int gi = 0; gi < n_gates; ++gi) {
        const GateSync gs{gi == 0 ? &full[b] : nullptr, (uint32_t)((it / NB) & 1), ncons};
        if (sg[gi].kind == TQB_GATE_CHAIN && sg[gi].off_a >= 4u) {
          const RotDesc rd = srd[gi];
          chain_rot_dispatch<T, GateSync, (CT > 128 && sizeof(T) == 8) ? 3 : 4, (CT < 128 ? 2 : 1)>(tile, gbase, rd, sm, tid, ncons, gs, bm);
        } else {
          gs();
          tile_apply_gate_lean<T, false>(tile, m, gbase, sg[gi], sm, tid, ncons, padL, bm);
        }
      }
    }
    fence_proxy_async();
    mbar_arrive(&done[b]);
  }
}

template <typename T>
__global__ void init_basis_kernel(cplx<T> *state, int n, long long batch, unsigned long long local_index,
                                  int present) {
  const unsigned long long total = (unsigned long long)batch << n;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const unsigned long long li = i & ((1ull << n) - 1ull);
    cplx<T> v{0, 0};
    if (present && li == local_index) v.x = (T)1;
    state[i] = v;
  }
}

template <typename T, int V, int MAXK>
static int launch_pass(void *state, const TileGeom &geo, int64_t batch, const tqb_gate *gates, int n_gates,
                       const void *mats, int threads, int ctas_per_sm, const Workspace &ws, cudaStream_t st) {
  const int max_threads = 256;
  if (threads > max_threads) threads = max_threads;
  const size_t smem = tile_region_bytes<T>(geo.m) + (sizeof(uint64_t) << geo.h) + (size_t)geo.mat_count * sizeof(cplx<T>) +
                      (size_t)n_gates * sizeof(tqb_gate);
  TQB_REQUIRE(smem <= (size_t)ws.max_smem_optin, "tqb_run_passes: tile + gate list exceed shared memory");
  auto kern = tile_pass_kernel<T, V, MAXK>;
  static thread_local bool configured = false;  // one per template instantiation and host thread
  if (!configured) {
    TQB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ws.max_smem_optin));
    TQB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  const unsigned long long total = (unsigned long long)batch << (geo.n - geo.m);
  // persistent grid: exactly the CTAs that are resident at once (a second wave would run half empty)
  int resident = 0;
  TQB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, threads, smem));
  if (resident < 1) return fail("tqb_run_passes: tile does not fit on an SM");
  int per_sm = ctas_per_sm > 0 && ctas_per_sm < resident ? ctas_per_sm : resident;
  unsigned long long grid = (unsigned long long)ws.sm_count * per_sm;
  if (grid > total) grid = total;
  kern<<<(unsigned)grid, threads, smem, st>>>(reinterpret_cast<cplx<T> *>(state), geo, (long long)batch, gates,
                                               n_gates, reinterpret_cast<const cplx<T> *>(mats));
  TQB_CHECK_LAUNCH("tile_pass_kernel");
  return 0;
}

static std::atomic<int> g_use_tma{1};
static std::atomic<int> g_dbg{0};  // profiling switches of the TMA kernel (tqb_set_tma(256 + flags))

template <typename T, int MAXK, int NB>
static int launch_pass_tma(void *state, const TileGeom &geo, int64_t batch, const tqb_gate *gates, int n_gates,
                           const void *mats, int threads, int ctas_per_sm, const Workspace &ws, cudaStream_t st, bool *used) {
  *used = false;
  const int max_threads = 128;  // consumers; + 1 producer warp = 160 threads, <= 136 registers, 3 CTAs per SM
  if (threads > max_threads) threads = max_threads;
  const size_t smem = NB * (sizeof(cplx<T>) << geo.m) + 64 + (sizeof(uint64_t) << geo.h) +
                      (size_t)geo.mat_count * sizeof(cplx<T>) + (size_t)n_gates * sizeof(tqb_gate);
  // caller falls back to the single-buffer kernel
  if (smem > (size_t)ws.max_smem_optin) return 0;
  auto kern = tile_pass_tma_kernel<T, MAXK, NB>;
  static thread_local bool configured = false;
  if (!configured) {
    TQB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ws.max_smem_optin));
    TQB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  const unsigned long long total = (unsigned long long)batch << (geo.n - geo.m);
  int resident = 0;
  const int block = threads + 32;  // consumers + the producer warp
  TQB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, block, smem));
  if (resident < 1) return 0;
  int per_sm = ctas_per_sm > 0 && ctas_per_sm < resident ? ctas_per_sm : resident;
  unsigned long long grid = (unsigned long long)ws.sm_count * per_sm;
  if (grid > total) grid = total;
  kern<<<(unsigned)grid, block, smem, st>>>(reinterpret_cast<cplx<T> *>(state), geo, (long long)batch, gates, n_gates,
                                             reinterpret_cast<const cplx<T> *>(mats), g_dbg.load());
  TQB_CHECK_LAUNCH("tile_pass_tma_kernel");
  *used = true;
  return 0;
}

static std::atomic<int> g_lean{1};  // tqb_set_tma(512 + v): 0 = never use the lean kernel

template <typename T, int NB, int CT, bool GM, bool PAD>
static int launch_pass_lean(void *state, const TileGeom &geo, int64_t batch, const tqb_gate *gates, int n_gates,
                            const void *mats, int threads, int ctas_per_sm, const Workspace &ws, cudaStream_t st, bool *used) {
  *used = false;
  threads = CT;
  const bool padded = PAD && geo.L >= 1 && geo.L <= 7 && geo.h > 0;   // must match the kernel's padL
  const size_t smem = NB * ((sizeof(cplx<T>) << geo.m) + (padded ? ((size_t)16 << geo.h) : 0)) + 64 +
                      (size_t)n_gates * (sizeof(tqb_gate) + sizeof(RotDesc)) +
                      (size_t)((geo.mat_count + 1) & ~1) * sizeof(cplx<T>) + (sizeof(uint64_t) << geo.h);
  if (smem > (size_t)ws.max_smem_optin) return 0;
  auto kern = tile_pass_lean_kernel<T, NB, CT, GM, PAD>;
  static thread_local bool configured = false;
  if (!configured) {
    TQB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ws.max_smem_optin));
    TQB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  const unsigned long long total = (unsigned long long)batch << (geo.n - geo.m);
  int resident = 0;
  const int block = threads + 32;
  TQB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, block, smem));
  if (resident < 1) return 0;
  int per_sm = ctas_per_sm > 0 && ctas_per_sm < resident ? ctas_per_sm : resident;
  unsigned long long grid = (unsigned long long)ws.sm_count * per_sm;
  if (grid > total) grid = total;
  kern<<<(unsigned)grid, block, smem, st>>>(reinterpret_cast<cplx<T> *>(state), geo, (long long)batch, gates, n_gates,
                                             reinterpret_cast<const cplx<T> *>(mats), g_dbg.load());
  TQB_CHECK_LAUNCH("tile_pass_lean_kernel");
  *used = true;
  return 0;
}

}  // namespace tqb

using namespace tqb;

extern "C" {

int tqb_set_tma(int mode) {
  if (mode >= 512) {  // 512 + v: lean kernel variant on (1, default) / off (0)
    return g_lean.exchange(mode - 512 ? 1 : 0);
  }
  if (mode >= 256) {  // 256 + flags: profiling switches (results are WRONG with any flag set)
    g_dbg.store(mode - 256);
    return g_use_tma.load();
  }
  const int old = g_use_tma.exchange(mode < 0 ? 0 : (mode > 2 ? 1 : mode));
  return old;
}

int tqb_abi_version(void) { return TQB_ABI_VERSION; }
const char *tqb_last_error(void) { return t_err.c_str(); }
int64_t tqb_launch_count(void) { return g_launches.load(); }

int tqb_init(int device) {
  TQB_REQUIRE(device >= 0 && device < 64, "tqb_init: bad device");
  std::lock_guard<std::mutex> lk(g_ws_mu);
  if (g_ws[device].ptr) return 0;
  int cur = 0;
  TQB_CHECK_CUDA(cudaGetDevice(&cur));
  TQB_CHECK_CUDA(cudaSetDevice(device));
  Workspace w;
  TQB_CHECK_CUDA(cudaDeviceGetAttribute(&w.sm_count, cudaDevAttrMultiProcessorCount, device));
  TQB_CHECK_CUDA(cudaDeviceGetAttribute(&w.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  w.bytes = (size_t)32 << 20;
  TQB_CHECK_CUDA(cudaMalloc(&w.ptr, w.bytes));
  g_ws[device] = w;
  TQB_CHECK_CUDA(cudaSetDevice(cur));
  return 0;
}

int tqb_workspace_slot(int slot, int n_slots) {
  TQB_REQUIRE(n_slots >= 1 && n_slots <= 64 && slot >= 0 && slot < n_slots, "tqb_workspace_slot: bad arguments (1 <= n_slots <= 64)");
  t_ws_slot = slot;
  t_ws_slots = n_slots;
  return 0;
}

int tqb_shutdown(int device) {
  TQB_REQUIRE(device >= 0 && device < 64, "tqb_shutdown: bad device");
  std::lock_guard<std::mutex> lk(g_ws_mu);
  if (g_ws[device].ptr) {
    cudaFree(g_ws[device].ptr);
    g_ws[device] = Workspace();
  }
  return 0;
}

int tqb_device_info(int device, int *sm_count, int *max_smem_optin) {
  TQB_CHECK_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, device));
  TQB_CHECK_CUDA(cudaDeviceGetAttribute(max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  return 0;
}

int tqb_init_basis(void *state, int n, int64_t batch, int dtype, uint64_t global_base, uint64_t basis_index,
                   void *stream) {
  TQB_REQUIRE(state && n >= 0 && n < 48 && batch >= 1, "tqb_init_basis: bad arguments");
  Workspace *ws = workspace();
  if (!ws) return -1;
  const uint64_t dim = 1ull << n;
  const int present = basis_index >= global_base && (basis_index - global_base) < dim;
  const unsigned long long local = present ? (basis_index - global_base) : 0ull;
  const unsigned long long total = (unsigned long long)batch << n;
  unsigned long long blocks = (total + 255) / 256;
  const unsigned long long cap = (unsigned long long)ws->sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (dtype == TQB_C64)
    init_basis_kernel<float><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
        reinterpret_cast<cplx<float> *>(state), n, (long long)batch, local, present);
  else if (dtype == TQB_C128)
    init_basis_kernel<double><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
        reinterpret_cast<cplx<double> *>(state), n, (long long)batch, local, present);
  else
    return fail("tqb_init_basis: bad dtype");
  TQB_CHECK_LAUNCH("init_basis_kernel");
  return 0;
}

int tqb_run_passes(void *state, int n, int64_t batch, int dtype, uint64_t global_base, const tqb_pass *passes,
                   int n_passes, const tqb_gate *gates_dev, const void *mats_dev, int threads, int ctas_per_sm,
                   void *stream) {
  return tqb_run_passes2(state, n, batch, dtype, global_base, passes, n_passes, gates_dev, nullptr, mats_dev, threads,
                         ctas_per_sm, stream);
}

int tqb_run_passes2(void *state, int n, int64_t batch, int dtype, uint64_t global_base, const tqb_pass *passes,
                    int n_passes, const tqb_gate *gates_dev, const tqb_gate *gates_host, const void *mats_dev,
                    int threads, int ctas_per_sm, void *stream) {
  TQB_REQUIRE(state && n >= 0 && n < 48 && batch >= 1, "tqb_run_passes: bad state arguments");
  TQB_REQUIRE(dtype == TQB_C64 || dtype == TQB_C128, "tqb_run_passes: bad dtype");
  TQB_REQUIRE(threads >= 32 && threads <= 1024 && threads % 32 == 0, "tqb_run_passes: bad CTA size");
  TQB_REQUIRE(n_passes == 0 || (passes && gates_dev), "tqb_run_passes: null pass/gate list");
  Workspace *ws = workspace();
  if (!ws) return -1;
  for (int p = 0; p < n_passes; ++p) {
    const tqb_pass &ps = passes[p];
    TQB_REQUIRE(ps.m >= 0 && ps.m <= n && ps.L >= 0 && ps.L <= ps.m, "tqb_run_passes: bad tile shape");
    const int h = ps.m - ps.L;
    TQB_REQUIRE(h <= TQB_MAX_TILE_HIGH && h <= 12, "tqb_run_passes: too many high tile bits");
    TQB_REQUIRE(ps.n_gates >= 1 && ps.gate_begin >= 0, "tqb_run_passes: empty pass");
    TileGeom geo;
    geo.n = n; geo.m = ps.m; geo.L = ps.L; geo.h = h; geo.global_base = global_base;
    geo.mat_begin = ps.mat_begin; geo.mat_count = ps.mat_count;
    TQB_REQUIRE(ps.mat_begin >= 0 && ps.mat_count >= 0 && ps.mat_count <= 4096, "tqb_run_passes: bad staged matrix range");
    int prev = ps.L - 1;
    for (int i = 0; i < TQB_MAX_TILE_HIGH; ++i) {
      geo.hb[i] = i < h ? ps.hb[i] : 0;
      if (i < h) {
        TQB_REQUIRE(ps.hb[i] > prev && ps.hb[i] < n, "tqb_run_passes: hb must be ascending, >= L and < n");
        prev = ps.hb[i];
      }
    }
    const tqb_gate *g = gates_dev + ps.gate_begin;
    const bool heavy = ps.max_dense_k > 2;
    cudaStream_t st = as_stream(stream);
    int rc;
    // TMA staging needs runs of >= 128 bytes and room for the double buffer; else vector LDG/STG
    const size_t run_bytes = (dtype == TQB_C128 ? (size_t)16 : (size_t)8) << ps.L;
    if (g_use_tma.load() && run_bytes >= 128 && n > ps.m) {
      bool used = false;
      rc = 0;
      if (ps.max_dense_k < 0 && gates_host && g_lean.load()) {
        // the pass's own specialised kernel (tqb_jit.cu), when it is compiled and loaded
        rc = spec_try_launch(state, geo.n, batch, dtype, global_base, ps, gates_host, mats_dev, *ws, st, &used);
        if (rc) return rc;
        if (used) continue;
      }
      if (ps.max_dense_k < 0 && g_lean.load()) {
#define TQB_LEAN(T, GM, PAD) launch_pass_lean<T, 2, 128, GM, PAD>(state, geo, batch, g, ps.n_gates, mats_dev, threads, ctas_per_sm, *ws, st, &used)
        const bool pad = ps.max_dense_k == -2;   // the planner's request for the padded tile layout
        if (dtype == TQB_C128) {
          if (ps.mat_count == 0) rc = pad ? TQB_LEAN(double, true, true) : TQB_LEAN(double, true, false);
          else rc = pad ? TQB_LEAN(double, false, true) : TQB_LEAN(double, false, false);
        } else {
          if (ps.mat_count == 0) rc = pad ? TQB_LEAN(float, true, true) : TQB_LEAN(float, true, false);
          else rc = pad ? TQB_LEAN(float, false, true) : TQB_LEAN(float, false, false);
        }
#undef TQB_LEAN
        if (rc) return rc;
        if (used) continue;
      }
#ifndef TQB_LEAN_ONLY  // (defined only for quick SASS inspection builds of the lean kernel: tools/sass_lean.sh)
#define TQB_TMA(T, MK, NB) launch_pass_tma<T, MK, NB>(state, geo, batch, g, ps.n_gates, mats_dev, threads, ctas_per_sm, *ws, st, &used)
      if (!used) {
        if (dtype == TQB_C128) rc = heavy ? TQB_TMA(double, 4, 2) : TQB_TMA(double, 2, 2);
        else rc = heavy ? TQB_TMA(float, 4, 2) : TQB_TMA(float, 2, 2);
        if (rc) return rc;
      }
#undef TQB_TMA
      if (used) continue;
    }
#define TQB_LAUNCH(T, V, MK) launch_pass<T, V, MK>(state, geo, batch, g, ps.n_gates, mats_dev, threads, ctas_per_sm, *ws, st)
    if (dtype == TQB_C128)
      rc = heavy ? TQB_LAUNCH(double, 1, 4) : TQB_LAUNCH(double, 1, 2);
    else if (ps.L >= 1)
      rc = heavy ? TQB_LAUNCH(float, 2, 4) : TQB_LAUNCH(float, 2, 2);
    else
      rc = heavy ? TQB_LAUNCH(float, 1, 4) : TQB_LAUNCH(float, 1, 2);
#undef TQB_LAUNCH
    if (rc) return rc;
#else
    }
    return fail("lean-only build");
#endif
  }
  return 0;
}

int tqb_copy(void *dst, const void *src, int64_t count, int dtype, void *stream) {
  TQB_REQUIRE(dst && src && count >= 0, "tqb_copy: bad arguments");
  const size_t bytes = (size_t)count * (dtype == TQB_C128 ? 16 : 8);
  TQB_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
  return 0;
}

}  // extern "C"
