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
Workspace *workspace() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || g_ws[dev].ptr == nullptr) {
    set_error("tqb_init(device) has not been called for the current device");
    return nullptr;
  }
  return &g_ws[dev];
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
// MAXK = 2: light variant (<= 512 threads, <= 64 registers); MAXK = 4: heavy variant
// (<= 256 threads, <= 128 registers) for passes that carry dense 3- and 4-qubit blocks.
template <typename T, int V, int MAXK>
__global__ void __launch_bounds__(MAXK <= 2 ? 512 : 256, 2)
tile_pass_kernel(cplx<T> *__restrict__ state, const TileGeom geo, const long long batch,
                 const tqb_gate *__restrict__ gates, const int n_gates, const cplx<T> *__restrict__ mats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx<T> *tile = reinterpret_cast<cplx<T> *>(smem_raw);
  uint64_t *roff = reinterpret_cast<uint64_t *>(smem_raw + (sizeof(cplx<T>) << geo.m));
  tqb_gate *sg = reinterpret_cast<tqb_gate *>(roff + (1u << geo.h));

  const int tid = threadIdx.x, nthreads = blockDim.x;
  for (uint32_t j = tid; j < (1u << geo.h); j += nthreads) roff[j] = run_offset(geo, j);
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(gates);
    uint32_t *dst = reinterpret_cast<uint32_t *>(sg);
    const int nw = n_gates * (int)(sizeof(tqb_gate) / 4);
    for (int i = tid; i < nw; i += nthreads) dst[i] = src[i];
  }
  __syncthreads();

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
      const tqb_gate &g = sg[gi];
      const cplx<T> *mat = mats + g.mat_off + (size_t)b * g.mat_bstride;
      tile_apply_gate<T, MAXK>(tile, geo, roff, geo.global_base | base, g, mat, tid, nthreads);
      __syncthreads();
    }
    tile_store<T, V>(tile, sb, geo, roff, base, tid, nthreads);
    __syncthreads();
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
  const int max_threads = MAXK <= 2 ? 512 : 256;
  if (threads > max_threads) threads = max_threads;
  const size_t smem = (sizeof(cplx<T>) << geo.m) + (sizeof(uint64_t) << geo.h) + (size_t)n_gates * sizeof(tqb_gate);
  TQB_REQUIRE(smem <= (size_t)ws.max_smem_optin, "tqb_run_passes: tile + gate list exceed shared memory");
  auto kern = tile_pass_kernel<T, V, MAXK>;
  static thread_local size_t conf = 0;  // one per template instantiation
  if (smem > 48 * 1024 && smem > conf) {
    TQB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ws.max_smem_optin));
    conf = ws.max_smem_optin;
  }
  const unsigned long long total = (unsigned long long)batch << (geo.n - geo.m);
  int per_sm = ctas_per_sm;
  if (per_sm <= 0) {
    per_sm = (int)((size_t)(200 * 1024) / (smem + 1024));
    const int by_threads = 2048 / threads;
    if (per_sm > by_threads) per_sm = by_threads;
    if (per_sm < 1) per_sm = 1;
  }
  unsigned long long grid = (unsigned long long)ws.sm_count * per_sm;
  if (grid > total) grid = total;
  kern<<<(unsigned)grid, threads, smem, st>>>(reinterpret_cast<cplx<T> *>(state), geo, (long long)batch, gates,
                                               n_gates, reinterpret_cast<const cplx<T> *>(mats));
  TQB_CHECK_LAUNCH("tile_pass_kernel");
  return 0;
}

}  // namespace tqb

using namespace tqb;

extern "C" {

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
#define TQB_LAUNCH(T, V, MK) launch_pass<T, V, MK>(state, geo, batch, g, ps.n_gates, mats_dev, threads, ctas_per_sm, *ws, st)
    if (dtype == TQB_C128)
      rc = heavy ? TQB_LAUNCH(double, 1, 4) : TQB_LAUNCH(double, 1, 2);
    else if (ps.L >= 1)
      rc = heavy ? TQB_LAUNCH(float, 2, 4) : TQB_LAUNCH(float, 2, 2);
    else
      rc = heavy ? TQB_LAUNCH(float, 1, 4) : TQB_LAUNCH(float, 1, 2);
#undef TQB_LAUNCH
    if (rc) return rc;
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
