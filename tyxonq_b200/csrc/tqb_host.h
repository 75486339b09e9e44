// tqb_host.h -- host-side helpers shared by the translation units of libtyxonq_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/tyxonq_b200.h"

namespace tqb {

void set_error(const std::string &msg);
int fail(const std::string &msg);  // records msg, returns -1
extern std::atomic<int64_t> g_launches;

// Per-device scratch for two-stage reductions (allocated by tqb_init).
struct Workspace {
  void *ptr = nullptr;
  size_t bytes = 0;
  int sm_count = 0;
  int max_smem_optin = 0;
};
Workspace *workspace();  // workspace of the current device, nullptr (with error set) if tqb_init was not called

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// tqb_jit.cu: run one lean-eligible pass with its NVRTC-specialised kernel (*used = false: caller falls back)
int spec_try_launch(void *state, int n, int64_t batch, int dtype, uint64_t global_base, const tqb_pass &ps,
                    const tqb_gate *gates_host, const void *mats_dev, const Workspace &ws, cudaStream_t st, bool *used);

#define TQB_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return ::tqb::fail(std::string(#expr) + " failed: " + cudaGetErrorString(_e));      \
  } while (0)

#define TQB_CHECK_LAUNCH(name)                                                            \
  do {                                                                                    \
    ::tqb::g_launches.fetch_add(1, std::memory_order_relaxed);                            \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess)                                                                \
      return ::tqb::fail(std::string(name) + " launch failed: " + cudaGetErrorString(_e)); \
  } while (0)

#define TQB_REQUIRE(cond, msg) \
  do {                         \
    if (!(cond)) return ::tqb::fail(msg); \
  } while (0)

}  // namespace tqb
