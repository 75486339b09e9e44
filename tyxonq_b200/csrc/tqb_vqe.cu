// tqb_vqe.cu -- a whole variational evaluation (energy + adjoint gradient) in ONE shared-memory-resident CTA.
//
// For small registers (n <= 12 qubits: ket + bra = 2 x 64 KiB of complex128) an energy + gradient evaluation is
// latency, not bandwidth: round 1 replayed a CUDA graph of ~60 kernels (one fused pass per un-applied gate, one
// reduction per parameter) at 376 us per TFIM-10 evaluation.  Here one CTA keeps ket and bra in shared memory and walks
// the whole evaluation -- forward circuit, bra = H ket, E = Re<ket|bra>, reverse sweep with the gradient inner products --
// with __syncthreads() between steps; the grid is the BATCH: one CTA per parameter vector (line searches, multi-start,
// parameter shift, population optimisers evaluate hundreds of vectors per launch on the otherwise idle SMs).
//
// Replaces, for such registers: value_and_grad of the reference's numerics backends on the VQE path
// (numerics/backends/numpy_backend.py:386-454 finite differences, pytorch_backend.py:446-564 autograd tape) around
// examples/vqetfim_benchmark.py:70-103 (exact_energy), and the adjoint sweep modelled on
// applications/chem/chem_libs/quantum_chem_library/civector_ops.py:141-200.
#include <cuda_runtime.h>

#include <string>

#include "tqb_host.h"

namespace tqb {

struct c128 {
  double x, y;
};
__device__ __forceinline__ c128 cmul(c128 a, c128 b) { return c128{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ c128 cadd(c128 a, c128 b) { return c128{a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ c128 cscale(c128 a, double s) { return c128{a.x * s, a.y * s}; }
// i^k * a
__device__ __forceinline__ c128 imul(c128 a, int k) {
  switch (k & 3) {
    case 0: return a;
    case 1: return c128{-a.y, a.x};
    case 2: return c128{-a.x, -a.y};
    default: return c128{a.y, -a.x};
  }
}

constexpr int VT = 256;

// (P v)_j for the Pauli string P = i^ny X^x Z^z:  i^ny (-1)^popc((j^x) & z) v[j ^ x]
__device__ __forceinline__ c128 pauli_at(const c128 *v, uint32_t j, uint32_t x, uint32_t z, int ny) {
  const uint32_t k = j ^ x;
  c128 a = imul(v[k], ny);
  if (__popc(k & z) & 1) a = c128{-a.x, -a.y};
  return a;
}

// exp(-i theta/2 P) on v (in place): c = cos(theta/2), s = sin(theta/2); pairs (j, j ^ x) are owned by one thread
__device__ __forceinline__ void rotate(c128 *v, int n, uint32_t x, uint32_t z, int ny, double c, double s, int tid) {
  const uint32_t dim = 1u << n;
  if (x == 0u) {
    for (uint32_t j = tid; j < dim; j += VT) {
      const double sg = (__popc(j & z) & 1) ? -s : s;   // (c - i s sign) v_j
      const c128 a = v[j];
      v[j] = c128{c * a.x + sg * a.y, c * a.y - sg * a.x};
    }
    return;
  }
  const uint32_t pb = (uint32_t)(__ffs((int)x) - 1), plow = (1u << pb) - 1u;
  for (uint32_t k = tid; k < (dim >> 1); k += VT) {
    const uint32_t j = ((k & ~plow) << 1) | (k & plow), jp = j ^ x;
    const c128 a = v[j], b = v[jp];
    const c128 pa = pauli_at(v, j, x, z, ny), pb2 = pauli_at(v, jp, x, z, ny);   // (P v)_j (from b), (P v)_jp (from a)
    // v' = c v - i s P v
    v[j] = c128{c * a.x + s * pa.y, c * a.y - s * pa.x};
    v[jp] = c128{c * b.x + s * pb2.y, c * b.y - s * pb2.x};
  }
}

// fixed dense 1-qubit gate on index bit b (M row-major 2x2), optionally its conjugate transpose
__device__ __forceinline__ void dense1(c128 *v, int n, int b, const double *M, bool dag, int tid) {
  c128 m00{M[0], M[1]}, m01{M[2], M[3]}, m10{M[4], M[5]}, m11{M[6], M[7]};
  if (dag) {
    const c128 t = m01;
    m00.y = -m00.y; m11.y = -m11.y;
    m01 = c128{m10.x, -m10.y};
    m10 = c128{t.x, -t.y};
  }
  const uint32_t low = (1u << b) - 1u;
  for (uint32_t k = tid; k < (1u << (n - 1)); k += VT) {
    const uint32_t j = ((k & ~low) << 1) | (k & low), jp = j | (1u << b);
    const c128 a = v[j], c = v[jp];
    v[j] = cadd(cmul(m00, a), cmul(m01, c));
    v[jp] = cadd(cmul(m10, a), cmul(m11, c));
  }
}

// fixed dense 2-qubit gate: matrix index = 2 * bit(bhi) + bit(blo) (M row-major 4x4)
__device__ __forceinline__ void dense2(c128 *v, int n, int bhi, int blo, const double *M, bool dag, int tid) {
  const int p0 = bhi < blo ? bhi : blo, p1 = bhi < blo ? blo : bhi;
  const uint32_t l0 = (1u << p0) - 1u, l1 = (1u << p1) - 1u;
  for (uint32_t k = tid; k < (1u << (n - 2)); k += VT) {
    uint32_t j = ((k & ~l0) << 1) | (k & l0);
    j = ((j & ~l1) << 1) | (j & l1);
    const uint32_t idx[4] = {j, j | (1u << blo), j | (1u << bhi), j | (1u << bhi) | (1u << blo)};
    c128 a[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = v[idx[r]];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      c128 acc{0.0, 0.0};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double *e = dag ? M + 2 * (c * 4 + r) : M + 2 * (r * 4 + c);
        const c128 m{e[0], dag ? -e[1] : e[1]};
        acc = cadd(acc, cmul(m, a[c]));
      }
      v[idx[r]] = acc;
    }
  }
}

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
  return x;
}

__global__ void __launch_bounds__(VT) vqe_resident_kernel(int n, const tqb_vqe_op *__restrict__ ops, int n_ops,
                                                          const double *__restrict__ fixed_mats, const uint32_t *__restrict__ ham_x,
                                                          const int *__restrict__ ham_ptr, int n_groups,
                                                          const uint32_t *__restrict__ ham_z, const double *__restrict__ ham_coef,
                                                          const double *__restrict__ params, int n_params, double *out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t dim = 1u << n;
  c128 *ket = reinterpret_cast<c128 *>(smem_raw);
  c128 *bra = ket + dim;
  double *acc = reinterpret_cast<double *>(bra + dim);   // [0] = energy, [1 + p] = dE/dtheta_p
  const int tid = threadIdx.x;
  const double *theta = params + (size_t)blockIdx.x * n_params;
  for (uint32_t j = tid; j < dim; j += VT) ket[j] = c128{j == 0u ? 1.0 : 0.0, 0.0};
  for (int i = tid; i <= n_params; i += VT) acc[i] = 0.0;
  __syncthreads();
  // ---- forward
  for (int k = 0; k < n_ops; ++k) {
    const tqb_vqe_op op = ops[k];
    if (op.kind == 0) {
      const double th = 0.5 * (op.param >= 0 ? op.scale * theta[op.param] : op.scale);
      double s, c;
      sincos(th, &s, &c);
      rotate(ket, n, op.xmask, op.zmask, __popc(op.xmask & op.zmask), c, s, tid);
    } else if (op.kind == 1) {
      dense1(ket, n, op.bit0, fixed_mats + 2 * op.mat_off, false, tid);
    } else {
      dense2(ket, n, op.bit0, op.bit1, fixed_mats + 2 * op.mat_off, false, tid);
    }
    __syncthreads();
  }
  // ---- bra = H ket,  E = Re <ket|bra>
  double e = 0.0;
  for (uint32_t j = tid; j < dim; j += VT) {
    c128 h{0.0, 0.0};
    for (int g = 0; g < n_groups; ++g) {
      const uint32_t x = ham_x[g], src = j ^ x;
      double pr = 0.0, pi = 0.0;
      for (int t = ham_ptr[g]; t < ham_ptr[g + 1]; ++t) {
        const bool odd = __popc(src & ham_z[t]) & 1;
        pr += odd ? -ham_coef[2 * t] : ham_coef[2 * t];
        pi += odd ? -ham_coef[2 * t + 1] : ham_coef[2 * t + 1];
      }
      h = cadd(h, cmul(c128{pr, pi}, ket[src]));
    }
    bra[j] = h;
    e += ket[j].x * h.x + ket[j].y * h.y;
  }
  e = warp_sum(e);
  if ((tid & 31) == 0) atomicAdd(&acc[0], e);
  __syncthreads();
  // ---- reverse sweep: dE/dtheta_k = scale_k * Im <bra| P_k |ket> on the states AFTER gate k, then un-apply gate k on both
  for (int k = n_ops - 1; k >= 0; --k) {
    const tqb_vqe_op op = ops[k];
    if (op.kind == 0) {
      const int ny = __popc(op.xmask & op.zmask);
      if (op.param >= 0) {
        double g = 0.0;
        for (uint32_t j = tid; j < dim; j += VT) {
          const c128 pk = pauli_at(ket, j, op.xmask, op.zmask, ny), b = bra[j];
          g += b.x * pk.y - b.y * pk.x;   // Im(conj(b) * pk)
        }
        g = warp_sum(g);
        if ((tid & 31) == 0) atomicAdd(&acc[1 + op.param], op.scale * g);
        __syncthreads();
      }
      const double th = 0.5 * (op.param >= 0 ? op.scale * theta[op.param] : op.scale);
      double s, c;
      sincos(th, &s, &c);
      rotate(ket, n, op.xmask, op.zmask, ny, c, -s, tid);
      rotate(bra, n, op.xmask, op.zmask, ny, c, -s, tid);
    } else if (op.kind == 1) {
      dense1(ket, n, op.bit0, fixed_mats + 2 * op.mat_off, true, tid);
      dense1(bra, n, op.bit0, fixed_mats + 2 * op.mat_off, true, tid);
    } else {
      dense2(ket, n, op.bit0, op.bit1, fixed_mats + 2 * op.mat_off, true, tid);
      dense2(bra, n, op.bit0, op.bit1, fixed_mats + 2 * op.mat_off, true, tid);
    }
    __syncthreads();
  }
  for (int i = tid; i <= n_params; i += VT) out[(size_t)blockIdx.x * (n_params + 1) + i] = acc[i];
}

}  // namespace tqb

using namespace tqb;

extern "C" int tqb_vqe_resident(int n, const tqb_vqe_op *ops_dev, int n_ops, const double *fixed_mats_dev, const uint32_t *ham_x_dev,
                                const int32_t *ham_ptr_dev, int n_groups, const uint32_t *ham_z_dev, const double *ham_coef_dev,
                                const double *params_dev, int n_params, int64_t batch, double *out_dev, void *stream) {
  TQB_REQUIRE(n >= 1 && n <= 12 && ops_dev && n_ops >= 0 && ham_x_dev && ham_ptr_dev && n_groups >= 1 && ham_z_dev && ham_coef_dev &&
                  params_dev && n_params >= 0 && n_params <= 4096 && batch >= 1 && out_dev,
              "tqb_vqe_resident: bad arguments (n <= 12 qubits)");
  Workspace *ws = workspace();
  if (!ws) return -1;
  const size_t smem = ((size_t)32 << n) + (size_t)(n_params + 1) * sizeof(double) + 16;
  TQB_REQUIRE(smem + 1024 <= (size_t)ws->max_smem_optin, "tqb_vqe_resident: state does not fit shared memory");
  static thread_local bool configured = false;
  if (!configured) {
    TQB_CHECK_CUDA(cudaFuncSetAttribute(vqe_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ws->max_smem_optin - 1024));
    configured = true;
  }
  vqe_resident_kernel<<<(unsigned)batch, VT, smem, as_stream(stream)>>>(n, ops_dev, n_ops, fixed_mats_dev, ham_x_dev, ham_ptr_dev, n_groups,
                                                                         ham_z_dev, ham_coef_dev, params_dev, n_params, out_dev);
  TQB_CHECK_LAUNCH("vqe_resident_kernel");
  return 0;
}
