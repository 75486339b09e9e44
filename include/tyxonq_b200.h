/*
 * tyxonq_b200 -- C ABI of the B200-native statevector engine (libtyxonq_b200.so).
 *
 * The reference (QureGenAI-Biotech/TyxonQ) is pure Python and has no FFI of its own; the
 * drop-in seams are Python.  This header is the boundary BELOW those
 * seams: plain pointers and sizes, no torch types.  Each entry point names the reference
 * interface it replaces (paths relative to the reference's src/tyxonq/).
 *
 * Conventions
 *   - A state is `batch` consecutive arrays of 2^n complex amplitudes in device memory,
 *     complex64 (TQB_C64: float re,im) or complex128 (TQB_C128: double re,im).
 *   - "bit p" is bit p of the flat amplitude index.  TyxonQ qubit q of an n-qubit register is
 *     bit n-1-q (big-endian, libs/quantum_library/kernels/statevector.py:28-42).
 *   - For a sharded state the local array holds 2^n amplitudes and `global_base` carries the
 *     rank's high-order index bits (rank << n); diagonal tables and parity masks see
 *     global_base | local_index.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  No entry
 *     point synchronises, allocates or frees device memory except tqb_init/tqb_shutdown,
 *     so everything else can be captured into a CUDA graph.
 *   - Return value: 0 on success, negative on error; tqb_last_error() gives the message.
 */
#ifndef TYXONQ_B200_H
#define TYXONQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TQB_ABI_VERSION 2

#define TQB_C64 0
#define TQB_C128 1
#define TQB_F64 2 /* tqb_norm2 / tqb_expect_z_bits / tqb_cdf_chunks / tqb_chunk_totals / tqb_sample only: */
                  /* the array holds float64 PROBABILITIES (a noise-mixed distribution, a density-matrix  */
                  /* diagonal) instead of amplitudes: p_i is the entry itself                            */

#define TQB_GATE_DENSE 0 /* dense 2^k x 2^k matrix, k <= 4, row-major, row = output index      */
#define TQB_GATE_DIAG 1  /* diagonal: table of 2^k entries indexed by k GLOBAL index bits       */
#define TQB_GATE_PAIR 2  /* 2x2 matrix on the pair (pattern A, pattern B) of k target bits,     */
                         /* all other patterns untouched (cry, iswap, controlled-U, UCC Givens)*/
#define TQB_GATE_SWAP 3  /* exchange the amplitudes of pattern A and pattern B (x, cx, swap):   */
                         /* a PAIR gate with the matrix [[0,1],[1,0]], done without arithmetic  */
#define TQB_GATE_MUX 4   /* 1-qubit gate on tile-local bit bits[0] whose 2x2 is selected by one  */
                         /* control bit bits[1] (< 64 tile-local, 64+p index bit p outside the  */
                         /* tile): U0 at mat_off, U1 at mat_off+4.  cx fused with a 1-qubit gate */
                         /* on its target costs one sweep and needs only the TARGET tile-local. */
#define TQB_GATE_CHAIN 5 /* k = R (2 or 3) 1-qubit layers on tile-local bits bits[0..R) applied  */
                         /* to 2^R register-resident amplitudes per thread (one shared-memory   */
                         /* round trip for R gates).  Layer 0's 2x2 is selected by the control   */
                         /* bit bits[R] (< 64 tile-local, 64+p outside the tile, 127 = none),    */
                         /* layer i > 0 by the value of bit bits[i-1] after layer i-1.  Matrices: */
                         /* layer i at mat_off + 8i (selector 0) and + 8i + 4 (selector 1).      */
                         /* sbits[] = ascending positions of the targets (+ a tile-local control) */
                         /* Rotation form (off_a = 4 + 2*type + muxed, R <= 4): the data are a    */
                         /* phase table and one coefficient per layer (tqb_core.cuh              */
                         /* gate_chain_rot); off_b = extras (bits 0..1) | unit table (bit 7) |    */
                         /* bit 8+i-1: scaled layer i runs in the c form | bit 11: bits 8..10 are  */
                         /* valid (they repeat the imaginary parts of the coefficients, which the */
                         /* generic kernels read; the specialised kernels compile them in) | bit   */
                         /* 12: layer 0 runs in the t form too, its factor is in the table         */

#define TQB_MAX_DENSE_K 4
#define TQB_MAX_GATE_BITS 8
#define TQB_MAX_TILE_HIGH 16

/* One gate of a pass.  Matrix-index bit j (j = 0 least significant) lives on bits[j].
 * DENSE/PAIR/SWAP: bits[] are TILE-LOCAL positions (see tqb_pass), k <= 4.
 * DIAG (k <= 6): bits[j] < 64 is a tile-local position; bits[j] = 64 + p is index bit p that
 * is NOT in the tile (p may be a shard bit >= n of a sharded state).                          */
typedef struct tqb_gate {
  int32_t kind;
  int32_t k;
  int8_t bits[TQB_MAX_GATE_BITS];
  int8_t sbits[TQB_MAX_GATE_BITS]; /* DENSE/PAIR/SWAP: bits[] sorted ascending               */
  uint32_t off_a, off_b;           /* PAIR/SWAP: tile-local offsets of patterns A and B       */
  uint32_t mat_off;                /* offset (complex elements) into the matrix buffer        */
  uint32_t mat_bstride;            /* per-batch-member stride (complex elements), 0 = shared  */
  uint64_t zmask;                  /* PAIR: parity of popc(global_index & zmask) picks the     */
                                   /* 2x2 at mat_off (even) or mat_off+4 (odd); 0 = no parity */
} tqb_gate;                        /* 48 bytes */

/* One pass = one read-modify-write sweep over the whole state.  Every CTA stages tiles of 2^m
 * amplitudes in shared memory: the L lowest index bits (contiguous runs of 2^L amplitudes)
 * plus the m-L high bits hb[] (ascending, each >= L).  Tile-local bit j is index bit j for
 * j < L and index bit hb[j-L] otherwise.  The descriptor stream gates[gate_begin .. gate_begin+
 * n_gates) is applied to the tile before it is written back.                                             */
typedef struct tqb_pass {
  int32_t m;
  int32_t L;
  int32_t gate_begin;
  int32_t n_gates;
  int32_t max_dense_k; /* largest k of a DENSE gate in the pass (selects the kernel variant); */
                       /* -1 = the pass holds only DENSE k = 1, DIAG, MUX and CHAIN gates:      */
                       /* eligible for the lean kernel variant (matrices staged when            */
                       /* mat_count > 0, else read per batch member from global memory);        */
                       /* -2 = the same, staged in the PADDED tile layout (16 bytes after every  */
                       /* run: for passes whose gates keep index bits below 128 bytes in          */
                       /* registers, which would bank-conflict 8-way in the plain layout)        */
  int32_t mat_begin;   /* the pass's matrices are mats[mat_begin .. mat_begin+mat_count): they  */
  int32_t mat_count;   /* are staged in shared memory once per CTA; 0 = read from global memory */
  int8_t hb[TQB_MAX_TILE_HIGH];
} tqb_pass; /* 44 bytes */

/* ---- library ------------------------------------------------------------------------- */
int tqb_abi_version(void);
const char *tqb_last_error(void);
/* Allocates the per-device reduction workspace (call once per device before anything else). */
int tqb_init(int device);
int tqb_shutdown(int device);
/* Reductions write partial sums into the device's workspace.  Calls that run concurrently on different streams
 * (independent evaluations) must use different slices of it: the calling host thread selects slot `slot` of
 * `n_slots` (<= 64) equal slices for everything it enqueues afterwards, graph captures included.  (0, 1) = whole. */
int tqb_workspace_slot(int slot, int n_slots);
/* Number of kernels this library has launched since load (bench.py's gpu_launches).         */
int64_t tqb_launch_count(void);
/* max dynamic shared memory per CTA (bytes) and SM count of `device`.                        */
int tqb_device_info(int device, int *sm_count, int *max_smem_optin);

/* ---- state construction -------------------------------------------------------------- */
/* replaces init_statevector (libs/quantum_library/kernels/statevector.py:19-25): every batch
 * member becomes the basis state |basis_index> (amplitude 1 at local index basis_index if it
 * lies in this shard: basis_index - global_base in [0, 2^n)).                                */
int tqb_init_basis(void *state, int n, int64_t batch, int dtype, uint64_t global_base,
                   uint64_t basis_index, void *stream);

/* ---- gate application ---------------------------------------------------------------- */
/* replaces apply_1q_statevector / apply_2q_statevector / apply_kqubit_unitary
 * (statevector.py:28-129) and the per-op loop of StatevectorEngine.run/state
 * (devices/simulators/statevector/engine.py:52-374, 914-1038): runs n_passes fused passes in
 * place.  passes: HOST array; gates, mats: DEVICE arrays (mats in the state's dtype).
 * threads: CTA size (multiple of 32, <= 1024).  ctas_per_sm: 0 = auto.                      */
int tqb_run_passes(void *state, int n, int64_t batch, int dtype, uint64_t global_base,
                   const tqb_pass *passes, int n_passes, const tqb_gate *gates_dev,
                   const void *mats_dev, int threads, int ctas_per_sm, void *stream);

/* The same with a HOST copy of the gate descriptors (gates_host[i] == gates_dev[i]; NULL = none).  With the host
 * copy the library may run a lean-eligible pass (tqb_pass.max_dense_k < 0, matrices staged) through a kernel
 * SPECIALISED for that pass shape: the gate list, every bit position and the thread mapping become compile-time
 * constants of one hand-written kernel template (csrc/tqb_spec.cuh) compiled by NVRTC for sm_100a and cached in
 * memory and on disk; outside-the-tile bit positions, hb[] and matrix values stay run-time parameters.  Results are
 * the same as the generic kernels' to rounding.  Passes that are not eligible, or whose kernel is not ready yet
 * (asynchronous mode), run the generic kernels.                                                          */
int tqb_run_passes2(void *state, int n, int64_t batch, int dtype, uint64_t global_base,
                    const tqb_pass *passes, int n_passes, const tqb_gate *gates_dev,
                    const tqb_gate *gates_host, const void *mats_dev, int threads, int ctas_per_sm,
                    void *stream);
/* Specialisation mode: 0 = off, 1 = asynchronous (default: new shapes compile on background threads while the
 * generic kernel runs them), 2 = synchronous (compile on first use; compile errors are returned).  256 + flags:
 * profiling switches of the specialised kernels (1 = skip gates, 2 = skip bulk loads, 4 = skip bulk stores; results
 * are WRONG with any flag set).  512 + v: tile staging of the specialised kernels, v = 1 (default) one tensor copy
 * per tile (cp.async.bulk.tensor through a CUtensorMap of the state, unpadded layouts), v = 0 one bulk copy per
 * contiguous run.  1024 + v: code shape of the specialised kernels, v = 0 unrolled (one copy of the code per gate,
 * every bit position an immediate), v = 1 looped (gates of one class share a body, positions from a constant table:
 * small enough for the instruction caches), v = 2 (default) looped for complex64, unrolled for complex128.
 * Returns the old mode.                                                                                       */
int tqb_set_jit(int mode);
/* Directory of the on-disk cubin cache (NULL or "" = none).                                              */
int tqb_set_jit_cache(const char *dir);
/* Block until every queued specialisation has been compiled (asynchronous mode).                          */
int tqb_jit_wait(void);
/* Stop the background compilation threads (drops queued shapes, waits for the compilations in flight).  Called at
 * process exit; a host that unloads the library earlier calls it first.                                   */
int tqb_jit_shutdown(void);
/* out4 = { specialised launches, NVRTC compilations, disk-cache hits, shapes known }.                    */
int tqb_jit_stats(int64_t *out4);
/* Source of one pass's specialised kernel: the generated constants (+ the kernel template when with_template != 0)
 * copied into buf (cap bytes, 0-terminated); returns the size needed, negative when the pass is not eligible.
 * tqb_spec_compile: NVRTC-compile it for sm_100a without launching (needs no GPU; fills the disk cache); returns
 * the cubin size.  Both are used by the CPU test tier and by the build step that pre-warms the cache.     */
int64_t tqb_spec_source(const tqb_pass *pass, const tqb_gate *gates_host, int dtype, int with_template,
                        char *buf, int64_t cap);
int tqb_spec_compile(const tqb_pass *pass, const tqb_gate *gates_host, int dtype);

/* Tile staging mode of tqb_run_passes.  0 = vectorised LDG/STG, single buffer.  Otherwise TMA bulk
 * copies (cp.async.bulk) with mbarrier completion into a ring of tile buffers, used whenever the
 * contiguous runs are >= 128 bytes: 1 (default) and 2 = two buffers (three CTAs per SM at 32 KiB
 * tiles), 3 = three buffers when two CTAs per SM still fit (measured slower: profiles/).  A
 * producer warp issues all bulk copies.  Returns the old mode.                                   */
int tqb_set_tma(int mode);

/* ---- reductions ---------------------------------------------------------------------- */
/* out_dev[b] = sum_i |psi_b,i|^2 (float64).                                                  */
int tqb_norm2(const void *state, int n, int64_t batch, int dtype, double *out_dev, void *stream);
/* replaces expect_z_statevector called once per qubit (statevector.py:62-68, engine.py:467-473):
 * out_dev[b*n + p] = sum_i |psi_i|^2 (1 - 2 bit_p(i)) for every LOCAL bit p in one read.     */
int tqb_expect_z_bits(const void *state, int n, int64_t batch, int dtype, double *out_dev,
                      void *stream);
/* replaces the per-term Python loops of postprocessing/counts_expval.py:55-84 and
 * examples/vqetfim_benchmark.py:96-102: out_dev[b*n_masks + t] = sum_i |psi_i|^2
 * (-1)^popc((global_base|i) & masks_dev[t]).                                                  */
int tqb_expect_zmasks(const void *state, int n, int64_t batch, int dtype, uint64_t global_base,
                      const uint64_t *masks_dev, int n_masks, double *out_dev, void *stream);
/* replaces pauli_string_sum_dense + dynamics.expectation (kernels/pauli.py:74-87,
 * dynamics.py:117-126), matrix free.  Term t is coef[t] * i^ny * X^xmask Z^zmask with the
 * i^ny phase already folded into the complex coef (re,im pairs, float64).  Terms must be
 * grouped by xmask: group g covers terms [group_ptr[g], group_ptr[g+1]) and has xmask
 * group_x[g].  out_dev[b] (2 doubles: re, im) = <psi_b| H |psi_b>.  All arrays on device.
 * Local xmasks only (xmask < 2^n).                                                            */
int tqb_expect_pauli_sum(const void *state, int n, int64_t batch, int dtype, uint64_t global_base,
                         const uint64_t *group_x, const int32_t *group_ptr, int n_groups,
                         const uint64_t *term_z, const double *term_coef, double *out_dev,
                         void *stream);
/* The same expectation value, tile-staged: the xmask groups are sorted into tile LAYOUTS (tile bits = the L lowest index
 * bits + hb[], as in tqb_pass; every xmask of a layout's groups lies inside its tile bits) and each layout is ONE read of
 * the state: a CTA stages a tile in shared memory and evaluates all the layout's groups from there.  Per group g the
 * xmask is given in TILE-LOCAL bit positions (group_xl), per term t the z mask is split into its tile-local part
 * (term_zl, tile-local positions) and the rest (term_zout, index bits; shard bits >= n included).  group_ptr indexes the
 * term arrays as in tqb_expect_pauli_sum; a layout covers groups [group_begin, group_begin + n_groups) and n_terms terms.
 * flags bit 0: every group is a Hermitian operator (real Pauli coefficients): the pairs (j, j ^ x) are visited once and
 * the result is real; bit 1: every coefficient (with i^#Y folded in) is real.  layouts: HOST array; everything else on the device.  out_dev[b] = (re, im), overwritten.     */
typedef struct tqb_pauli_layout {
  int32_t m, L;
  int32_t group_begin, n_groups, n_terms;
  int8_t hb[TQB_MAX_TILE_HIGH];
} tqb_pauli_layout; /* 36 bytes */
int tqb_expect_pauli_tiled(const void *state, int n, int64_t batch, int dtype, uint64_t global_base,
                           const tqb_pauli_layout *layouts_host, int n_layouts, const uint32_t *group_xl_dev,
                           const int32_t *group_ptr_dev, const uint32_t *term_zl_dev,
                           const uint64_t *term_zout_dev, const double *term_coef_dev, int flags,
                           double *out_dev, void *stream);
/* replaces apply_op (applications/chem/chem_libs/hamiltonians_chem_library/
 * hamiltonian_builders.py:283-318) without densifying H: out = H psi, same term layout.    */
int tqb_apply_pauli_sum(const void *state, void *out, int n, int64_t batch, int dtype,
                        uint64_t global_base, const uint64_t *group_x, const int32_t *group_ptr,
                        int n_groups, const uint64_t *term_z, const double *term_coef,
                        void *stream);
/* out_dev[b] (re,im float64) = <a_b|b_b> = sum_i conj(a_i) b_i.                              */
int tqb_inner(const void *a, const void *b, int n, int64_t batch, int dtype, double *out_dev,
              void *stream);
/* Gradient reductions of the adjoint sweep (model: civector_ops.py:141-200).
 * PAIR generator: out_dev[slot] += scale * Re sum_groups s(i) (conj(bra_A) ket_B - conj(bra_B) ket_A)
 * with s = (-1)^popc(index & zmask); bits/off as in a PAIR gate with m = n (whole state).   */
int tqb_grad_pair(const void *bra, const void *ket, int n, int dtype, const tqb_gate *gate_host,
                  double scale, double *out_dev, int slot, void *stream);
/* DENSE generator D (2^k x 2^k, float64 re,im pairs on HOST, k <= 2):
 * out_dev[slot] += scale * Re <bra| D_bits |ket>.                                            */
int tqb_grad_dense(const void *bra, const void *ket, int n, int dtype, int k, const int *bits,
                   const double *gen_host, double scale, double *out_dev, int slot, void *stream);

/* Single-qubit transition matrices of a (bra, ket) pair for up to 6 index bits in ONE read of both states:
 * out_dev[8 i .. 8 i + 8) = (re, im) of T[0][0], T[0][1], T[1][0], T[1][1] for bit i of the call, T[a][b] = sum over the other
 * bits of conj(bra[.., a, ..]) ket[.., b, ..].  <bra| A_q |ket> = sum_ab A[a][b] T_q[a][b] for any single-qubit operator A:
 * one call gives the adjoint gradient of every single-qubit gate acting on those qubits in a layer (the reverse sweep of
 * civector_ops.py:141-200 layer by layer instead of gate by gate).  The bits of interest must be tile bits: tile = the L
 * lowest index bits + hb[] (m <= 12 bits in all), tile_bits[] = their tile-local positions.  out_dev is overwritten.   */
int tqb_transition_1q(const void *bra, const void *ket, int n, int dtype, int m, int L, const int8_t *hb,
                      const int8_t *tile_bits, int n_bits, double *out_dev, void *stream);

/* Whole sweeps of PAIR rotations on a SMALL state (2^n amplitudes resident in L2) in ONE launch: a
 * persistent grid walks the steps with a grid-wide barrier between them instead of one launch per gate.
 * Step j rotates the pair (pattern A, pattern B) of its k target bits by [[c, -s], [s, c]] where the
 * parity of popc(index & zmask) is even and by [[c, s], [-s, c]] where it is odd (UCC excitation
 * exp(theta G), statevector_ops.py:140-168).  mode 0: apply steps 0..n_steps-1 to `ket` only (forward).
 * mode 1: reverse sweep of the adjoint gradient (civector_ops.py:141-200): for every step,
 * out_dev[slot] += scale * Re sum s(i) (conj(bra_A) ket_B - conj(bra_B) ket_A), then the rotation is
 * applied to ket AND bra (pass the un-apply angles).  steps_dev: device array of tqb_pair_step.
 * sync_dev: one zero-initialised 64-bit word in device memory per call (grid barrier counter).    */
typedef struct tqb_pair_step {
  int32_t k;
  int32_t slot;
  int8_t sbits[TQB_MAX_GATE_BITS]; /* target bits ascending (index bits: the whole state is the tile) */
  uint64_t off_a, off_b, zmask;
  double c, s, scale;
} tqb_pair_step; /* 64 bytes */
int tqb_pair_sweep(void *ket, void *bra, int n, int dtype, const tqb_pair_step *steps_dev, int n_steps,
                   int mode, double *out_dev, unsigned long long *sync_dev, void *stream);

/* A whole variational evaluation -- forward circuit, bra = H ket, E = Re<ket|bra>, reverse sweep with the gradient inner
 * products -- in ONE shared-memory-resident CTA per parameter vector (n <= 12 qubits, complex128): replaces value_and_grad
 * of the reference's numerics backends (numpy_backend.py:386-454, pytorch_backend.py:446-564) around
 * examples/vqetfim_benchmark.py:70-103 for small registers, where an evaluation is latency, not bandwidth.
 * ops (device array): kind 0 = Pauli rotation exp(-i theta/2 P), P = i^popc(x&z) X^xmask Z^zmask on index bits,
 * theta = scale * params[param] (param < 0: theta = scale); kind 1 / 2 = fixed 1- / 2-qubit matrix at fixed_mats[mat_off]
 * (complex128 entries, row-major; 2-qubit index = 2 * bit(bit0) + bit(bit1)).  Hamiltonian in the layout of
 * tqb_expect_pauli_sum with 32-bit masks.  params_dev: [batch][n_params]; out_dev: [batch][1 + n_params] = energy, gradient. */
typedef struct tqb_vqe_op {
  int32_t kind;
  int32_t param;
  uint32_t xmask, zmask;
  int32_t bit0, bit1;
  int32_t mat_off;
  int32_t reserved;
  double scale;
} tqb_vqe_op; /* 40 bytes */
int tqb_vqe_resident(int n, const tqb_vqe_op *ops_dev, int n_ops, const double *fixed_mats_dev,
                     const uint32_t *ham_x_dev, const int32_t *ham_ptr_dev, int n_groups,
                     const uint32_t *ham_z_dev, const double *ham_coef_dev, const double *params_dev,
                     int n_params, int64_t batch, double *out_dev, void *stream);

/* Reduced density matrix of index bit `bit`, per batch member: out_dev[4b..4b+4) = rho00, rho11,
 * Re rho01, Im rho01 with rho01 = sum psi_0 conj(psi_1).  Gives the Born probabilities
 * p_i = tr(K_i^+ K_i rho) of a 1-qubit Kraus channel in ONE read of the state (replaces the m
 * trial applications of apply_kraus_statevector, libs/quantum_library/kernels/statevector.py:176-186). */
int tqb_reduced_1q(const void *state, int n, int64_t batch, int dtype, int bit, double *out_dev,
                   void *stream);

/* Density matrices ride on the same kernels as a 2n-bit vector, row bits high (rho[r, c] at index
 * r * 2^n + c): U rho U^+ is U on the row bit and conj(U) on the column bit, a Kraus channel is
 * the 4x4 sum_i K_i (x) conj(K_i) on (row bit, column bit) -- replaces apply_1q_density /
 * apply_2q_density / apply_kraus_density (libs/quantum_library/kernels/density_matrix.py:20-140).
 * tqb_dm_diag: out_dev[b * 2^n + i] = Re rho_b[i, i] (the probabilities DensityMatrixEngine.run
 * samples from and takes <Z> of, devices/simulators/density_matrix/engine.py:108-148).        */
int tqb_dm_diag(const void *rho, int n, int64_t batch, int dtype, double *out_dev, void *stream);

/* ---- measurement --------------------------------------------------------------------- */
/* replaces StatevectorEngine._project_z (engine.py:1075-1087): zero the half with
 * bit != keep, then scale by 1/sqrt(sum |psi|^2) when that sum is > 0.                      */
int tqb_project_z(void *state, int n, int64_t batch, int dtype, int bit, int keep, void *stream);
/* state *= factor (host scalar).                                                             */
int tqb_scale(void *state, int n, int64_t batch, int dtype, double factor, void *stream);
/* out_dev[b*2^n + i] = |psi_i|^2 as float64 (re*re + im*im, no FMA).                         */
int tqb_probabilities(const void *state, int n, int64_t batch, int dtype, double *out_dev,
                      void *stream);

/* Sampler: replaces numpy Generator.choice + bincount in StatevectorEngine.run
 * (engine.py:377-418).  The CDF is the documented blocked prefix sum (oracle/sv_oracle.py
 * blocked_cdf): sequential float64 sums inside chunks of TQB_SCAN_BLOCK amplitudes, then a
 * sequential sum of the chunk totals.                                                        */
#define TQB_SCAN_BLOCK 4096
/* chunk_prefix_dev: batch * (n_chunks + 1) doubles, n_chunks = max(1, 2^n / TQB_SCAN_BLOCK);
 * entry [c] = sum of chunks < c, entry [n_chunks] = total.                                   */
int tqb_cdf_chunks(const void *state, int n, int64_t batch, int dtype, double *chunk_prefix_dev,
                   void *stream);
/* idx_dev[b*shots + s] = #{ i : cdf_i / cdf_last <= uniforms_dev[b*shots + s] }  (int64).    */
int tqb_sample(const void *state, int n, int64_t batch, int dtype, const double *chunk_prefix_dev,
               const double *uniforms_dev, int64_t shots, int64_t *idx_dev, void *stream);

/* The same two calls with a table of in-chunk partial sums: sub_prefix_dev (batch * n_chunks * 32 doubles, NULL = none;
 * used when n >= 12) receives every chunk's sequential running sum after each 128 amplitudes -- the values the in-chunk
 * scan passes through anyway -- and tqb_sample2 enters a chunk at the 128-amplitude block that holds the sample, so a
 * shot reads at most 128 amplitudes instead of up to 4096.  Same CDF, same indices, bit for bit.                    */
int tqb_cdf_chunks2(const void *state, int n, int64_t batch, int dtype, double *chunk_prefix_dev,
                    double *sub_prefix_dev, void *stream);
int tqb_sample2(const void *state, int n, int64_t batch, int dtype, const double *chunk_prefix_dev,
                const double *sub_prefix_dev, const double *uniforms_dev, int64_t shots, int64_t *idx_dev,
                void *stream);

/* replaces term_expectation_from_counts / expval_pauli_sum (postprocessing/counts_expval.py:7-20,
 * 87-112) for the shots > 0 VQE path (applications/chem/runtimes/hea_device_runtime.py:120-262):
 * idx_dev holds `batch` rows of `shots` sampled indices (tqb_sample); row b belongs to measurement
 * group b % n_groups, whose terms are term_ptr[g] .. term_ptr[g+1].  zmask = index bits of the
 * term's qubits (qubit q -> bit n-1-q).  <Z_S> = (shots - 2 * #odd-parity samples) / shots, the
 * exact value the reference gets from the counts dict; energy_dev[b] = sum_t coef_t <Z_S_t> in
 * term order; expvals_dev (optional, NULL to skip) [b * expvals_stride + local term index].      */
int tqb_expval_from_samples(const int64_t *idx_dev, int64_t batch, int64_t shots, int n_groups,
                            const int32_t *term_ptr_dev, const uint64_t *term_z_dev,
                            const double *term_coef_dev, double *energy_dev, double *expvals_dev,
                            int64_t expvals_stride, void *stream);

/* ---- sharded states ------------------------------------------------------------------ */
/* The sampler of a state sharded over ranks (rank = highest index bits; engine.py:377-418 at sizes
 * one GPU cannot hold).  Same blocked CDF as above, bit for bit: every rank computes the totals
 * of its own chunks (tqb_chunk_totals: batch * n_chunks doubles), the ranks all-gather them in
 * rank order, every rank runs the sequential prefix over ALL chunks (tqb_chunk_prefix:
 * batch * (n_chunks + 1) doubles) and resolves the uniforms whose chunk it owns
 * (tqb_sample_shard: chunks [chunk_first, chunk_first + 2^n_local / TQB_SCAN_BLOCK) of the
 * n_chunks_total the prefix covers; idx = GLOBAL index, -1 for samples owned by another rank,
 * tail_index for u beyond the last chunk).  A max-reduction over ranks assembles the result.   */
int tqb_chunk_totals(const void *state, int n, int64_t batch, int dtype, double *totals_dev,
                     void *stream);
int tqb_chunk_prefix(const double *totals_dev, int64_t n_chunks, int64_t batch,
                     double *chunk_prefix_dev, void *stream);
int tqb_sample_shard(const void *state, int n_local, int dtype, const double *chunk_prefix_dev,
                     int64_t n_chunks_total, int64_t chunk_first, int64_t tail_index,
                     const double *uniforms_dev, int64_t shots, int64_t *idx_dev, void *stream);
/* Local half of a global<->local qubit exchange: dst/src are chunk views; copies `count`
 * complex elements (used to unpack received blocks back into the shard).                     */
int tqb_copy(void *dst, const void *src, int64_t count, int dtype, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TYXONQ_B200_H */
