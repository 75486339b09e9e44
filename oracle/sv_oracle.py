"""CPU oracle for the TyxonQ statevector hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a plain-numpy restatement of the reference algorithm.  It is the
*checker* for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing
under ``tyxonq_b200/`` imports it, and the product path has no CPU fallback.

Parity pinning: every function here is checked against the live reference
(``/root/reference/src/tyxonq``) by ``tests/golden/make_golden.py`` (run in the build
container, fixtures committed under ``tests/golden/``) and against the known-answer
vectors held by the reference's own tests (``tests/test_oracle_golden.py``).

Conventions (reference ``libs/quantum_library/kernels/statevector.py:28-59``):
qubit ``q`` is tensor axis ``q`` of ``state.reshape((2,)*n)`` == bit ``n-1-q`` of the
flat index (big-endian; qubit 0 is the most significant bit).  All arithmetic is
complex128, every gate is out-of-place.
"""
from __future__ import annotations

import math
from typing import Any, Dict, Iterable, List, Sequence, Tuple

import numpy as np

C128 = np.complex128

# --------------------------------------------------------------------------------------
# gate matrices  (reference: libs/quantum_library/kernels/gates.py)
# --------------------------------------------------------------------------------------
_I2 = np.eye(2, dtype=C128)
_X = np.array([[0, 1], [1, 0]], dtype=C128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=C128)
_Z = np.array([[1, 0], [0, -1]], dtype=C128)
_I4 = np.eye(4, dtype=C128)


def gate_h() -> np.ndarray:  # gates.py:11-16
    return np.array([[1, 1], [1, -1]], dtype=C128) * (C128(1.0) / np.sqrt(np.float64(2.0)))


def gate_rz(theta: float) -> np.ndarray:  # gates.py:19-35  diag(e^{-i t/2}, e^{+i t/2})
    t = np.float64(theta)
    return np.cos(t * 0.5) * _I2 + (-1j * np.sin(t * 0.5)) * _Z


def gate_rx(theta: float) -> np.ndarray:  # gates.py:38-48
    t = np.float64(theta)
    return np.cos(t * 0.5) * _I2 + (-1j * np.sin(t * 0.5)) * _X


def gate_ry(theta: float) -> np.ndarray:  # gates.py:51-65 (real matrix cast to complex)
    t = np.float64(theta)
    c, s = np.cos(t * 0.5), np.sin(t * 0.5)
    return np.array([[c, -s], [s, c]], dtype=np.float64).astype(C128)


def gate_phase(theta: float) -> np.ndarray:  # gates.py:68-79  diag(1, e^{i t})
    return np.array([[1.0, 0.0], [0.0, np.exp(1j * np.float64(theta))]], dtype=C128)


def gate_x() -> np.ndarray:  # gates.py:163-170
    return _X.copy()


def gate_y() -> np.ndarray:  # gates.py:173-177
    return _Y.copy()


def gate_z() -> np.ndarray:  # gates.py:180-184
    return _Z.copy()


def gate_s() -> np.ndarray:  # gates.py:187-188
    return gate_phase(np.pi / 2.0)


def gate_sd() -> np.ndarray:  # gates.py:191-192
    return gate_phase(-np.pi / 2.0)


def gate_t() -> np.ndarray:  # gates.py:195-196
    return gate_phase(np.pi / 4.0)


def gate_td() -> np.ndarray:  # gates.py:199-200
    return gate_phase(-np.pi / 4.0)


def gate_cx_4x4() -> np.ndarray:  # gates.py:82-90, control = first qubit = high bit
    return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=C128)


def gate_cz_4x4() -> np.ndarray:  # gates.py:98-107
    return np.diag(np.array([1, 1, 1, -1], dtype=C128))


def gate_iswap_4x4() -> np.ndarray:  # gates.py:110-133
    return np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]], dtype=C128)


def gate_swap_4x4() -> np.ndarray:  # gates.py:136-160
    return np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=C128)


def _rpp(theta: float, P: np.ndarray) -> np.ndarray:
    t = np.float64(theta)
    return np.cos(t * 0.5) * _I4 + (-1j * np.sin(t * 0.5)) * np.kron(P, P)


def gate_rxx(theta: float) -> np.ndarray:  # gates.py:203-214
    return _rpp(theta, _X)


def gate_ryy(theta: float) -> np.ndarray:  # gates.py:217-233
    return _rpp(theta, _Y)


def gate_rzz(theta: float) -> np.ndarray:  # gates.py:236-252
    return _rpp(theta, _Z)


def gate_cry_4x4(theta: float) -> np.ndarray:  # gates.py:265-284, control first
    t = np.float64(theta)
    c, s = np.cos(t * 0.5), np.sin(t * 0.5)
    return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, c, -s], [0, 0, s, c]], dtype=C128)


def build_controlled_unitary(U: np.ndarray, num_controls: int, ctrl_state: Sequence[int] | None = None) -> np.ndarray:
    """gates.py:287-319: layout [controls..., targets...]; U acts when controls == ctrl_state."""
    U = np.asarray(U, dtype=C128)
    if num_controls < 1:
        return U
    dt = U.shape[0]
    if ctrl_state is None:
        ctrl_state = [1] * num_controls
    want = 0
    for b in ctrl_state:  # ctrl_state[0] is the most significant control bit
        want = (want << 1) | int(b)
    out = np.eye((1 << num_controls) * dt, dtype=C128)
    out[want * dt:(want + 1) * dt, want * dt:(want + 1) * dt] = U
    return out


# --------------------------------------------------------------------------------------
# state kernels (reference: libs/quantum_library/kernels/statevector.py)
# --------------------------------------------------------------------------------------
def init_statevector(n: int) -> np.ndarray:  # statevector.py:19-25
    psi = np.zeros(1 << max(n, 0), dtype=C128)
    psi[0] = 1.0
    return psi


def apply_1q(state: np.ndarray, g: np.ndarray, q: int, n: int) -> np.ndarray:
    """statevector.py:28-42  psi'[..a..] = sum_b g[a,b] psi[..b..] on axis q (einsum)."""
    psi = np.asarray(state, dtype=C128).reshape((2,) * n)
    axes = list(range(2, 2 + n))
    ia = axes.copy(); ia[q] = 1
    oa = axes.copy(); oa[q] = 0
    return np.einsum(np.asarray(g, dtype=C128), [0, 1], psi, ia, oa).reshape(-1)


def apply_2q(state: np.ndarray, g: np.ndarray, q0: int, q1: int, n: int) -> np.ndarray:
    """statevector.py:45-59  gate4 reshaped (2,2,2,2) = [out q0, out q1, in q0, in q1]."""
    if q0 == q1:
        return state
    psi = np.asarray(state, dtype=C128).reshape((2,) * n)
    g4 = np.asarray(g, dtype=C128).reshape(2, 2, 2, 2)
    axes = list(range(4, 4 + n))
    ia = axes.copy(); ia[q0] = 2; ia[q1] = 3
    oa = axes.copy(); oa[q0] = 0; oa[q1] = 1
    return np.einsum(g4, [0, 1, 2, 3], psi, ia, oa).reshape(-1)


def apply_kq(state: np.ndarray, U: np.ndarray, qubits: Sequence[int], n: int) -> np.ndarray:
    """statevector.py:71-129  targets moved last, (2^{n-k},2^k) @ U^T, moved back.
    First listed qubit is the most significant bit of U's row/column index."""
    k = len(qubits)
    if k == 0:
        return state
    psi = np.asarray(state, dtype=C128)
    U = np.asarray(U, dtype=C128).reshape(1 << k, 1 << k)
    tgt = [int(q) for q in qubits]
    keep = [a for a in range(n) if a not in tgt]
    fwd = keep + tgt
    inv = np.argsort(fwd).tolist()
    m = psi.reshape((2,) * n).transpose(fwd).reshape(1 << (n - k), 1 << k)
    out = m @ U.T
    return out.reshape((2,) * n).transpose(inv).reshape(-1)


def expect_z(state: np.ndarray, q: int, n: int) -> float:
    """statevector.py:62-68  sum_{bit_q=0}|psi|^2 - sum_{bit_q=1}|psi|^2."""
    s = np.moveaxis(np.asarray(state).reshape((2,) * n), q, 0).reshape(2, -1)
    sums = np.sum(np.abs(s) ** 2, axis=1)
    return float(sums[0] - sums[1])


def project_z(state: np.ndarray, q: int, keep: int, n: int) -> np.ndarray:
    """engine.py:1075-1087  zero the other half, renormalise (if norm > 0)."""
    t = np.array(state, dtype=C128).reshape((2,) * n)
    t = np.moveaxis(t, q, 0)
    t[1 - int(bool(keep))] = 0  # keep==0 zeroes bit=1, any other value zeroes bit=0
    out = np.moveaxis(t, 0, q).reshape(-1)
    nrm = np.linalg.norm(out)
    return out / nrm if nrm > 0 else out


def apply_kraus(state: np.ndarray, kraus: Sequence[np.ndarray], q: int, n: int, status: float) -> np.ndarray:
    """statevector.py:132-218  pick first i with status <= cumsum(p)_i (else 0), renormalise."""
    ks = [np.asarray(k, dtype=C128) for k in kraus]
    p = []
    for k in ks:
        v = apply_1q(state, k, q, n)
        p.append(np.real(np.sum(np.conj(v) * v)))
    p = np.array(p, dtype=np.float64)
    cum = np.cumsum(p / np.sum(p))
    sel = 0
    for i, c in enumerate(cum):
        if status <= float(c):
            sel = i
            break
    out = apply_1q(state, ks[sel], q, n)
    return out / np.sqrt(np.real(np.sum(np.conj(out) * out)))


# --------------------------------------------------------------------------------------
# op interpreter (reference: devices/simulators/statevector/engine.py)
# --------------------------------------------------------------------------------------
_ONE_Q_FIXED = {"h": gate_h, "x": gate_x, "s": gate_s, "sdg": gate_sd}
_ONE_Q_PARAM = {"rz": gate_rz, "rx": gate_rx, "ry": gate_ry}
_TWO_Q_FIXED = {"cx": gate_cx_4x4, "cz": gate_cz_4x4, "iswap": gate_iswap_4x4, "swap": gate_swap_4x4}
_TWO_Q_PARAM = {"rxx": gate_rxx, "ryy": gate_ryy, "rzz": gate_rzz}


def evolve_ops(n: int, ops: Iterable[Sequence[Any]], *, mode: str = "state",
               initial: np.ndarray | None = None,
               unitary_cache: Dict[str, np.ndarray] | None = None,
               kraus_cache: Dict[str, Sequence[np.ndarray]] | None = None) -> Tuple[np.ndarray, List[int]]:
    """Interpret op tuples exactly like ``StatevectorEngine.run`` (mode="run",
    engine.py:43-374) or ``StatevectorEngine.state`` (mode="state", engine.py:897-1039).

    Quirks kept on purpose: unknown names (y, z, t, tdg, cy, ...) are skipped
    (engine.py:372-374); ``cry`` exists in run() only (engine.py:76-79 vs 918-949);
    run() ignores the initial state; measure_z only records the qubit.
    """
    unitary_cache = unitary_cache or {}
    kraus_cache = kraus_cache or {}
    if mode == "state" and initial is not None:
        psi = np.array(initial, dtype=C128).reshape(-1)
    else:
        psi = init_statevector(n)
    measures: List[int] = []
    for op in ops:
        if not isinstance(op, (list, tuple)) or not op:
            continue
        nm = op[0]
        if nm in _ONE_Q_FIXED:
            psi = apply_1q(psi, _ONE_Q_FIXED[nm](), int(op[1]), n)
        elif nm in _ONE_Q_PARAM:
            psi = apply_1q(psi, _ONE_Q_PARAM[nm](float(op[2])), int(op[1]), n)
        elif nm in _TWO_Q_FIXED:
            psi = apply_2q(psi, _TWO_Q_FIXED[nm](), int(op[1]), int(op[2]), n)
        elif nm in _TWO_Q_PARAM:
            psi = apply_2q(psi, _TWO_Q_PARAM[nm](float(op[3])), int(op[1]), int(op[2]), n)
        elif nm == "cry" and mode == "run":
            psi = apply_2q(psi, gate_cry_4x4(float(op[3])), int(op[1]), int(op[2]), n)
        elif nm == "measure_z" and mode == "run":
            measures.append(int(op[1]))
        elif nm == "project_z":
            psi = project_z(psi, int(op[1]), int(op[2]), n)
        elif nm == "reset":
            psi = project_z(psi, int(op[1]), 0, n)
        elif nm == "unitary":
            if len(op) == 3 and unitary_cache.get(str(op[2])) is not None:
                psi = apply_kq(psi, unitary_cache[str(op[2])], [int(op[1])], n)
            elif len(op) == 4 and unitary_cache.get(str(op[3])) is not None:
                psi = apply_kq(psi, unitary_cache[str(op[3])], [int(op[1]), int(op[2])], n)
        elif nm == "kraus":
            ks = kraus_cache.get(str(op[2]))
            if ks is not None:
                if len(op) <= 3:
                    raise ValueError("oracle needs an explicit kraus status draw")
                psi = apply_kraus(psi, ks, int(op[1]), n, float(op[3]))
        # everything else: silently skipped, like the reference
    return psi, measures


def run_expectations(n: int, ops: Iterable[Sequence[Any]]) -> Dict[str, float]:
    """engine.run(shots=0) result: {"Z{q}": <Z_q>} for each measured qubit (engine.py:467-473)."""
    psi, measures = evolve_ops(n, ops, mode="run")
    return {f"Z{q}": expect_z(psi, q, n) for q in measures}


# --------------------------------------------------------------------------------------
# sampler (reference: engine.py:377-465 + numpy Generator.choice)
# --------------------------------------------------------------------------------------
SCAN_BLOCK = 4096  # amplitudes per scan block: the documented blocked order of the CDF


def probabilities(state: np.ndarray) -> np.ndarray:
    """|psi|^2 in float64 as re*re + im*im (two roundings of the products, one of the sum;
    no FMA) -- the formula the CUDA kernel uses.  The reference takes ``abs(psi)**2``
    (hypot then square, engine.py:379), equal to within 1 ulp per element."""
    s = np.asarray(state)
    re = s.real.astype(np.float64)
    im = s.imag.astype(np.float64)
    return re * re + im * im


def blocked_cdf(p: np.ndarray, block: int = SCAN_BLOCK) -> np.ndarray:
    """Inclusive prefix sum of p in the fixed blocked order used by the CUDA scan:
    sequential float64 sums inside each ``block``-sized chunk, sequential sum of the chunk
    totals, then prefix[chunk] + within[chunk].  (np.cumsum is the block=len(p) case.)"""
    p = np.asarray(p, dtype=np.float64)
    nblk = max(1, (p.size + block - 1) // block)
    pad = nblk * block - p.size
    pp = np.concatenate([p, np.zeros(pad)]) if pad else p
    within = np.cumsum(pp.reshape(nblk, block), axis=1)
    totals = within[:, -1]
    prefix = np.concatenate([[0.0], np.cumsum(totals)[:-1]])
    return (prefix[:, None] + within).reshape(-1)[: p.size]


def sample_indices(p: np.ndarray, uniforms: np.ndarray, block: int = SCAN_BLOCK) -> np.ndarray:
    """idx_s = #{ i : cdf_i / cdf_last <= u_s }  == searchsorted(cdf/cdf[-1], u, 'right')
    (numpy ``Generator.choice(dim, size, p=p)`` with its uniforms made explicit)."""
    cdf = blocked_cdf(p, block)
    cdf = cdf / cdf[-1]
    return np.searchsorted(cdf, np.asarray(uniforms, dtype=np.float64), side="right").astype(np.int64)


def sample_indices_numpy_formula(p: np.ndarray, uniforms: np.ndarray) -> np.ndarray:
    """The reference formula verbatim: p/=p.sum() (engine.py:411-412) then
    cdf=cumsum(p); cdf/=cdf[-1]; searchsorted(u, 'right') (numpy Generator.choice)."""
    p = np.asarray(p, dtype=np.float64)
    p = p / float(p.sum())
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    return cdf.searchsorted(np.asarray(uniforms, dtype=np.float64), side="right").astype(np.int64)


def counts_from_indices(idx: np.ndarray, n: int) -> Dict[str, int]:
    """bincount -> nonzero -> big-endian bitstrings over ALL n qubits (engine.py:418-463)."""
    out: Dict[str, int] = {}
    vals, cnts = np.unique(np.asarray(idx, dtype=np.int64), return_counts=True)
    for v, c in zip(vals.tolist(), cnts.tolist()):
        out[format(v, f"0{n}b") if n > 0 else ""] = int(c)
    return out


# --------------------------------------------------------------------------------------
# Pauli sums (reference: kernels/pauli.py:53-87, dynamics.py:117-126, counts_expval.py)
# --------------------------------------------------------------------------------------
_PAULI = (_I2, _X, _Y, _Z)


def pauli_string_to_matrix(ps: Sequence[int]) -> np.ndarray:  # pauli.py:65-71
    out = np.array([[1.0 + 0.0j]], dtype=C128)
    for code in ps:
        out = np.kron(out, _PAULI[int(code)])
    return out


def pauli_sum_dense(terms: Sequence[Sequence[int]], weights: Sequence[float] | None = None) -> np.ndarray:
    """pauli.py:74-87 (dense 2^n x 2^n; small n only)."""
    n = len(terms[0])
    H = np.zeros((1 << n, 1 << n), dtype=C128)
    w = [1.0] * len(terms) if weights is None else weights
    for ps, c in zip(terms, w):
        H = H + float(c) * pauli_string_to_matrix(ps)
    return H


def apply_pauli_string(state: np.ndarray, ps: Sequence[int]) -> np.ndarray:
    """P|psi> matrix-free (same result as pauli_string_to_matrix(ps) @ psi)."""
    n = len(ps)
    out = np.asarray(state, dtype=C128)
    for q, code in enumerate(ps):
        if int(code):
            out = apply_1q(out, _PAULI[int(code)], q, n)
    return out


def expect_pauli_sum(state: np.ndarray, terms: Sequence[Sequence[int]], weights: Sequence[float]) -> float:
    """Re <psi| sum_j w_j P_j |psi>  (dynamics.py:117-126 with H never materialised)."""
    psi = np.asarray(state, dtype=C128)
    e = 0.0
    for ps, w in zip(terms, weights):
        e += float(w) * float(np.real(np.vdot(psi, apply_pauli_string(psi, ps))))
    return e


def apply_pauli_sum(state: np.ndarray, terms: Sequence[Sequence[int]], weights: Sequence[complex]) -> np.ndarray:
    psi = np.asarray(state, dtype=C128)
    out = np.zeros_like(psi)
    for ps, w in zip(terms, weights):
        out = out + complex(w) * apply_pauli_string(psi, ps)
    return out


def term_expectation_from_counts(counts: Dict[str, int], idxs: Sequence[int]) -> float:
    """counts_expval.py:7-20."""
    tot = sum(counts.values()) or 1
    acc = 0.0
    for b, c in counts.items():
        sgn = 1.0
        for q in idxs:
            sgn *= 1.0 if b[q] == "0" else -1.0
        acc += sgn * c
    return acc / tot


def zproduct_from_probabilities(p: np.ndarray, idxs: Sequence[int], n: int) -> float:
    """counts_expval.py:64-76 (vectorised): sum_i p_i * prod_q (-1)^{bit_{n-1-q}(i)}."""
    i = np.arange(1 << n, dtype=np.int64)
    sgn = np.ones(1 << n)
    for q in idxs:
        sgn = sgn * (1.0 - 2.0 * ((i >> (n - 1 - q)) & 1))
    return float(np.sum(sgn * np.asarray(p, dtype=np.float64)))


def heisenberg_terms(n: int, edges: Sequence[Tuple[int, int]], *, hzz=1.0, hxx=1.0, hyy=1.0,
                     hz=0.0, hx=0.0, hy=0.0) -> Tuple[List[List[int]], List[float]]:
    """Term enumeration of pauli.py:145-171 (pairs: ZZ, XX, YY per edge; then fields Z, X, Y per qubit)."""
    terms: List[List[int]] = []
    w: List[float] = []
    for a, b in edges:
        for code, c in ((3, hzz), (1, hxx), (2, hyy)):
            if c != 0.0:
                ps = [0] * n; ps[a] = code; ps[b] = code
                terms.append(ps); w.append(c)
    for q in range(n):
        for code, c in ((3, hz), (1, hx), (2, hy)):
            if c != 0.0:
                ps = [0] * n; ps[q] = code
                terms.append(ps); w.append(c)
    return terms, w


# --------------------------------------------------------------------------------------
# gradients (reference: kernels/common.py:11-23, numpy_backend.py:386-454)
# --------------------------------------------------------------------------------------
def parameter_shift_gradient(energy_fn, params: Sequence[float]) -> np.ndarray:  # common.py:11-23
    base = np.asarray(params, dtype=np.float64)
    g = np.zeros_like(base)
    for i in range(base.size):
        pp = base.copy(); pp[i] += 0.5 * np.pi
        pm = base.copy(); pm[i] -= 0.5 * np.pi
        g[i] = 0.5 * (float(energy_fn(pp)) - float(energy_fn(pm)))
    return g


def central_fd_gradient(energy_fn, params: Sequence[float], eps: float = 1e-6) -> np.ndarray:
    base = np.asarray(params, dtype=np.float64)
    g = np.zeros_like(base)
    for i in range(base.size):
        pp = base.copy(); pp[i] += eps
        pm = base.copy(); pm[i] -= eps
        g[i] = (float(energy_fn(pp)) - float(energy_fn(pm))) / (2 * eps)
    return g


# --------------------------------------------------------------------------------------
# workload generators (op sequences of the reference's circuit builders)
# --------------------------------------------------------------------------------------
def hea_ops(n: int, nlayers: int, params: np.ndarray) -> List[tuple]:
    """libs/circuits_library/blocks.py:14-56 (example_block): h on all; per layer the cx chain,
    then rz(q), rx(q) interleaved per qubit; params flat [layer][rz(n) | rx(n)]."""
    flat = np.asarray(params, dtype=np.float64).reshape(-1)
    ops: List[tuple] = [("h", q) for q in range(n)]
    for j in range(nlayers):
        base = j * 2 * n
        ops += [("cx", q, q + 1) for q in range(n - 1)]
        for q in range(n):
            ops.append(("rz", q, float(flat[base + q])))
            ops.append(("rx", q, float(flat[base + n + q])))
    return ops


def hwe_ry_ops(n: int, nlayers: int, params: np.ndarray) -> List[tuple]:
    """libs/circuits_library/blocks.py:60-85 (build_hwe_ry_ops), barriers dropped (no-ops)."""
    mat = np.asarray(params, dtype=np.float64).reshape(nlayers + 1, n)
    ops: List[tuple] = [("ry", i, float(mat[0, i])) for i in range(n)]
    for l in range(nlayers):
        ops += [("cx", i, i + 1) for i in range(n - 1)]
        ops += [("ry", i, float(mat[l + 1, i])) for i in range(n)]
    return ops


def qaoa_ring_ops(n: int, nlayers: int, params: np.ndarray) -> List[tuple]:
    """libs/circuits_library/qaoa_ising.py:8-66 with ZZ terms on a ring, weights 1, mixer X."""
    p = np.asarray(params, dtype=np.float64).reshape(-1)
    ops: List[tuple] = [("h", q) for q in range(n)]
    edges = [(i, (i + 1) % n) for i in range(n)] if n > 2 else [(0, 1)]
    for j in range(nlayers):
        for a, b in edges:
            q0, q1 = min(a, b), max(a, b)  # one_indices come out ascending (qaoa_ising.py:31-36)
            ops.append(("rzz", q0, q1, float(1.0 * p[2 * j])))
        for q in range(n):
            ops.append(("rx", q, float(p[2 * j + 1])))
    return ops


def _trotter_term(ps: Sequence[int], theta: float) -> List[tuple]:
    """libs/circuits_library/trotter_circuit.py:8-72 (_apply_single_term)."""
    nz = [i for i, v in enumerate(ps) if v != 0]
    ops: List[tuple] = []
    if not nz:
        return ops
    if len(nz) == 1:
        q = nz[0]
        if ps[q] == 1:
            return [("h", q), ("rz", q, 2.0 * theta), ("h", q)]
        if ps[q] == 2:
            return [("sdg", q), ("h", q), ("rz", q, 2.0 * theta), ("h", q), ("s", q)]
        return [("rz", q, 2.0 * theta)]
    for q in nz:
        if ps[q] == 1:
            ops.append(("h", q))
        elif ps[q] == 2:
            ops += [("sdg", q), ("h", q)]
    for i in range(len(nz) - 1):
        ops.append(("cx", nz[i], nz[i + 1]))
    ops.append(("rz", nz[-1], 2.0 * theta))
    for i in range(len(nz) - 2, -1, -1):
        ops.append(("cx", nz[i], nz[i + 1]))
    for q in reversed(nz):
        if ps[q] == 1:
            ops.append(("h", q))
        elif ps[q] == 2:
            ops += [("h", q), ("s", q)]
    return ops


def trotter_ops(terms: Sequence[Sequence[int]], weights: Sequence[float], time: float, steps: int) -> List[tuple]:
    """libs/circuits_library/trotter_circuit.py:75-122 (first order; ends with measure_z on all)."""
    n = len(terms[0])
    dt = float(time) / float(max(1, int(steps)))
    ops: List[tuple] = []
    for _ in range(max(1, int(steps))):
        for ps, c in zip(terms, weights):
            ops += _trotter_term(ps, float(c) * dt)
    ops += [("measure_z", q) for q in range(n)]
    return ops


def tfim_terms(n: int, J: float = 1.0, h: float = 1.0) -> Tuple[List[List[int]], List[float]]:
    """TFIM Pauli list for config 4: ZZ on (i,i+1) weight J, X on i weight h."""
    terms: List[List[int]] = []
    w: List[float] = []
    for i in range(n - 1):
        ps = [0] * n; ps[i] = 3; ps[i + 1] = 3
        terms.append(ps); w.append(J)
    for i in range(n):
        ps = [0] * n; ps[i] = 1
        terms.append(ps); w.append(h)
    return terms, w


def tfim_vqe_ops(n: int, nlayers: int, param: np.ndarray) -> List[tuple]:
    """examples/vqetfim_benchmark.py:21-34 (ansatz_ops_xx_rz): rxx chain then rz on each wire."""
    p = np.asarray(param, dtype=np.float64).reshape(2 * nlayers, n)
    ops: List[tuple] = []
    t = 0
    for _ in range(nlayers):
        ops += [("rxx", i, i + 1, float(p[t, i])) for i in range(n - 1)]
        t += 1
        ops += [("rz", i, float(p[t, i])) for i in range(n)]
        t += 1
    return ops


def tfim_vqe_energy(n: int, nlayers: int, param: np.ndarray, Jx: float = 1.0, h: float = -1.0) -> float:
    """examples/vqetfim_benchmark.py:70-103 (exact_energy): h*sum<Z_i> + Jx*sum<X_i X_{i+1}>."""
    psi, _ = evolve_ops(n, tfim_vqe_ops(n, nlayers, param))
    e = 0.0
    for i in range(n):
        e += h * expect_z(psi, i, n)
    psix = psi
    for q in range(n):
        psix = apply_1q(psix, gate_h(), q, n)
    p = np.abs(psix) ** 2
    for i in range(n - 1):
        e += Jx * zproduct_from_probabilities(p, [i, i + 1], n)
    return float(e)
