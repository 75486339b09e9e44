"""Host-side gate definitions: the parity contract of the reference's gate constructors.

Every matrix here equals the reference's ``libs/quantum_library/kernels/gates.py`` output
(complex128) for the same angle; tests/test_gates_parity.py pins that against golden
fixtures generated from the live reference.  Matrices are built on the host with numpy
(2x2 / 4x4, a few hundred bytes per gate) and shipped to the device in one buffer per
circuit; the 2^n-sized work happens only in the CUDA kernels.

A gate is lowered to one of three device kinds (include/tyxonq_b200.h):
  DENSE  full 2^k x 2^k matrix (h, rx, ry, rxx, ryy, user unitaries)
  DIAG   2^k-entry table       (rz, s, sdg, phase, cz, rzz) -- needs no qubit to be tile-local
  PAIR   2x2 block on two basis patterns, identity elsewhere (cry, iswap, UCC excitation
         rotations) -- touches only 2/2^k of the amplitudes
  SWAP   PAIR with the matrix X: a pure exchange of two patterns (x, cx, swap), no arithmetic
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

C128 = np.complex128
_SQ2 = 1.0 / math.sqrt(2.0)

# ---------------------------------------------------------------------------------------
# matrices (reference gates.py line numbers in comments)
# ---------------------------------------------------------------------------------------
H_MAT = np.array([[1, 1], [1, -1]], dtype=C128) * C128(1.0 / np.sqrt(np.float64(2.0)))  # :11-16
X_MAT = np.array([[0, 1], [1, 0]], dtype=C128)       # :163-170
Y_MAT = np.array([[0, -1j], [1j, 0]], dtype=C128)    # :173-177
Z_MAT = np.array([[1, 0], [0, -1]], dtype=C128)      # :180-184
CX_MAT = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=C128)   # :82-90
CZ_MAT = np.diag(np.array([1, 1, 1, -1], dtype=C128))                                      # :98-107
ISWAP_MAT = np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]], dtype=C128)  # :110-133
SWAP_MAT = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=C128)     # :136-160


def _cs(theta: float) -> Tuple[float, float]:
    t = float(theta) * 0.5
    return math.cos(t), math.sin(t)


def rz_diag(theta: float) -> np.ndarray:      # :19-35  c*I - i s Z
    c, s = _cs(theta)
    return np.array([complex(c, -s), complex(c, s)], dtype=C128)


def rx_mat(theta: float) -> np.ndarray:       # :38-48  c*I - i s X
    c, s = _cs(theta)
    return np.array([[c, -1j * s], [-1j * s, c]], dtype=C128)


def ry_mat(theta: float) -> np.ndarray:       # :51-65
    c, s = _cs(theta)
    return np.array([[c, -s], [s, c]], dtype=C128)


def phase_diag(theta: float) -> np.ndarray:   # :68-79  diag(1, e^{i theta})
    return np.array([1.0, np.exp(1j * np.float64(theta))], dtype=C128)


S_DIAG = phase_diag(np.pi / 2.0)     # :187-188
SD_DIAG = phase_diag(-np.pi / 2.0)   # :191-192


def rzz_diag(theta: float) -> np.ndarray:     # :236-252  c*I4 - i s Z(x)Z
    c, s = _cs(theta)
    a, b = complex(c, -s), complex(c, s)
    return np.array([a, b, b, a], dtype=C128)


def rxx_mat(theta: float) -> np.ndarray:      # :203-214
    c, s = _cs(theta)
    m = np.zeros((4, 4), dtype=C128)
    for i in range(4):
        m[i, i] = c
        m[i, 3 - i] = -1j * s
    return m


def ryy_mat(theta: float) -> np.ndarray:      # :217-233  Y(x)Y = antidiag(-1, 1, 1, -1)
    c, s = _cs(theta)
    m = np.zeros((4, 4), dtype=C128)
    yy = (-1.0, 1.0, 1.0, -1.0)
    for i in range(4):
        m[i, i] = c
        m[i, 3 - i] = -1j * s * yy[i]
    return m


# generators d/dtheta U = G U for the parametrised gates (used by the adjoint sweep)
def _gen_rot(P: np.ndarray) -> np.ndarray:
    return (-0.5j) * P


GEN = {
    "rx": _gen_rot(X_MAT), "ry": _gen_rot(Y_MAT), "rz": _gen_rot(Z_MAT),
    "rxx": _gen_rot(np.kron(X_MAT, X_MAT)), "ryy": _gen_rot(np.kron(Y_MAT, Y_MAT)),
    "rzz": _gen_rot(np.kron(Z_MAT, Z_MAT)),
    # cry: |1><1| (x) (-i/2) Y
    "cry": np.kron(np.array([[0, 0], [0, 1]], dtype=C128), _gen_rot(Y_MAT)),
}

# ---------------------------------------------------------------------------------------
# lowered gates
# ---------------------------------------------------------------------------------------
DENSE, DIAG, PAIR, SWAP, MUX, CHAIN = 0, 1, 2, 3, 4, 5


@dataclass
class LGate:
    """A gate lowered to index-bit space.

    bits[j] is the index bit carrying matrix-index bit j (j = 0 least significant), i.e. for
    reference qubits (q0, .., qk-1) on n qubits: bits[j] = n-1-q_{k-1-j}.
    data: DENSE (2^k,2^k) | DIAG (2^k,) | PAIR (4,) or (8,) [even-parity 2x2, odd-parity 2x2]
          | MUX (8,) [2x2 for control = 0, 2x2 for control = 1], bits = (target, control).
    mask = all index bits the gate touches (ordering); local_mask = the bits that must be inside
    the shared-memory tile (none for DIAG, only the target for MUX).
    """
    kind: int
    bits: Tuple[int, ...]
    data: np.ndarray
    pat_a: int = 0
    pat_b: int = 0
    zmask: int = 0
    # bookkeeping for gradients: (parameter slot, op name) of a parametrised gate
    param: Optional[int] = None
    name: str = ""
    batched: bool = False   # data has a leading axis: one matrix per batch member
    mask: int = field(default=0, init=False)
    local_mask: int = field(default=0, init=False)

    def __post_init__(self) -> None:
        m = 0
        for b in self.bits:
            m |= 1 << int(b)
        self.mask = m
        if self.kind == DIAG:
            self.local_mask = 0
        elif self.kind == MUX:
            self.local_mask = 1 << int(self.bits[0])
        elif self.kind == CHAIN:   # bits = targets in layer order (+ outer control when pat_a & 1) (+ pat_a >> 1 extra table bits)
            nt = len(self.bits) - (self.pat_a & 1) - (self.pat_a >> 1)
            self.local_mask = 0
            for b in self.bits[:nt]:
                self.local_mask |= 1 << int(b)
        else:
            self.local_mask = m

    @property
    def k(self) -> int:
        return len(self.bits)


def _bits_of(qubits: Sequence[int], n: int) -> Tuple[int, ...]:
    return tuple(n - 1 - int(q) for q in reversed(list(qubits)))


def dense_gate(mat: np.ndarray, qubits: Sequence[int], n: int, **kw: Any) -> LGate:
    k = len(qubits)
    d = 1 << k
    m = np.asarray(mat, dtype=C128)
    if m.size == d * d:
        return LGate(DENSE, _bits_of(qubits, n), m.reshape(d, d), **kw)
    return LGate(DENSE, _bits_of(qubits, n), m.reshape(-1, d * d), batched=True, **kw)  # leading axis = batch


def diag_gate(tab: np.ndarray, qubits: Sequence[int], n: int, **kw: Any) -> LGate:
    return LGate(DIAG, _bits_of(qubits, n), np.asarray(tab, dtype=C128).reshape(-1), **kw)


def pair_gate(m2: np.ndarray, qubits: Sequence[int], n: int, pat_a: int, pat_b: int, *,
              m2_odd: Optional[np.ndarray] = None, zmask: int = 0, **kw: Any) -> LGate:
    d = np.asarray(m2, dtype=C128).reshape(4)
    if zmask:
        d = np.concatenate([d, np.asarray(m2_odd, dtype=C128).reshape(4)])
    return LGate(PAIR, _bits_of(qubits, n), d, pat_a=int(pat_a), pat_b=int(pat_b), zmask=int(zmask), **kw)


def swap_gate(qubits: Sequence[int], n: int, pat_a: int, pat_b: int, **kw: Any) -> LGate:
    """Exchange of two basis patterns (x, cx, swap): a PAIR gate with matrix X, done without arithmetic."""
    return LGate(SWAP, _bits_of(qubits, n), X_MAT.reshape(4).copy(), pat_a=int(pat_a), pat_b=int(pat_b), **kw)


MUX_GENERAL, MUX_XU, LAYER_PLAIN = 0, 2, 3   # how a chain layer was BUILT: arbitrary pair | u1 = X.u0 (cx after the gate) | u1 = u0


def mux_gate(u0: np.ndarray, u1: np.ndarray, target_bit: int, control_bit: int, structure: int = MUX_GENERAL, **kw: Any) -> LGate:
    """1-qubit gate on index bit ``target_bit``: u0 where index bit ``control_bit`` is 0, u1 where it is 1.
    ``structure`` (kept in ``pat_b``) records that the pair is a gate followed by a fused cx (u1 = X.u0)."""
    u0 = np.asarray(u0, dtype=C128)
    u1 = np.asarray(u1, dtype=C128)
    if u0.size == 4 and u1.size == 4:
        d = np.concatenate([u0.reshape(4), u1.reshape(4)])
        return LGate(MUX, (int(target_bit), int(control_bit)), d, pat_b=int(structure), **kw)
    B = max(u0.size, u1.size) // 4   # one matrix pair per batch member
    d = np.concatenate([np.broadcast_to(u0.reshape(-1, 4), (B, 4)), np.broadcast_to(u1.reshape(-1, 4), (B, 4))], axis=1)
    return LGate(MUX, (int(target_bit), int(control_bit)), np.ascontiguousarray(d), pat_b=int(structure), batched=True, **kw)


# A scaled rotation layer runs as x + t.(i x') with t = r / a as long as |t| <= ROT_T_MAX, else as c.x + i x' with c = a / r
# (tqb_core.cuh rot_layer_scaled).  The product of the divided-out factors multiplies layer 0, so the rounding errors stay
# relative to |a x| + |r x'| in either form; the t form is preferred far beyond |t| = 1 because the specialised kernels
# compile the form in (tqb_gate.off_b bits 8..11) and circuits whose angles move (VQE) should keep their kernel shapes.
ROT_T_MAX = 1024.0


def rot_decompose(u: np.ndarray) -> Dict[int, Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]]:
    """U = M.diag(d0, d1) with M = [[a, i r], [i r, a]] (type 0: rz.rx products) or M = [[a, -r], [r, a]] (type 1: ry, h),
    a and r real, for every member of u [B, 2, 2].  Returns {type: (a, r, d0, d1)} for the types that reproduce U to
    1e-13 on ALL members (tqb_core.cuh gate_chain_rot runs such layers with half the multiply-adds)."""
    u = np.asarray(u, dtype=C128).reshape(-1, 2, 2)
    if u.shape[0] == 1:
        return _rot_decompose_scalar(*[complex(x) for x in u.reshape(4)])
    u00, u01, u10, u11 = u[:, 0, 0], u[:, 0, 1], u[:, 1, 0], u[:, 1, 1]
    av, rv = np.abs(u00), np.abs(u10)
    big = av >= rv
    sa, sr = np.where(av > 0, av, 1.0), np.where(rv > 0, rv, 1.0)
    out: Dict[int, Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]] = {}
    first = _rot_decompose_scalar(*[complex(x) for x in u[0].reshape(4)])   # a type that fails on member 0 fails on the batch
    for typ, f10, f01 in ((0, 1j, 1j), (1, 1.0, -1.0)):      # M10 = f10 * r, M01 = f01 * r
        if typ not in first:
            continue
        d0 = np.where(big, u00 / sa, u10 / (f10 * sr))
        d1 = np.where(big, u11 / sa, u01 / (f01 * sr))
        neg = (np.abs(d0 + 1.0) < 1e-14) & (np.abs(d1 + 1.0) < 1e-14)   # M.(-1) = (-M).1: keep the table a unit table
        d0, d1 = np.where(neg, 1.0 + 0.0j, d0), np.where(neg, 1.0 + 0.0j, d1)
        with np.errstate(all="ignore"):                       # a zero d0 fails the |d| = 1 check below
            a = (u00 / d0).real
            r = (u10 / (f10 * d0)).real
        rec = np.stack([a * d0, f01 * r * d1, f10 * r * d0, a * d1], axis=1)
        if np.all(np.abs(np.abs(d0) - 1) < 1e-13) and np.all(np.abs(np.abs(d1) - 1) < 1e-13) and \
                np.abs(rec - u.reshape(-1, 4)).max() < 1e-13:
            out[typ] = (a, r, d0, d1)
    return out


def _rot_decompose_scalar(u00: complex, u01: complex, u10: complex, u11: complex):
    """rot_decompose for one matrix in plain Python complex arithmetic (the planner calls it once per gate)."""
    av, rv = abs(u00), abs(u10)
    out = {}
    for typ, f10, f01 in ((0, 1j, 1j), (1, 1.0, -1.0)):
        if av >= rv:
            if av == 0.0:
                continue
            d0, d1 = u00 / av, u11 / av
        else:
            d0, d1 = u10 / (f10 * rv), u01 / (f01 * rv)
        if abs(abs(d0) - 1) > 1e-13 or abs(abs(d1) - 1) > 1e-13:
            continue
        if abs(d0 + 1.0) < 1e-14 and abs(d1 + 1.0) < 1e-14:   # M.(-1) = (-M).1: keep the table a unit table
            d0 = d1 = 1.0 + 0.0j
        a = (u00 / d0).real
        r = (u10 / (f10 * d0)).real
        if max(abs(a * d0 - u00), abs(f01 * r * d1 - u01), abs(f10 * r * d0 - u10), abs(a * d1 - u11)) < 1e-13:
            out[typ] = (np.array([a]), np.array([r]), np.array([d0], dtype=C128), np.array([d1], dtype=C128))
    return out


def rot_plan(decs: Sequence[Dict[int, Any]], tags: Sequence[int], has_control: bool):
    """(type, muxed, decompositions) when a chain whose selector-0 matrices decompose as ``decs`` (rot_decompose) with
    structure ``tags`` can run in rotation form (tqb_core.cuh gate_chain_rot), else None."""
    tags = [int(t) for t in tags]
    rest = set(tags[1:])
    if tags[0] != (MUX_XU if has_control else LAYER_PLAIN) or len(rest) != 1 or not rest <= {MUX_XU, LAYER_PLAIN}:
        return None
    for typ in (0, 1):
        if all(typ in d for d in decs):
            return typ, (1 if tags[1] == MUX_XU else 0), [d[typ] for d in decs]
    return None


def chain_gate(layers: Sequence[Tuple[int, np.ndarray, np.ndarray]], control_bit: Optional[int] = None,
               structure: Optional[Sequence[int]] = None, decs: Optional[Sequence[Dict[int, Any]]] = None,
               pre_diags: Sequence["LGate"] = (), extra_bits: Sequence[int] = (), **kw: Any) -> LGate:
    """R = 2..4 one-qubit layers [(target_bit, M_sel0, M_sel1), ...]: layer 0 is selected by
    ``control_bit`` (None: M_sel0 is used), layer i > 0 by the value of layer i-1's target bit.
    ``structure``: per layer MUX_GENERAL / MUX_XU / LAYER_PLAIN.  When every layer is plain or a gate followed by a
    fused cx, and all selector-0 matrices factor as rotation x diagonal of one type (``rot_decompose``), the gate is
    emitted in rotation form: data = [P (2^R entries), S.(a_0 + i r_0), (t_i or c_i) + i mode_i for layers i > 0]
    (tqb_core.cuh rot_layer_scaled: the larger of a_i, r_i is divided out and collected in S), pat_b = 4 + 2*type + muxed
    (R = 4 exists in rotation form only).  ``pre_diags``: diagonal gates that act right BEFORE the chain, on chain
    targets plus the ``extra_bits`` (<= 2): they are multiplied into the table, which then has 2^(R+E) entries
    (index = register index + extras << R); the extras are appended to ``bits`` and counted in pat_a >> 1."""
    assert 2 <= len(layers) <= 4
    R = len(layers)
    bits = [int(t) for t, _, _ in layers]
    B = max(max(np.asarray(a).size, np.asarray(b).size) for _, a, b in layers) // 4
    if control_bit is not None:
        bits.append(int(control_bit))
    if structure is not None and decs is None:
        decs = [rot_decompose(a) for _, a, _ in layers]
    rot = rot_plan(decs, structure, control_bit is not None) if structure is not None else None
    assert R <= 3 or rot is not None, "4-layer chains need the rotation form"
    assert not pre_diags or (rot is not None and B == 1), "diagonal gates can only be absorbed by unbatched rotation-form chains"
    E = len(extra_bits)
    assert E <= 2
    unit = False
    inv_bits = 0   # bit i-1: layer i runs in the c form; bit 3: the bits are valid; bit 4: layer 0 is scaled too (factor in the table)
    if rot is not None:
        typ, muxed, dec = rot
        if B == 1:   # plain Python complex arithmetic: the planner builds one of these per chain
            tab = [1.0 + 0.0j]
            for i in range(R):   # register index bit i <-> layer i
                d0, d1 = complex(dec[i][2][0]), complex(dec[i][3][0])
                tab = [p * d0 for p in tab] + [p * d1 for p in tab]
            if E or pre_diags:
                tab = tab * (1 << E)                      # index = register index + (extras << R)
                targets = bits[:R]
                for d in pre_diags:
                    src = [("t", targets.index(b)) if b in targets else ("x", list(extra_bits).index(b)) for b in d.bits]
                    for idx in range(len(tab)):
                        sv, xv = idx & ((1 << R) - 1), idx >> R
                        di = 0
                        for j, (kind, pos) in enumerate(src):
                            di |= (((sv >> pos) if kind == "t" else (xv >> pos)) & 1) << j
                        tab[idx] *= complex(d.data[di])
            unit = all(abs(t - 1.0) < 1e-15 for t in tab)   # (phases of exactly 1 up to the rounding of their normalisation)
            if control_bit is not None:   # second copy for control = 1: register-index bit 0 flipped (gate_chain_rot)
                tab = tab + [tab[i ^ 1] for i in range(len(tab))]
            coef, scale = [], 1.0
            for i in range(1, R):   # scaled layers: (t, 0) with t = r / a, or (c, 1) with c = a / r; the factor goes to layer 0
                ai, ri = float(dec[i][0][0]), float(dec[i][1][0])
                if abs(ai) * ROT_T_MAX >= abs(ri):
                    coef.append(complex(ri / ai, 0.0))
                    scale *= ai
                else:
                    coef.append(complex(ai / ri, 1.0))
                    scale *= ri
                    inv_bits |= 1 << (i - 1)
            inv_bits |= 8     # the forms are the same for every state of the batch: the specialised kernels compile them in
            a0, r0 = float(dec[0][0][0]), float(dec[0][1][0])
            if not unit and a0 != 0.0 and abs(a0) * ROT_T_MAX >= abs(r0):
                # the table is multiplied in anyway: it takes layer 0's factor as well, and layer 0 runs in the t form
                # (2 multiply-adds per amplitude instead of 4)
                tab = [p * (a0 * scale) for p in tab]
                tab += [complex(r0 / a0, 0.0)] + coef
                inv_bits |= 16
            else:
                tab += [complex(a0 * scale, r0 * scale)] + coef
            data = np.array(tab, dtype=C128).reshape(1, -1)
        else:
            P = np.ones((B, 1 << R), dtype=C128)
            if not all(np.all(dec[i][2] == 1.0) and np.all(dec[i][3] == 1.0) for i in range(R)):   # (ry layers: all ones)
                for s in range(1 << R):
                    for i in range(R):
                        P[:, s] *= dec[i][3] if (s >> i) & 1 else dec[i][2]
            unit = bool(np.abs(P - 1.0).max() < 1e-15)   # plain rotations (ry, rx layers): the kernel skips the table multiply
            if control_bit is not None:
                P = np.concatenate([P, P[:, np.arange(1 << R) ^ 1]], axis=1)
            cols, scale, all_t = [], np.ones(B), True
            for i in range(1, R):   # scaled layers, per batch member (see the B == 1 branch)
                ai, ri = np.broadcast_to(dec[i][0], (B,)), np.broadcast_to(dec[i][1], (B,))
                big = np.abs(ai) * ROT_T_MAX >= np.abs(ri)
                all_t = all_t and bool(np.all(big))
                with np.errstate(all="ignore"):
                    cols.append(np.where(big, ri / ai, ai / ri) + 1j * np.where(big, 0.0, 1.0))
                scale = scale * np.where(big, ai, ri)
            if all_t:
                inv_bits = 8   # every member runs every scaled layer in the t form
            cols.insert(0, (np.broadcast_to(dec[0][0], (B,)) + 1j * np.broadcast_to(dec[0][1], (B,))) * scale)
            coef = np.stack(cols, axis=1)
            data = np.ascontiguousarray(np.concatenate([P, coef], axis=1))
        kw["pat_b"] = (4 + 2 * typ + muxed) | (E << 4) | (128 if unit else 0) | (inv_bits << 8)
    else:
        blocks = []
        for _, a, b in layers:
            for m in (a, b):
                blocks.append(np.broadcast_to(np.asarray(m, dtype=C128).reshape(-1, 4), (B, 4)))
        data = np.ascontiguousarray(np.concatenate(blocks, axis=1))
    bits += [int(b) for b in extra_bits]
    pat_a = (1 if control_bit is not None else 0) | (E << 1)
    if B == 1:
        return LGate(CHAIN, tuple(bits), data.reshape(-1), pat_a=pat_a, **kw)
    return LGate(CHAIN, tuple(bits), data, pat_a=pat_a, batched=True, **kw)


def classify_unitary(mat: np.ndarray, qubits: Sequence[int], n: int) -> LGate:
    """Lower a user-supplied matrix (``unitary`` op / apply_kqubit_unitary) structurally."""
    k = len(qubits)
    d = 1 << k
    M = np.asarray(mat, dtype=C128).reshape(d, d)
    off = M - np.diag(np.diag(M))
    if not np.any(off):
        return diag_gate(np.diag(M).copy(), qubits, n)
    if k >= 2:
        rows = [i for i in range(d) if np.any(off[i]) or np.any(off[:, i]) or M[i, i] != 1.0]
        if len(rows) == 2:
            a, b = rows
            blk = M[np.ix_([a, b], [a, b])]
            if np.array_equal(blk, X_MAT):
                return swap_gate(qubits, n, a, b)
            return pair_gate(blk, qubits, n, a, b)
    return dense_gate(M, qubits, n)


def lower_op(op: Sequence[Any], n: int, *, mode: str, unitary_cache: Optional[Dict[str, Any]] = None,
             param: Optional[int] = None) -> Optional[LGate]:
    """Lower one op tuple.  Returns None for ops that are not gates or that the reference
    engine silently skips (engine.py:372-374; ``cry`` is skipped by state(), engine.py:918-949)."""
    nm = op[0]
    if nm == "h":
        return dense_gate(H_MAT, [op[1]], n, name=nm)
    if nm == "x":
        return swap_gate([op[1]], n, 0, 1, name=nm)
    if nm == "rx":
        return dense_gate(rx_mat(op[2]), [op[1]], n, name=nm, param=param)
    if nm == "ry":
        return dense_gate(ry_mat(op[2]), [op[1]], n, name=nm, param=param)
    if nm == "rz":
        return diag_gate(rz_diag(op[2]), [op[1]], n, name=nm, param=param)
    if nm == "s":
        return diag_gate(S_DIAG, [op[1]], n, name=nm)
    if nm == "sdg":
        return diag_gate(SD_DIAG, [op[1]], n, name=nm)
    if nm in ("cx", "cz", "iswap", "swap", "rxx", "ryy", "rzz", "cry"):
        q0, q1 = int(op[1]), int(op[2])
        if q0 == q1:  # apply_2q_statevector returns the input unchanged (statevector.py:46-47)
            return None
        if nm == "cx":
            return swap_gate([q0, q1], n, 0b10, 0b11, name=nm)
        if nm == "cz":
            return diag_gate(np.array([1, 1, 1, -1], dtype=C128), [q0, q1], n, name=nm)
        if nm == "swap":
            return swap_gate([q0, q1], n, 0b01, 0b10, name=nm)
        if nm == "iswap":
            return pair_gate(np.array([[0, 1j], [1j, 0]], dtype=C128), [q0, q1], n, 0b01, 0b10, name=nm)
        if nm == "rzz":
            return diag_gate(rzz_diag(op[3]), [q0, q1], n, name=nm, param=param)
        if nm == "rxx":
            return dense_gate(rxx_mat(op[3]), [q0, q1], n, name=nm, param=param)
        if nm == "ryy":
            return dense_gate(ryy_mat(op[3]), [q0, q1], n, name=nm, param=param)
        if nm == "cry":
            if mode != "run":
                return None
            return pair_gate(ry_mat(op[3]), [q0, q1], n, 0b10, 0b11, name=nm, param=param)
    if nm == "unitary":
        cache = unitary_cache or {}
        if len(op) == 3:
            mat = cache.get(str(op[2]))
            return None if mat is None else classify_unitary(_to_np(mat), [int(op[1])], n)
        if len(op) == 4:
            mat = cache.get(str(op[3]))
            q0, q1 = int(op[1]), int(op[2])
            if mat is None:
                return None
            if q0 == q1:
                raise ValueError("unitary on a repeated qubit")
            return classify_unitary(_to_np(mat), [q0, q1], n)
    return None


def _to_np(x: Any) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=C128)
