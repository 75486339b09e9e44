"""Host-side gate fusion before pass planning.

The reference applies one einsum per op (devices/simulators/statevector/engine.py:52-374).
Inside a fused pass every gate still costs one sweep over the shared-memory tile, so gates
are first merged where that is free:

  * consecutive 1-qubit gates on the same qubit        -> one 2x2 (rz then rx of the HEA layer)
  * consecutive diagonal gates (rz, s, cz, rzz, ...)   -> one table over the union of their bits (<= 6)
  * a 1-qubit gate next to a dense 2-qubit gate        -> folded into the 4x4 (same arithmetic cost)
  * cx next to a 1-qubit gate on its target            -> one MUX gate: the 2x2 is U (control 0) or
    X.U / U.X (control 1); costs one 1-qubit sweep and only the target has to be tile-local

Gates only move past gates they share no index bit with, so the circuit's unitary is unchanged
(products are formed in complex128 on the host).  Fused gates lose their gradient bookkeeping:
the adjoint paths plan the unfused list.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .gates import DENSE, DIAG, LAYER_PLAIN, MUX, MUX_XU, SWAP, X_MAT, LGate, chain_gate, mux_gate, rot_decompose, rot_plan

C128 = np.complex128
_I2 = np.eye(2, dtype=C128)


def _batched(g: LGate) -> bool:
    return bool(g.batched)


def _is_1q(g: Optional[LGate]) -> bool:
    """1-qubit DENSE/DIAG gate (DENSE may carry one matrix per batch member)."""
    return g is not None and g.kind in (DENSE, DIAG) and len(g.bits) == 1 and not (g.kind == DIAG and g.batched)


def _is_cx(g: LGate) -> bool:
    """SWAP gate that flips matrix bit 0 (target = bits[0]) when matrix bit 1 (control = bits[1]) is set."""
    return g.kind == SWAP and len(g.bits) == 2 and g.pat_a == 0b10 and g.pat_b == 0b11


def _m1(g: LGate) -> np.ndarray:
    """2x2 matrix, or [B, 2, 2] for a batched gate (numpy matmul broadcasts over the batch axis)."""
    if g.kind == DIAG:
        return np.diag(g.data)
    return g.data.reshape(-1, 2, 2) if g.batched else g.data


def _mul_1q(later: LGate, earlier: LGate) -> LGate:
    if later.kind == DIAG and earlier.kind == DIAG:
        return LGate(DIAG, later.bits, later.data * earlier.data, name="fused")
    m = _m1(later) @ _m1(earlier)
    if m.ndim == 3:
        return LGate(DENSE, later.bits, np.ascontiguousarray(m.reshape(-1, 4)), batched=True, name="fused")
    return LGate(DENSE, later.bits, m, name="fused")


def _embed_1q(u: np.ndarray, j: int) -> np.ndarray:
    """2x2 on matrix-index bit j of a 2-bit index (bit 0 = least significant = right kron factor)."""
    return np.kron(_I2, u) if j == 0 else np.kron(u, _I2)


def _merge_diag(a: LGate, b: LGate) -> LGate:
    """Table of a*b over the union of the bits (a's bits first)."""
    bits = list(a.bits) + [x for x in b.bits if x not in a.bits]
    k = len(bits)
    idx = np.arange(1 << k)

    def sub(g: LGate) -> np.ndarray:
        t = np.zeros(1 << k, dtype=np.int64)
        for j, bit in enumerate(g.bits):
            t |= ((idx >> bits.index(bit)) & 1) << j
        return g.data[t]

    return LGate(DIAG, tuple(bits), sub(a) * sub(b), name="fused")


def _simplify_mux(g: LGate) -> LGate:
    """A MUX whose two matrices are both diagonal is a 2-bit diagonal gate (cx.rz.cx = rzz):
    table index bit 0 = target, bit 1 = control."""
    if g.kind != MUX or g.batched:
        return g
    u0, u1 = g.data[:4].reshape(2, 2), g.data[4:].reshape(2, 2)
    if u0[0, 1] == 0 and u0[1, 0] == 0 and u1[0, 1] == 0 and u1[1, 0] == 0:
        return LGate(DIAG, (g.bits[0], g.bits[1]), np.array([u0[0, 0], u0[1, 1], u1[0, 0], u1[1, 1]], dtype=C128), name="fused")
    return g


def _as_layer(g: LGate):
    """(target_bit, control_bit | None, M_sel0, M_sel1, structure tag) for gates a CHAIN layer can carry."""
    if _is_1q(g):
        m = _m1(g)
        return g.bits[0], None, m, m, LAYER_PLAIN, _rot_of(g, m)
    if g.kind == MUX:
        u0, u1 = _mux_blocks(g)
        return g.bits[0], g.bits[1], u0, u1, g.pat_b, _rot_of(g, u0)
    return None


def _rot_of(g: LGate, u0: np.ndarray):
    d = g.__dict__.get("_rot")
    if d is None:
        d = rot_decompose(u0)
        g.__dict__["_rot"] = d
    return d


def _mux_blocks(g: LGate):
    """(U_control0, U_control1) of a MUX gate: 2x2 each, or [B, 2, 2] when batched."""
    if g.batched:
        return g.data[:, :4].reshape(-1, 2, 2), g.data[:, 4:].reshape(-1, 2, 2)
    return g.data[:4].reshape(2, 2), g.data[4:].reshape(2, 2)


def chain_fuse(gates: Sequence[LGate], R: int = 3, R_rot: int = 4, local_bits: Optional[set] = None) -> List[LGate]:
    """Group ADJACENT 1-qubit / MUX gates into CHAIN gates of up to R layers: one shared-memory round
    trip for R gates.  Layer i > 0 must be uncontrolled or controlled by layer i-1's target (the shape
    of a cx ladder); layer 0 may carry any control.  Chains whose layers all factor as rotation x diagonal
    (gates.rot_plan) may grow to R_rot layers: the rotation form needs few registers per layer.  ``local_bits``: the
    index bits that are tile-local (grouping after scheduling); layers must have their target among them."""
    out: List[LGate] = []
    cur: List[tuple] = []   # (layer, original gate)

    def flush() -> None:
        if len(cur) == 1:
            out.append(cur[0][1])
        elif cur:
            c0 = cur[0][0][1]
            out.append(chain_gate([(l[0], l[2], l[3]) for l, _ in cur], control_bit=c0,
                                  structure=[l[4] for l, _ in cur], decs=[l[5] for l, _ in cur], name="chain"))
        cur.clear()

    for g in gates:
        lay = _as_layer(g)
        if lay is not None and local_bits is not None and lay[0] not in local_bits:
            lay = None   # a diagonal 1-qubit gate riding along on a bit outside the tile: it cannot be a chain layer
        if lay is None:
            flush()
            out.append(g)
            continue
        t, c = lay[0], lay[1]
        if cur:
            tgts = [l[0] for l, _ in cur]
            c0 = cur[0][0][1]
            if t not in tgts and t != c0 and (c is None or c == tgts[-1]):
                if len(cur) < R:
                    cur.append((lay, g))
                    continue
                if len(cur) < R_rot and rot_plan([l[5] for l, _ in cur] + [lay[5]], [l[4] for l, _ in cur] + [lay[4]],
                                                 c0 is not None) is not None:
                    cur.append((lay, g))
                    continue
            flush()
        cur.append((lay, g))
    flush()
    return out


def _conjugated_diagonals(gates: Sequence[LGate]) -> List[LGate]:
    """Peephole: cx(c,t) . D(t) . cx(c,t) with D diagonal on t and nothing else on c or t in between is the
    2-bit diagonal table[c*2+t] = D[t xor c] (the ZZ rotation of a Trotter step, trotter_circuit.py:52-60)."""
    gates = list(gates)
    n = len(gates)
    nxt: List[Dict[int, int]] = [dict() for _ in range(n)]   # nxt[i][bit] = index of the next gate touching bit
    last: Dict[int, int] = {}
    for i in range(n - 1, -1, -1):
        for b in gates[i].bits:
            if b in last:
                nxt[i][b] = last[b]
            last[b] = i
    dead = [False] * n
    for i, g in enumerate(gates):
        if dead[i] or not _is_cx(g):
            continue
        t, c = g.bits[0], g.bits[1]
        j = nxt[i].get(t)
        if j is None or dead[j]:
            continue
        d = gates[j]
        if not (d.kind == DIAG and d.bits == (t,)):
            continue
        k = nxt[j].get(t)
        if k is None or dead[k] or nxt[i].get(c) != k:
            continue
        g2 = gates[k]
        if not (_is_cx(g2) and g2.bits == (t, c)):
            continue
        tab = np.array([d.data[0], d.data[1], d.data[1], d.data[0]], dtype=C128)
        gates[i] = LGate(DIAG, (t, c), tab, name="fused")
        dead[j] = dead[k] = True
    return [g for i, g in enumerate(gates) if not dead[i]]


class FusedGates(list):
    """Output of ``fuse`` when CHAIN grouping is deferred: ``compile_program`` groups the gates of every pass into
    chains AFTER scheduling (a chain needs all its targets tile-local at once; grouping first would tie the
    scheduler's hands: 605 instead of 400 passes for a 30-qubit depth-100 hardware-efficient ansatz)."""
    chain_after_schedule = True


def fuse(gates: Sequence[LGate], max_diag_k: int = 6, chain: int = 0) -> List[LGate]:
    """Algebraic merges (see module docstring).  ``chain`` >= 2 groups CHAINs of up to that many layers right away
    (the round-1 first version); the default 0 leaves the grouping to ``compile_program`` (per pass, after scheduling)."""
    if chain >= 2:
        return chain_fuse(_merge(_conjugated_diagonals(list(gates)), max_diag_k), chain, chain)
    # deferred mode: multi-bit diagonal gates stay small as well -- group_pass absorbs them into the chains of their
    # pass or merges what is left into tables there
    return FusedGates(_merge(_conjugated_diagonals(list(gates)), max_diag_k, merge_diag=False))


def _merge(gates: Sequence[LGate], max_diag_k: int = 6, merge_diag: bool = True) -> List[LGate]:
    out: List[Optional[LGate]] = []
    last: Dict[int, int] = {}

    def touch(g: LGate, j: int) -> None:
        for b in g.bits:
            last[b] = j

    for g in gates:
        if g.batched and not _is_1q(g):   # only batched 1-qubit gates take part in fusion
            out.append(g)
            touch(g, len(out) - 1)
            continue
        if _is_1q(g):
            t = g.bits[0]
            j = last.get(t)
            p = out[j] if j is not None else None
            if p is not None:
                if _is_1q(p):
                    out[j] = _mul_1q(g, p)
                    continue
                if p.kind == DENSE and len(p.bits) == 2 and not p.batched and not g.batched:
                    out[j] = LGate(DENSE, p.bits, _embed_1q(_m1(g), p.bits.index(t)) @ p.data, name="fused")
                    continue
                # 1q gate after a MUX on the same target -- unless the MUX is "gate, then cx" (MUX_XU): that structure is
                # what lets it ride in a rotation-form chain, and the 1q gate will pair up with the NEXT cx the same way
                # (last qubit of a cx ladder: otherwise one lone general MUX sweep per layer)
                if p.kind == MUX and p.bits[0] == t and p.pat_b != MUX_XU:
                    v = _m1(g)
                    u0, u1 = _mux_blocks(p)
                    out[j] = mux_gate(v @ u0, v @ u1, t, p.bits[1], name="fused")
                    continue
                if _is_cx(p) and p.bits[0] == t:          # 1q gate after cx on its target: V (c=0), V.X (c=1)
                    v = _m1(g)
                    out[j] = mux_gate(v, v @ X_MAT, t, p.bits[1], name="fused")
                    continue
        if _is_cx(g):
            t, c = g.bits[0], g.bits[1]
            j = last.get(t)
            p = out[j] if j is not None else None
            if _is_1q(p):                                  # cx after a 1q gate on its target: U (c=0), X.U (c=1)
                u = _m1(p)
                out[j] = None
                g = mux_gate(u, X_MAT @ u, t, c, structure=MUX_XU, name="fused")
            elif p is not None and p.kind == MUX and p.bits == (t, c) and last.get(c) == j:
                u0, u1 = _mux_blocks(p)
                out[j] = _simplify_mux(mux_gate(u0, X_MAT @ u1, t, c, name="fused"))
                continue
        if g.kind == DIAG and merge_diag:
            js = [last[b] for b in g.bits if b in last]
            if js:
                j = max(js)
                p = out[j]
                if p is not None and p.kind == DIAG and len(set(p.bits) | set(g.bits)) <= max_diag_k:
                    out[j] = _merge_diag(p, g)
                    touch(g, j)
                    continue
        if g.kind == DENSE and len(g.bits) == 2:
            M = g.data
            for b in g.bits:
                j = last.get(b)
                p = out[j] if j is not None else None
                if _is_1q(p) and not p.batched:
                    M = M @ _embed_1q(_m1(p), g.bits.index(b))
                    out[j] = None
            if M is not g.data:
                g = LGate(DENSE, g.bits, M, name="fused")
        out.append(g)
        touch(g, len(out) - 1)
    return [x for x in out if x is not None]


# ------------------------------------------------------------------------------------------
# per-pass grouping (after scheduling)
# ------------------------------------------------------------------------------------------
def sink_diagonals(gates: Sequence[LGate], sched: List[tuple]) -> List[tuple]:
    """Move every diagonal gate forward to just before the next non-diagonal gate that touches one of its bits (it
    commutes with everything in between), so that it lands in the pass -- and next to the chain -- that can absorb
    it.  The scheduler hands diagonal gates out as early as possible because they need no tile bit."""
    flat = [(pi, gi) for pi, (_, chosen) in enumerate(sched) for gi in chosen]
    new_lists: List[List[int]] = [[] for _ in sched]
    waiting: List[int] = []          # diagonal gates (indices) not yet placed, in order
    for pi, gi in flat:
        g = gates[gi]
        if g.kind == DIAG and not g.batched:
            waiting.append(gi)
            continue
        if waiting:
            keep = []
            for d in waiting:
                if gates[d].mask & g.mask:
                    new_lists[pi].append(d)
                else:
                    keep.append(d)
            waiting = keep
        new_lists[pi].append(gi)
    if waiting:                      # nothing touches them any more: the last pass takes them
        new_lists[-1].extend(waiting)
    return [(hb, lst) for (hb, _), lst in zip(sched, new_lists)]


def _merge_diag_run(diags: List[LGate], max_k: int = 6) -> List[LGate]:
    out: List[LGate] = []
    for d in diags:
        if out and not out[-1].batched and not d.batched and len(set(out[-1].bits) | set(d.bits)) <= max_k:
            out[-1] = _merge_diag(out[-1], d)
        else:
            out.append(d)
    return out


def group_pass(gates: Sequence[LGate], local_bits: set, R: int = 3, R_rot: int = 4, E_max: int = 2) -> List[LGate]:
    """Gates of ONE pass (execution order) -> chains + leftover gates.  1-qubit / MUX gates on tile-local bits become
    CHAIN layers like in ``chain_fuse``; diagonal gates that sit right before a rotation-form chain (or between its
    layers without touching an earlier layer's target) are multiplied into the chain's pre-diagonal table when their
    bits are chain targets plus at most ``E_max`` other bits -- the ZZ terms of a Trotter / QAOA layer then cost no
    sweep of their own.  Leftover diagonal gates are merged into tables of <= 6 bits."""
    out: List[LGate] = []
    pend: List[LGate] = []           # diagonal gates waiting right before the next non-diagonal gate
    N = len(gates)

    def layer_of(g: LGate):
        lay = _as_layer(g)
        if lay is not None and lay[0] not in local_bits:
            return None
        return lay

    i = 0
    while i < N:
        g = gates[i]
        if g.kind == DIAG and not g.batched and len(g.bits) > 1:
            pend.append(g)
            i += 1
            continue
        lay = layer_of(g)
        if lay is None:
            out.extend(_merge_diag_run(pend))
            pend = []
            out.append(g)
            i += 1
            continue
        cur = [(lay, g)]
        tmask = 1 << lay[0]
        hoist: List[LGate] = []      # diagonal gates between layers that commute to the front of the chain
        n_hoist_ok = 0
        j = i + 1
        end = j
        while j < N:
            h = gates[j]
            if h.kind == DIAG and not h.batched and len(h.bits) > 1:
                if h.mask & tmask:
                    break
                hoist.append(h)
                j += 1
                continue
            l2 = layer_of(h)
            if l2 is None:
                break
            t, c = l2[0], l2[1]
            tgts = [l[0] for l, _ in cur]
            c0 = cur[0][0][1]
            if not (t not in tgts and t != c0 and (c is None or c == tgts[-1])):
                break
            if len(cur) >= R and not (len(cur) < R_rot and rot_plan([l[5] for l, _ in cur] + [l2[5]],
                                                                     [l[4] for l, _ in cur] + [l2[4]], c0 is not None) is not None):
                break
            cur.append((l2, h))
            tmask |= 1 << t
            j += 1
            end = j
            n_hoist_ok = len(hoist)
        hoist = hoist[:n_hoist_ok]   # diagonal gates after the last accepted layer stay where they are (i resumes at ``end``)
        if len(cur) == 1:
            # no chain: emit the pending diagonals, the hoisted ones never moved (n_hoist_ok == 0 here)
            out.extend(_merge_diag_run(pend))
            pend = []
            out.append(g)
            i += 1
            continue
        c0 = cur[0][0][1]
        tags = [l[4] for l, _ in cur]
        decs = [l[5] for l, _ in cur]
        cands = pend + hoist
        absorbed: List[LGate] = []
        extras: List[int] = []
        if cands and not any(gt.batched for _, gt in cur) and rot_plan(decs, tags, c0 is not None) is not None:
            targets = [l[0] for l, _ in cur]
            tset = set(targets)
            need = [tuple(b for b in d.bits if b not in tset) for d in cands]
            pool = sorted({b for nb in need for b in nb})
            best = (-1, ())
            combos = [()] + [(x,) for x in pool] + [(x, y) for ix, x in enumerate(pool) for y in pool[ix + 1:]]
            for cb in combos:
                if len(cb) > E_max:
                    continue
                cnt = sum(1 for nb in need if set(nb) <= set(cb))
                if cnt > best[0] or (cnt == best[0] and len(cb) < len(best[1])):
                    best = (cnt, cb)
            extras = list(best[1])
            absorbed = [d for d, nb in zip(cands, need) if set(nb) <= set(extras)]
        ab_ids = {id(d) for d in absorbed}
        out.extend(_merge_diag_run([d for d in cands if id(d) not in ab_ids]))
        pend = []
        out.append(chain_gate([(l[0], l[2], l[3]) for l, _ in cur], control_bit=c0, structure=tags, decs=decs,
                              pre_diags=absorbed, extra_bits=extras, name="chain"))
        i = end
    out.extend(_merge_diag_run(pend))
    return out
