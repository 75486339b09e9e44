"""Host-side gate fusion before pass planning.

The reference applies one einsum per op (devices/simulators/statevector/engine.py:52-374).
Inside a fused pass every gate still costs one sweep over the shared-memory tile, so gates
are first merged where that is free:

  * consecutive 1-qubit gates on the same qubit        -> one 2x2 (rz then rx of the HEA layer)
  * consecutive diagonal gates (rz, s, cz, rzz, ...)   -> one table over the union of their bits (<= 6)
  * a 1-qubit gate next to a dense 2-qubit gate        -> folded into the 4x4 (same arithmetic cost)

Gates only move past gates they share no index bit with, so the circuit's unitary is unchanged
(products are formed in complex128 on the host).  Fused gates lose their gradient bookkeeping:
the adjoint paths plan the unfused list.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .gates import DENSE, DIAG, LGate

C128 = np.complex128
_I2 = np.eye(2, dtype=C128)


def _batched(g: LGate) -> bool:
    return g.kind == DENSE and g.data.ndim == 2 and g.data.shape[0] != g.data.shape[1]


def _m1(g: LGate) -> np.ndarray:
    return np.diag(g.data) if g.kind == DIAG else g.data


def _mul_1q(later: LGate, earlier: LGate) -> LGate:
    if later.kind == DIAG and earlier.kind == DIAG:
        return LGate(DIAG, later.bits, later.data * earlier.data, name="fused")
    return LGate(DENSE, later.bits, _m1(later) @ _m1(earlier), name="fused")


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


def fuse(gates: Sequence[LGate], max_diag_k: int = 6) -> List[LGate]:
    out: List[Optional[LGate]] = []
    last: Dict[int, int] = {}

    def touch(g: LGate, j: int) -> None:
        for b in g.bits:
            last[b] = j

    for g in gates:
        if _batched(g) or g.data.ndim > 2:
            out.append(g)
            touch(g, len(out) - 1)
            continue
        one_q = g.k == 1 and g.kind in (DENSE, DIAG)
        if one_q:
            j = last.get(g.bits[0])
            p = out[j] if j is not None else None
            if p is not None and not _batched(p):
                if p.k == 1 and p.kind in (DENSE, DIAG):
                    out[j] = _mul_1q(g, p)
                    continue
                if p.kind == DENSE and p.k == 2:
                    out[j] = LGate(DENSE, p.bits, _embed_1q(_m1(g), p.bits.index(g.bits[0])) @ p.data, name="fused")
                    continue
        if g.kind == DIAG:
            js = [last[b] for b in g.bits if b in last]
            if js:
                j = max(js)
                p = out[j]
                if p is not None and p.kind == DIAG and len(set(p.bits) | set(g.bits)) <= max_diag_k:
                    out[j] = _merge_diag(p, g)
                    touch(g, j)
                    continue
        if g.kind == DENSE and g.k == 2:
            M = g.data
            for b in g.bits:
                j = last.get(b)
                p = out[j] if j is not None else None
                if p is not None and p.k == 1 and p.kind in (DENSE, DIAG) and not _batched(p):
                    M = M @ _embed_1q(_m1(p), g.bits.index(b))
                    out[j] = None
            if M is not g.data:
                g = LGate(DENSE, g.bits, M, name="fused")
        out.append(g)
        touch(g, len(out) - 1)
    return [x for x in out if x is not None]
