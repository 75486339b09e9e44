"""Workload generators: the op sequences of the reference's circuit builders, as plain tuples.

Benchmarks must run the circuits the reference builds; these functions emit
the same op lists as (paths relative to the reference's src/tyxonq/):
  hea_ops       libs/circuits_library/blocks.py:14-56   (example_block)
  hwe_ry_ops    libs/circuits_library/blocks.py:60-85   (build_hwe_ry_ops, barriers dropped)
  qaoa_ring_ops libs/circuits_library/qaoa_ising.py:8-66 (ring of ZZ terms, mixer X)
  trotter_ops   libs/circuits_library/trotter_circuit.py:8-122
tests/golden/make_golden.py asserts op-for-op equality with the live builders.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


class Circuit:
    """Minimal stand-in for tyxonq.Circuit: what StatevectorEngine reads (core/ir/circuit.py:60-63)."""

    def __init__(self, num_qubits: int, ops: Sequence[tuple] | None = None, inputs=None) -> None:
        self.num_qubits = int(num_qubits)
        self.ops = list(ops or [])
        self._initial_state = inputs
        self._unitary_cache: dict = {}
        self._kraus_cache: dict = {}


def hea_ops(n: int, nlayers: int, params) -> List[tuple]:
    flat = np.asarray(params, dtype=np.float64).reshape(-1)
    ops: List[tuple] = [("h", q) for q in range(n)]
    for j in range(nlayers):
        base = j * 2 * n
        ops += [("cx", q, q + 1) for q in range(n - 1)]
        for q in range(n):
            ops.append(("rz", q, float(flat[base + q])))
            ops.append(("rx", q, float(flat[base + n + q])))
    return ops


def hwe_ry_ops(n: int, nlayers: int, params) -> List[tuple]:
    mat = np.asarray(params, dtype=np.float64).reshape(nlayers + 1, n)
    ops: List[tuple] = [("ry", i, float(mat[0, i])) for i in range(n)]
    for l in range(nlayers):
        ops += [("cx", i, i + 1) for i in range(n - 1)]
        ops += [("ry", i, float(mat[l + 1, i])) for i in range(n)]
    return ops


def qaoa_ring_ops(n: int, nlayers: int, params) -> List[tuple]:
    p = np.asarray(params, dtype=np.float64).reshape(-1)
    ops: List[tuple] = [("h", q) for q in range(n)]
    edges = [(i, (i + 1) % n) for i in range(n)] if n > 2 else [(0, 1)]
    for j in range(nlayers):
        for a, b in edges:
            ops.append(("rzz", min(a, b), max(a, b), float(p[2 * j])))
        for q in range(n):
            ops.append(("rx", q, float(p[2 * j + 1])))
    return ops


def _trotter_term(ps: Sequence[int], theta: float) -> List[tuple]:
    nz = [i for i, v in enumerate(ps) if v != 0]
    if not nz:
        return []
    if len(nz) == 1:
        q = nz[0]
        if ps[q] == 1:
            return [("h", q), ("rz", q, 2.0 * theta), ("h", q)]
        if ps[q] == 2:
            return [("sdg", q), ("h", q), ("rz", q, 2.0 * theta), ("h", q), ("s", q)]
        return [("rz", q, 2.0 * theta)]
    ops: List[tuple] = []
    for q in nz:
        if ps[q] == 1:
            ops.append(("h", q))
        elif ps[q] == 2:
            ops += [("sdg", q), ("h", q)]
    ops += [("cx", nz[i], nz[i + 1]) for i in range(len(nz) - 1)]
    ops.append(("rz", nz[-1], 2.0 * theta))
    ops += [("cx", nz[i], nz[i + 1]) for i in range(len(nz) - 2, -1, -1)]
    for q in reversed(nz):
        if ps[q] == 1:
            ops.append(("h", q))
        elif ps[q] == 2:
            ops += [("h", q), ("s", q)]
    return ops


def trotter_ops(terms: Sequence[Sequence[int]], weights: Sequence[float], time: float, steps: int) -> List[tuple]:
    n = len(terms[0])
    dt = float(time) / float(max(1, int(steps)))
    ops: List[tuple] = []
    for _ in range(max(1, int(steps))):
        for ps, c in zip(terms, weights):
            ops += _trotter_term(ps, float(c) * dt)
    ops += [("measure_z", q) for q in range(n)]
    return ops


def tfim_terms(n: int, J: float = 1.0, h: float = 1.0) -> Tuple[List[List[int]], List[float]]:
    terms: List[List[int]] = []
    w: List[float] = []
    for i in range(n - 1):
        ps = [0] * n
        ps[i] = ps[i + 1] = 3
        terms.append(ps)
        w.append(J)
    for i in range(n):
        ps = [0] * n
        ps[i] = 1
        terms.append(ps)
        w.append(h)
    return terms, w
