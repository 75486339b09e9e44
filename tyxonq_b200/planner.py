"""Host planner: lowered gates -> fused passes (tqb_pass / tqb_gate / matrix buffer).

The reference interprets one op at a time and allocates a new 2^n array per gate
(devices/simulators/statevector/engine.py:52-374).  Here a circuit is compiled into a short
list of *passes*; each pass streams the state through shared-memory tiles once and applies
every gate whose target bits are tile-local.  Diagonal gates never need locality (their
table is indexed by the global amplitude index), so they ride along with any pass.

Scheduling is a greedy list scheduler over the gate dependency order: gates that share no
index bit commute, so a gate may join the current pass when (a) none of its bits is blocked
by an earlier gate that could not be scheduled and (b) its target bits fit into the tile's
budget of high bits.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .gates import DENSE, DIAG, PAIR, SWAP, LGate

C128 = np.complex128


@dataclass
class TileConfig:
    m: int          # tile bits
    L: int          # low contiguous bits
    threads: int = 256
    ctas_per_sm: int = 0
    max_gates: int = 384

    @property
    def h(self) -> int:
        return self.m - self.L


def default_tile(n: int, itemsize: int, batch: int = 1) -> TileConfig:
    """itemsize = 8 (complex64) or 16 (complex128).  Streaming regime: 32 KiB tiles (several CTAs
    per SM overlap their load / compute / store phases), 512-byte contiguous runs.  Small states
    use one big tile so that whole layers stay in shared memory."""
    big = 13 if itemsize == 16 else 14           # 128 KiB tile
    stream_m = 11 if itemsize == 16 else 12      # 32 KiB tile
    stream_L = 5 if itemsize == 16 else 6        # 512 B runs
    m = int(os.environ.get("TQB_TILE_M", 0)) or (min(n, big) if (n <= big + 3 and batch <= 64) else stream_m)
    m = min(m, n)
    L = int(os.environ.get("TQB_TILE_L", 0)) or stream_L
    L = min(L, m)
    if m - L > 12:
        L = m - 12
    threads = int(os.environ.get("TQB_THREADS", 0)) or (256 if m <= 12 else 512)
    cps = int(os.environ.get("TQB_CTAS_PER_SM", 0))
    return TileConfig(m=m, L=L, threads=threads, ctas_per_sm=cps)


@dataclass
class Program:
    """A compiled circuit: host copies of the ABI arrays (uploaded by program.DeviceProgram)."""
    n: int
    passes: np.ndarray          # _lib.PASS_DTYPE
    gates: np.ndarray           # _lib.GATE_DTYPE
    mats: np.ndarray            # complex128, flat
    n_gates_in: int             # gates before planning (for gates/s accounting)
    tile: TileConfig
    order: List[int]            # order[i] = index (into the input list) of the i-th scheduled gate
    itemsize: int = 16          # 16: compiled for complex128 (R = 3), 8: complex64 (R = 4)

    @property
    def n_passes(self) -> int:
        return int(self.passes.shape[0])


def _low_mask(L: int) -> int:
    return (1 << L) - 1


def schedule(gates: Sequence[LGate], n: int, tile: TileConfig) -> List[Tuple[List[int], List[int]]]:
    """Return [(high_bits, [gate indices in execution order]), ...]."""
    N = len(gates)
    done = [False] * N
    remaining = N
    first = 0
    all_bits = (1 << n) - 1
    low = _low_mask(tile.L)
    h = tile.h
    out: List[Tuple[List[int], List[int]]] = []
    masks = [g.mask for g in gates]
    needs = [0 if g.kind == DIAG else (g.mask & ~low) for g in gates]
    is_diag = [g.kind == DIAG for g in gates]
    while remaining:
        while done[first]:
            first += 1
        H = 0
        nH = 0
        blocked = 0      # bits of skipped gates
        blocked_nd = 0   # ... of skipped non-diagonal gates only
        chosen: List[int] = []
        i = first
        seen = 0
        while i < N and seen < 4096:
            if not done[i]:
                seen += 1
                mk = masks[i]
                if is_diag[i]:
                    if mk & blocked_nd:
                        blocked |= mk
                    else:
                        chosen.append(i)
                elif mk & blocked:
                    blocked |= mk
                    blocked_nd |= mk
                else:
                    need = needs[i] & ~H
                    cnt = bin(need).count("1")
                    if nH + cnt <= h:
                        H |= need
                        nH += cnt
                        chosen.append(i)
                    else:
                        blocked |= mk
                        blocked_nd |= mk
                if blocked_nd == all_bits or len(chosen) >= tile.max_gates:
                    break
            i += 1
        for c in chosen:
            done[c] = True
        remaining -= len(chosen)
        # fill the tile with the lowest free bits (longer contiguous runs)
        b = tile.L
        while nH < h and b < n:
            if not (H >> b) & 1:
                H |= 1 << b
                nH += 1
            b += 1
        out.append(([p for p in range(n) if (H >> p) & 1], chosen))
    return out


_PERM_SWAP_BITS = [0, 2, 1, 3]  # 2-bit index with its bits exchanged


def _reg_capable(g: LGate, R: int) -> bool:
    if g.kind == DENSE:
        return g.k <= 2
    if g.kind == DIAG:
        return g.k <= 6
    if g.kind == SWAP:
        return g.k <= R
    if g.kind == PAIR:
        return g.k == 2 and g.zmask == 0
    return False


def micro_schedule(chosen: Sequence[int], gates: Sequence[LGate], local_of: dict, R: int, m: int):
    """Split one pass's gate list into units: ("micro", reg_bits(local, ascending), [gate idx]) runs whose
    non-diagonal gates act inside R register bits, and ("smem", gate idx) single shared-memory sweeps."""
    units = []
    remaining = list(chosen)
    if m < R:
        return [("smem", i) for i in remaining]
    while remaining:
        first = gates[remaining[0]]
        if not _reg_capable(first, R):
            units.append(("smem", remaining.pop(0)))
            continue
        rset: List[int] = []
        picked: List[int] = []
        rest: List[int] = []
        blocked = 0
        blocked_nd = 0
        for i in remaining:
            g = gates[i]
            mk = g.mask
            if g.kind == DIAG and _reg_capable(g, R):
                if mk & blocked_nd:
                    blocked |= mk
                    rest.append(i)
                else:
                    picked.append(i)
                continue
            if (mk & blocked) or not _reg_capable(g, R):
                blocked |= mk
                blocked_nd |= mk
                rest.append(i)
                continue
            need = [local_of[b] for b in g.bits if local_of[b] not in rset]
            if len(rset) + len(need) <= R:
                rset += need
                picked.append(i)
            else:
                blocked |= mk
                blocked_nd |= mk
                rest.append(i)
        p = 0
        while len(rset) < R:  # pad with unused tile-local bits
            if p not in rset:
                rset.append(p)
            p += 1
        units.append(("micro", sorted(rset), picked))
        remaining = rest
    return units


def compile_program(gates: Sequence[LGate], n: int, tile: TileConfig, *, batch_mats: int = 1, micro: bool = True,
                    itemsize: int = 16) -> Program:
    """Pack scheduled gates into the ABI arrays.  ``batch_mats`` > 1: every gate's ``data`` has a
    leading batch axis (one matrix per batch member).  ``micro``: group gates into register
    micro-passes (R = 3 register bits for complex128, 4 for complex64 -- ``itemsize`` 16 / 8)."""
    m_eff = min(tile.m, n)
    tile = TileConfig(m=m_eff, L=min(tile.L, m_eff), threads=tile.threads, ctas_per_sm=tile.ctas_per_sm,
                      max_gates=tile.max_gates)
    R = 3 if itemsize == 16 else 4
    sched = schedule(gates, n, tile)
    passes = np.zeros(len(sched), dtype=_lib.PASS_DTYPE)
    descs: List[np.void] = []
    mats: List[np.ndarray] = []
    mat_off = 0
    order: List[int] = []

    def new_desc() -> np.ndarray:
        return np.zeros(1, dtype=_lib.GATE_DTYPE)

    def add_matrix(e, data: np.ndarray) -> None:
        nonlocal mat_off
        d = np.asarray(data, dtype=C128)
        if batch_mats > 1:
            d = d.reshape(batch_mats, -1)
            e["mat_off"] = mat_off
            e["mat_bstride"] = d.shape[1]
            mats.append(d.reshape(-1))
            mat_off += d.size
        else:
            d = d.reshape(-1)
            e["mat_off"] = mat_off
            e["mat_bstride"] = 0
            mats.append(d)
            mat_off += d.size

    for pi, (hb, chosen) in enumerate(sched):
        L = tile.L
        assert len(hb) == m_eff - L, (hb, m_eff, L)
        local_of = {p: p for p in range(L)}
        for j, p in enumerate(hb):
            local_of[p] = L + j
        ps = passes[pi]
        ps["m"] = m_eff
        ps["L"] = L
        ps["gate_begin"] = len(descs)
        ps["mat_begin"] = mat_off
        for j, p in enumerate(hb):
            ps["hb"][j] = p
        maxk = 0
        units = micro_schedule(chosen, gates, local_of, R, m_eff) if micro else [("smem", i) for i in chosen]
        for unit in units:
            if unit[0] == "smem":
                g = gates[unit[1]]
                e = new_desc()
                e["kind"] = g.kind
                e["k"] = g.k
                if g.kind == DIAG:
                    assert g.k <= 6, "diagonal tables are limited to 6 bits"
                    for j, b in enumerate(g.bits):  # tile-local position, or 64 + index bit when outside the tile
                        e["bits"][0, j] = local_of[b] if b in local_of else 64 + b
                else:
                    assert g.k <= 4, "dense / pair gates are limited to 4 bits"
                    loc = [local_of[b] for b in g.bits]
                    for j, b in enumerate(loc):
                        e["bits"][0, j] = b
                    for j, b in enumerate(sorted(loc)):
                        e["sbits"][0, j] = b
                    if g.kind in (PAIR, SWAP):
                        e["off_a"] = sum(((g.pat_a >> j) & 1) << loc[j] for j in range(g.k))
                        e["off_b"] = sum(((g.pat_b >> j) & 1) << loc[j] for j in range(g.k))
                        e["zmask"] = g.zmask
                    else:
                        maxk = max(maxk, g.k)
                add_matrix(e, g.data)
                descs.append(e)
                order.append(unit[1])
                continue
            _, rbits, picked = unit
            hdr = new_desc()
            hdr["kind"] = _lib.GATE_MICRO
            hdr["k"] = R
            for j, b in enumerate(rbits):
                hdr["bits"][0, j] = b
            hdr["off_a"] = len(picked)
            descs.append(hdr)
            rho_of = {b: j for j, b in enumerate(rbits)}
            for idx in picked:
                g = gates[idx]
                e = new_desc()
                e["k"] = g.k
                if g.kind == DIAG:
                    e["kind"] = _lib.GATE_RDIAG
                    for j, b in enumerate(g.bits):
                        if b in local_of and local_of[b] in rho_of:
                            rho = rho_of[local_of[b]]
                            e["bits"][0, j] = 32 + rho
                            e["sbits"][0, rho] = 1 << j
                        elif b in local_of:
                            e["bits"][0, j] = local_of[b]
                        else:
                            e["bits"][0, j] = 64 + b
                    add_matrix(e, g.data)
                elif g.kind == SWAP:
                    e["kind"] = _lib.GATE_RSWAP
                    rho = [rho_of[local_of[b]] for b in g.bits]
                    e["off_a"] = sum(1 << r for r in rho)
                    aval = sum(((g.pat_a >> j) & 1) << rho[j] for j in range(g.k))
                    bval = sum(((g.pat_b >> j) & 1) << rho[j] for j in range(g.k))
                    e["off_b"] = aval
                    e["zmask"] = aval ^ bval
                    add_matrix(e, g.data)
                else:
                    e["kind"] = _lib.GATE_RDENSE
                    data = g.data
                    if g.kind == PAIR:  # 2x2 block on (pat_a, pat_b) of a 2-bit index, identity elsewhere
                        assert batch_mats == 1
                        M4 = np.eye(4, dtype=C128)
                        blk = np.asarray(g.data[:4]).reshape(2, 2)
                        ab = [g.pat_a, g.pat_b]
                        for r_ in range(2):
                            for c_ in range(2):
                                M4[ab[r_], ab[c_]] = blk[r_, c_]
                        data = M4
                    rho = [rho_of[local_of[b]] for b in g.bits]
                    if g.k == 2 and rho[0] > rho[1]:  # canonical order: matrix-index bit 0 on the lower register bit
                        rho = [rho[1], rho[0]]
                        d4 = np.asarray(data, dtype=C128).reshape(-1, 4, 4)
                        data = d4[:, _PERM_SWAP_BITS][:, :, _PERM_SWAP_BITS]
                    for j, r in enumerate(rho):
                        e["bits"][0, j] = r
                    e["k"] = len(rho)
                    add_matrix(e, data)
                descs.append(e)
                order.append(idx)
        ps["n_gates"] = len(descs) - int(ps["gate_begin"])
        ps["max_dense_k"] = maxk
        cnt = mat_off - int(ps["mat_begin"])
        ps["mat_count"] = cnt if (batch_mats == 1 and cnt <= 4096) else 0
    garr = np.concatenate(descs) if descs else np.zeros(0, dtype=_lib.GATE_DTYPE)
    flat = np.concatenate(mats) if mats else np.zeros(0, dtype=C128)
    return Program(n=n, passes=passes, gates=garr, mats=flat, n_gates_in=len(gates), tile=tile, order=order, itemsize=itemsize)
