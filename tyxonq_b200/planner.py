"""Host planner: lowered gates -> fused passes (tqb_pass / tqb_gate / matrix buffer).

The reference interprets one op at a time and allocates a new 2^n array per gate
(devices/simulators/statevector/engine.py:52-374).  Here a circuit is compiled into a short
list of *passes*; each pass streams the state through shared-memory tiles once and applies
every gate whose target bits are tile-local.  Diagonal gates never need locality (their
table is indexed by the global amplitude index) and the control of a MUX gate does not either,
so they ride along with any pass.

Scheduling is a greedy list scheduler over the gate dependency order: gates that share no
index bit commute, so a gate may join the current pass when (a) none of its bits is blocked
by an earlier gate that could not be scheduled and (b) the bits it needs tile-local fit into
the tile's budget of high bits.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .gates import CHAIN, DENSE, DIAG, MUX, PAIR, SWAP, LGate

C128 = np.complex128


@dataclass
class TileConfig:
    m: int          # tile bits
    L: int          # low contiguous bits
    threads: int = 256
    ctas_per_sm: int = 0
    max_gates: int = 384
    rot_layers: int = 4   # longest rotation-form chain (4 needs the 128-thread kernel variant for complex128)
    max_work: int = 0     # cap on the non-diagonal gates (1-qubit layers after fusion) of one pass; 0 = no cap
    balance: int = 0      # target number of non-diagonal gates per pass (0 = off): low-bits-only gates beyond it wait

    @property
    def h(self) -> int:
        return self.m - self.L


def default_tile(n: int, itemsize: int, batch: int = 1) -> TileConfig:
    """itemsize = 8 (complex64) or 16 (complex128).  Streaming regime: 32 KiB tiles, double
    buffered by TMA (3 CTAs per SM: 128 consumer threads + a producer warp each), 512-byte
    contiguous runs, 6 high tile bits (profiles/r01_tile_sweep.md: one-gate pass 6.18 TB/s).
    Small states use one big tile so that whole layers stay in shared memory."""
    big = 13 if itemsize == 16 else 14           # 128 KiB tile
    stream_m = 11 if itemsize == 16 else 12      # 32 KiB tile
    stream_L = 5 if itemsize == 16 else 6        # 512 B runs
    m = int(os.environ.get("TQB_TILE_M", 0)) or (min(n, big) if (n <= big + 3 and batch <= 64) else stream_m)
    m = min(m, n)
    L = int(os.environ.get("TQB_TILE_L", 0)) or stream_L
    L = min(L, m)
    if m - L > 12:
        L = m - 12
    threads = int(os.environ.get("TQB_THREADS", 0)) or (128 if m <= (11 if itemsize == 16 else 12) else 256)
    cps = int(os.environ.get("TQB_CTAS_PER_SM", 0))
    return TileConfig(m=m, L=L, threads=threads, ctas_per_sm=cps)


@dataclass
class Program:
    """A compiled circuit: host copies of the ABI arrays (uploaded by program.DeviceProgram)."""
    n: int
    passes: np.ndarray          # _lib.PASS_DTYPE
    gates: np.ndarray           # _lib.GATE_DTYPE, in execution order
    mats: np.ndarray            # complex128, flat
    n_gates_in: int             # gates before planning (for gates/s accounting)
    tile: TileConfig
    order: List[int]            # order[i] = index (into the input list) of the i-th scheduled gate

    @property
    def n_passes(self) -> int:
        return int(self.passes.shape[0])


def schedule(gates: Sequence[LGate], n: int, tile: TileConfig) -> List[Tuple[List[int], List[int]]]:
    """Return [(high_bits, [gate indices in execution order]), ...].

    Greedy list scheduling over the dependency order (gates on disjoint bits commute; diagonal gates commute among
    themselves and need no tile bit).  For every pass two kinds of tile are tried and the one that runs the most
    non-diagonal gates wins: (i) first come, first served -- gates take high tile bits in program order until the
    budget h is spent; (ii) every window of h contiguous high bits that contains the bits of the oldest pending
    gate.  (ii) is what keeps a diagonal wavefront alive on ladder circuits (cx chains): with (i) alone the oldest
    layer takes the whole budget and every pass advances one layer by h qubits; a window lets 3-4 layers advance
    together (12 instead of 6 gates per pass on a 30-qubit hardware-efficient ansatz)."""
    N = len(gates)
    done = [False] * N
    remaining = N
    first = 0
    all_bits = (1 << n) - 1
    low = (1 << tile.L) - 1
    h = tile.h
    out: List[Tuple[List[int], List[int]]] = []
    masks = [g.mask for g in gates]
    needs = [g.local_mask & ~low for g in gates]   # bits that must become high tile bits
    is_diag = [g.kind == DIAG for g in gates]

    max_gates = tile.max_gates
    max_work = tile.max_work or (1 << 30)

    def select(start: int, window: int, low_quota: int = 1 << 30) -> Tuple[List[int], int, int, int]:
        """window < 0: budget mode (i); else only gates whose needs lie inside ``window`` run.  ``low_quota``: how many
        gates that need NO high tile bit (they can run in any pass) may join; the others wait.  Returns
        (chosen, H, number of high bits used, number of non-diagonal gates chosen)."""
        H = 0
        nH = 0
        blocked = 0      # bits of skipped gates
        blocked_nd = 0   # ... of skipped non-diagonal gates only
        chosen: List[int] = []
        push = chosen.append
        nd = 0
        seen = 0
        budget = window < 0
        outside = ~window
        for i in range(start, N):
            if done[i]:
                continue
            seen += 1
            mk = masks[i]
            if is_diag[i]:
                if mk & blocked_nd:
                    blocked |= mk
                else:
                    push(i)
            elif mk & blocked:
                blocked |= mk
                blocked_nd |= mk
            elif not needs[i] and low_quota <= 0:
                blocked |= mk          # a low-bits-only gate over the quota: it waits for a pass with room
                blocked_nd |= mk
            else:
                if not needs[i]:
                    low_quota -= 1
                need = needs[i] & ~H
                if budget:
                    cnt = need.bit_count()
                    ok = nH + cnt <= h
                else:
                    ok = not (need & outside)
                    cnt = need.bit_count() if ok else 0
                if ok:
                    H |= need
                    nH += cnt
                    push(i)
                    nd += 1
                else:
                    blocked |= mk
                    blocked_nd |= mk
            if blocked_nd == all_bits or len(chosen) >= max_gates or seen >= 4096:
                break
            if nd >= max_work:   # the pass is full: everything else waits (diagonal gates ride with their consumers)
                break
        return chosen, H, nH, nd

    while remaining:
        while done[first]:
            first += 1
        chosen, H, nH, nd = select(first, -1)
        # windows of h contiguous high bits around the oldest pending non-diagonal gate
        j = first
        while j < N and (done[j] or is_diag[j] or (tile.balance > 0 and not needs[j])):
            j += 1   # (balanced mode: the oldest pending gate that NEEDS a window; low-bits-only gates fit any pass)
        if j < N and needs[j] and h >= 2 and n - tile.L > h:
            nb = needs[j]
            lo_need = (nb & -nb).bit_length() - 1
            hi_need = nb.bit_length() - 1
            for start in range(max(tile.L, hi_need - h + 1), min(lo_need, n - h) + 1):
                window = ((1 << h) - 1) << start
                c2, H2, nH2, nd2 = select(first, window)
                if nd2 > nd:
                    chosen, H, nH, nd = c2, H2, nH2, nd2
        if tile.balance > 0 and nd > tile.balance:
            # Balance (tile.balance = target number of non-diagonal gates per pass).  Gates on the always-local low bits
            # can run in ANY pass; taking all of them as soon as they are ready piles the low 10 bits of several layers of
            # a ladder circuit into one shared-memory-bound pass while the passes around it wait for HBM.  Re-select with
            # the gates that need the window first and only as many low-bits-only gates as fit the target: the rest ride
            # in later passes (their dependants lag by one period, nothing is lost).
            best = None
            cands = [(-1,)]
            jj = first
            while jj < N and (done[jj] or is_diag[jj] or not needs[jj]):
                jj += 1
            if jj < N and needs[jj] and h >= 2 and n - tile.L > h:
                nb2 = needs[jj]
                lo2 = (nb2 & -nb2).bit_length() - 1
                hi2 = nb2.bit_length() - 1
                cands += [(((1 << h) - 1) << st_,) for st_ in range(max(tile.L, hi2 - h + 1), min(lo2, n - h) + 1)]
            for (w_,) in cands:
                c0, H0, nH0, nd0 = select(first, w_, 0)          # window gates only
                quota = max(0, tile.balance - nd0)
                c1, H1, nH1, nd1 = select(first, w_, quota) if quota else (c0, H0, nH0, nd0)
                score = (min(nd1, tile.balance), nd0)
                if best is None or score > best[0]:
                    best = (score, c1, H1, nH1, nd1)
            if best is not None and best[4] > 0:
                _, chosen, H, nH, nd = best
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


def _prepare(gates: Sequence[LGate], n: int, tile: TileConfig, chain: Optional[bool]):
    """Effective tile, the pass schedule (diagonals sunk to their consumers when chains are grouped) and the chain switch."""
    m_eff = min(tile.m, n)
    tile = TileConfig(m=m_eff, L=min(tile.L, m_eff), threads=tile.threads, ctas_per_sm=tile.ctas_per_sm,
                      max_gates=tile.max_gates, rot_layers=tile.rot_layers, max_work=tile.max_work, balance=tile.balance)
    sched = schedule(gates, n, tile)
    if chain is None:
        chain = bool(getattr(gates, "chain_after_schedule", False))
    if chain:
        from .fuse import sink_diagonals
        sched = sink_diagonals(gates, sched)
    return tile, sched, chain


def _group(gates: Sequence[LGate], sched, tile: TileConfig):
    """Group the 1-qubit / MUX gates of every pass into CHAIN gates now that their targets are known to be
    tile-local together (``order`` then no longer maps compiled gates to input gates)."""
    from .fuse import group_pass
    grouped: List[LGate] = []
    sched2 = []
    for hb, chosen in sched:
        sub = group_pass([gates[i] for i in chosen], set(range(tile.L)) | set(int(p) for p in hb), R_rot=tile.rot_layers)
        sched2.append((hb, list(range(len(grouped), len(grouped) + len(sub)))))
        grouped += sub
    return grouped, sched2


def compile_program(gates: Sequence[LGate], n: int, tile: TileConfig, *, batch_mats: int = 1, itemsize: int = 16,
                    chain: Optional[bool] = None) -> Program:
    """Pack scheduled gates into the ABI arrays.  ``batch_mats`` > 1: every gate's ``data`` has a
    leading batch axis (one matrix per batch member) when ``gate.batched`` is set.  ``itemsize`` (bytes per amplitude)
    only decides which passes ask for the padded tile layout."""
    tile, sched, chain = _prepare(gates, n, tile, chain)
    if chain:
        gates, sched = _group(gates, sched, tile)
    return _pack(gates, sched, n, tile, batch_mats, itemsize)


def compile_program_stream(gates: Sequence[LGate], n: int, tile: TileConfig, *, batch_mats: int = 1, itemsize: int = 16,
                           chain: Optional[bool] = None, first: int = 8, chunk: int = 32):
    """The same passes as ``compile_program`` as a sequence of Programs (``first`` passes, then ``chunk`` at a time), so
    that a caller can launch a chunk while the host groups and packs the next one: scheduling is global and happens
    before the first chunk, chain grouping and packing are per pass.  Concatenating the chunks' passes gives exactly
    compile_program's passes."""
    tile, sched, chain = _prepare(gates, n, tile, chain)
    lo = 0
    size = max(1, int(first))
    while lo < len(sched):
        part = sched[lo:lo + size]
        if chain:
            g2, part = _group(gates, part, tile)
        else:
            # re-index the chunk's gates from 0 so that the chunk is a self-contained Program
            idx = [i for _, c in part for i in c]
            pos = {i: j for j, i in enumerate(idx)}
            g2, part = [gates[i] for i in idx], [(hb, [pos[i] for i in c]) for hb, c in part]
        yield _pack(g2, part, n, tile, batch_mats, itemsize)
        lo += size
        size = max(1, int(chunk))


def _pack(gates: Sequence[LGate], sched, n: int, tile: TileConfig, batch_mats: int, itemsize: int) -> Program:
    m_eff = tile.m
    ng = sum(len(c) for _, c in sched)
    passes = np.zeros(len(sched), dtype=_lib.PASS_DTYPE)
    garr = np.zeros(ng, dtype=_lib.GATE_DTYPE)
    mats: List[np.ndarray] = []
    mat_off = 0
    gi = 0
    order: List[int] = []
    L = tile.L
    for pi, (hb, chosen) in enumerate(sched):
        assert len(hb) == m_eff - L, (hb, m_eff, L)
        local_of = {p: p for p in range(L)}
        for j, p in enumerate(hb):
            local_of[p] = L + j

        def enc(b: int) -> int:  # tile-local position, or 64 + index bit when outside the tile
            return local_of[b] if b in local_of else 64 + b

        ps = passes[pi]
        ps["m"] = m_eff
        ps["L"] = L
        ps["gate_begin"] = gi
        ps["n_gates"] = len(chosen)
        ps["mat_begin"] = mat_off
        for j, p in enumerate(hb):
            ps["hb"][j] = p
        maxk = 0
        any_batched = False
        lean = True   # only 1-qubit-layer gates: DENSE k = 1, DIAG, MUX, CHAIN (the lean kernel variant)
        low_regs = False   # a gate keeps index bits below 128 bytes in registers: bank conflicts unless the tile is padded
        low_bits = 3 if itemsize == 16 else 4
        for idx in chosen:
            g = gates[idx]
            if not (g.kind in (DIAG, MUX, CHAIN) or (g.kind == DENSE and g.k == 1)):
                lean = False
            e = garr[gi]
            e["kind"] = g.kind
            e["k"] = g.k
            if g.kind != DIAG and any(local_of.get(b, 99) < low_bits for b in g.bits):
                low_regs = True
            if g.kind == DIAG:
                assert g.k <= 6, "diagonal tables are limited to 6 bits"
                for j, b in enumerate(g.bits):
                    e["bits"][j] = enc(b)
            elif g.kind == MUX:
                e["k"] = 1
                e["bits"][0] = local_of[g.bits[0]]
                e["bits"][1] = enc(g.bits[1])
            elif g.kind == CHAIN:
                has_c, E = g.pat_a & 1, g.pat_a >> 1
                r = len(g.bits) - has_c - E
                e["k"] = r
                loc = [local_of[b] for b in g.bits[:r]]
                for j, b in enumerate(loc):
                    e["bits"][j] = b
                ctrl = enc(g.bits[r]) if has_c else 127
                e["bits"][r] = ctrl
                for j in range(2):      # extra index bits of the pre-diagonal table (rotation form only)
                    if r + 1 + j < 8:
                        e["bits"][r + 1 + j] = enc(g.bits[r + has_c + j]) if j < E else 127
                e["off_a"] = g.pat_b & 15          # 0 = two matrices per layer; 4..7 = rotation form (tqb_core.cuh gate_chain_rot)
                e["off_b"] = (E & 3) | (g.pat_b & 128) | (g.pat_b & 0x1F00)   # extras, unit table, forms of the scaled layers
                zs = sorted(loc + ([ctrl] if ctrl < 64 else []))
                for j, b in enumerate(zs):
                    e["sbits"][j] = b
            else:
                assert g.k <= 4, "dense / pair gates are limited to 4 bits"
                loc = [local_of[b] for b in g.bits]
                for j, b in enumerate(loc):
                    e["bits"][j] = b
                for j, b in enumerate(sorted(loc)):
                    e["sbits"][j] = b
                if g.kind in (PAIR, SWAP):
                    e["off_a"] = sum(((g.pat_a >> j) & 1) << loc[j] for j in range(g.k))
                    e["off_b"] = sum(((g.pat_b >> j) & 1) << loc[j] for j in range(g.k))
                    e["zmask"] = g.zmask
                else:
                    maxk = max(maxk, g.k)
            d = np.asarray(g.data, dtype=C128)
            if batch_mats > 1 and g.batched:
                any_batched = True
                d = d.reshape(batch_mats, -1)
                e["mat_off"] = mat_off
                e["mat_bstride"] = d.shape[1]
                mats.append(d.reshape(-1))
            else:
                d = d.reshape(-1)
                e["mat_off"] = mat_off
                e["mat_bstride"] = 0
                mats.append(d)
            mat_off += d.size
            order.append(idx)
            gi += 1
        cnt = mat_off - int(ps["mat_begin"])
        ps["mat_count"] = cnt if (not any_batched and cnt <= 4096) else 0   # staged in shared memory
        # -1: lean-eligible pass, matrices staged or per batch member (tqb_run_passes picks tile_pass_lean_kernel when the tile streams)
        # -2: the same with the padded tile layout (tqb_core.cuh pidx)
        ps["max_dense_k"] = (-2 if low_regs else -1) if (lean and (any_batched or (ps["mat_count"] > 0 and cnt <= 1024))) else maxk
    flat = np.concatenate(mats) if mats else np.zeros(0, dtype=C128)
    return Program(n=n, passes=passes, gates=garr, mats=flat, n_gates_in=len(gates), tile=tile, order=order)


def fp_ops_per_amplitude(prog: Program) -> float:
    """Floating-point pipe instructions (multiply, add or fused multiply-add, one lane each) the tile kernel issues per
    amplitude over the whole program: the arithmetic side of the roofline (bench.py reports it against the measured
    FP64 / FP32 instruction rate).  DENSE k: 2^k complex MACs per amplitude = 4 * 2^k; DIAG / MUX-selected 2x2 /
    general CHAIN layer: 4 / 8 / 8; rotation-form CHAIN: 4 per layer + 4 for the pre-diagonal table unless it is all
    ones; PAIR touches 2 of 2^k patterns with a 2x2; SWAP moves data only."""
    from . import _lib
    total = 0.0
    for g in prog.gates:
        kind, k = int(g["kind"]), int(g["k"])
        if kind == _lib.GATE_DENSE:
            total += 4.0 * (1 << k)
        elif kind == _lib.GATE_DIAG:
            total += 4.0
        elif kind == _lib.GATE_PAIR:
            total += 8.0 * 2.0 / (1 << k)
        elif kind == _lib.GATE_MUX:
            total += 8.0
        elif kind == _lib.GATE_CHAIN:
            if int(g["off_a"]) >= 4:
                total += 4.0 + 2.0 * (k - 1) + (0.0 if int(g["off_b"]) & 128 else 4.0)   # layer 0: 4, scaled layers: 2, table: 4
            else:
                total += 8.0 * k
    return total
