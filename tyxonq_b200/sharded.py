"""States too large for one GPU: shard by the high-order index bits, one process per GPU.

The reference has no distributed path at all (zero collectives).  Design:

  * rank r of G = 2^g owns the amplitudes whose g highest PHYSICAL index bits equal r
    (n_local = n - g local bits).  A logical->physical bit map is tracked on the host, so "which
    qubits are global" changes as the circuit runs; TyxonQ's big-endian convention (qubit 0 = most
    significant bit, libs/quantum_library/kernels/statevector.py:28-42) fixes the INITIAL map only.
  * Gates run as ordinary local fused passes with ``global_base = rank << n_local``: diagonal tables
    and MUX/CHAIN controls read the rank bits from global_base, so diagonal gates and controls on
    global qubits need no communication.
  * A gate whose target is a global bit forces a *global <-> local exchange*: the g global bits
    are swapped with the g top local bits by ONE all-to-all of contiguous chunks over NVLink
    (each rank keeps 1/G of its shard, sends (G-1)/G); victims (the local bits that become global)
    are chosen by farthest next use and moved to the top local positions by in-tile SWAP gates
    that ride in the preceding local passes.
  * Reductions (<Z_q>, norms, energies) are local reductions + one all-reduce of a few doubles.

The only data-path collective is that all-to-all.  ``exchange`` uses NCCL's all_to_all_single on
GPUs and grouped send/recv elsewhere (gloo has no all-to-all), which is what the CPU tests run.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Any, Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .gates import CHAIN, DENSE, DIAG, MUX, PAIR, SWAP, X_MAT, LGate


# ------------------------------------------------------------------------------------------
# host planning
# ------------------------------------------------------------------------------------------
def _bits(mask: int) -> List[int]:
    out = []
    b = 0
    while mask:
        if mask & 1:
            out.append(b)
        mask >>= 1
        b += 1
    return out


def remap_gate(g: LGate, phys: Sequence[int]) -> LGate:
    """Same gate on physical bit positions (phys[logical bit] = physical bit)."""
    ng = LGate(g.kind, tuple(phys[b] for b in g.bits), g.data, pat_a=g.pat_a, pat_b=g.pat_b,
               zmask=_remap_mask(g.zmask, phys), param=g.param, name=g.name)
    return ng


def _remap_mask(mask: int, phys: Sequence[int]) -> int:
    out = 0
    for b in _bits(mask):
        out |= 1 << phys[b]
    return out


@dataclass
class Segment:
    gates: List[LGate]            # on PHYSICAL bits; bits >= n_local are rank bits (controls / diagonals only)
    exchange_after: bool = False  # all-to-all of the g global bits with the g top local bits after the gates


@dataclass
class ShardPlan:
    n: int
    g: int
    segments: List[Segment]
    final_phys: List[int]         # final_phys[logical bit] = physical bit after the last segment
    n_exchanges: int = 0


def plan_sharded(gates: Sequence[LGate], n: int, g: int, phys0: Optional[Sequence[int]] = None) -> ShardPlan:
    """Split a gate list (logical index bits) into local segments separated by global<->local exchanges.
    ``phys0``: the logical->physical map the state is in when the plan starts (default: identity)."""
    n_local = n - g
    phys = list(range(n)) if phys0 is None else list(phys0)      # logical -> physical
    remaining = list(range(len(gates)))
    segments: List[Segment] = []
    n_ex = 0
    if g == 0:
        return ShardPlan(n, 0, [Segment(list(gates))], phys)
    guard = 0
    while remaining:
        guard += 1
        if guard > 4 * len(gates) + 8:
            raise RuntimeError("sharded planner failed to make progress")
        chosen: List[int] = []
        rest: List[int] = []
        blocked = 0
        blocked_nd = 0
        for i in remaining:
            gt = gates[i]
            mk = gt.mask
            if gt.kind == DIAG:
                if mk & blocked_nd:
                    blocked |= mk
                    rest.append(i)
                else:
                    chosen.append(i)
                continue
            ok = not (mk & blocked) and all(phys[b] < n_local for b in _bits(gt.local_mask))
            if ok:
                chosen.append(i)
            else:
                blocked |= mk
                blocked_nd |= mk
                rest.append(i)
        seg = Segment([remap_gate(gates[i], phys) for i in chosen])
        remaining = rest
        if remaining:
            # choose the g local logical bits that become global: farthest next use as a tile-local bit,
            # never a bit the first blocked gate needs
            inv = {p: l for l, p in enumerate(phys)}
            first_need = gates[remaining[0]].local_mask
            next_use = {}
            for pos, i in enumerate(remaining):
                for b in _bits(gates[i].local_mask):
                    next_use.setdefault(b, pos)
            local_logical = [l for l in range(n) if phys[l] < n_local]
            cands = [l for l in local_logical if not (first_need >> l) & 1]
            cands.sort(key=lambda l: -next_use.get(l, 1 << 30))
            victims = cands[:g]
            if len(victims) < g:
                raise RuntimeError("not enough local qubits to exchange")
            # move the victims to the top g local positions with in-tile SWAP gates (physical bit swaps)
            top = list(range(n_local - g, n_local))
            vict_set = set(victims)
            free_top = [p for p in top if inv[p] not in vict_set]
            for v in victims:
                pv = phys[v]
                if pv >= n_local - g:
                    continue
                pt = free_top.pop()
                other = inv[pt]
                seg.gates.append(LGate(SWAP, (pv, pt), X_MAT.reshape(4).copy(), pat_a=0b01, pat_b=0b10, name="bitswap"))
                phys[v], phys[other] = pt, pv
                inv[pt], inv[pv] = v, other
            seg.exchange_after = True
            n_ex += 1
            # the all-to-all swaps physical bit (n_local - g + i) with physical bit (n_local + i)
            for i in range(g):
                a, b = n_local - g + i, n_local + i
                la, lb = inv[a], inv[b]
                phys[la], phys[lb] = b, a
                inv[a], inv[b] = lb, la
        segments.append(seg)
    return ShardPlan(n, g, segments, phys, n_ex)


def _exchange_segment(phys: List[int], victims: Sequence[int], n: int, g: int) -> Segment:
    """Segment that moves the logical bits ``victims`` (local, len g) to the top g local positions with in-tile bit
    swaps and then exchanges them with the g rank bits; ``phys`` is updated in place."""
    n_local = n - g
    inv = {p: l for l, p in enumerate(phys)}
    seg = Segment([])
    vict_set = set(victims)
    free_top = [p for p in range(n_local - g, n_local) if inv[p] not in vict_set]
    for v in victims:
        pv = phys[v]
        if pv >= n_local - g:
            continue
        pt = free_top.pop()
        other = inv[pt]
        seg.gates.append(LGate(SWAP, (pv, pt), X_MAT.reshape(4).copy(), pat_a=0b01, pat_b=0b10, name="bitswap"))
        phys[v], phys[other] = pt, pv
        inv[pt], inv[pv] = v, other
    seg.exchange_after = True
    for i in range(g):
        a, b = n_local - g + i, n_local + i
        la, lb = inv[a], inv[b]
        phys[la], phys[lb] = b, a
        inv[a], inv[b] = lb, la
    return seg


def plan_localize(phys_in: Sequence[int], need_mask: int, n: int, g: int, cost: Optional[Sequence[int]] = None) -> ShardPlan:
    """Exchanges that make every logical bit of ``need_mask`` local: the all-to-all brings all g rank bits home and
    sends g local victims away.  Victims are local logical bits outside ``need_mask`` with the smallest ``cost``
    (e.g. how many pending Pauli groups flip that bit).  No segment if the mask is already local; one exchange when
    g bits outside the mask are local now; two when some of them sit on rank bits themselves (the first exchange sends
    g mask bits away to bring every rank bit home)."""
    n_local = n - g
    phys = list(phys_in)
    if g == 0 or all(phys[b] < n_local for b in _bits(need_mask)):
        return ShardPlan(n, g, [], phys, 0)
    if n - bin(need_mask).count("1") < g:
        raise RuntimeError("mask touches more than n - g qubits: it cannot be made local")
    segments: List[Segment] = []
    cands = [l for l in range(n) if phys[l] < n_local and not (need_mask >> l) & 1]
    if len(cands) < g:
        inside = [l for l in range(n) if phys[l] < n_local and (need_mask >> l) & 1]
        if len(inside) < g:
            raise RuntimeError("not enough local qubits to exchange")
        segments.append(_exchange_segment(phys, inside[:g], n, g))
        cands = [l for l in range(n) if phys[l] < n_local and not (need_mask >> l) & 1]
    cands.sort(key=lambda l: (cost[l] if cost is not None else 0, -phys[l]))
    segments.append(_exchange_segment(phys, cands[:g], n, g))
    return ShardPlan(n, g, segments, phys, len(segments))


def plan_restore(phys_in: Sequence[int], n: int, g: int) -> ShardPlan:
    """Segments that bring an arbitrary logical->physical map back to the identity (needed before sampling: the
    blocked CDF of the reference runs over LOGICAL index order = rank-major order of an identity layout).
    At most two exchanges: one that brings every currently-global bit home when a target bit is global in
    the wrong place, one that sends the g highest logical bits to their rank positions; then in-tile bit swaps
    sort the local bits."""
    n_local = n - g
    phys = list(phys_in)
    inv = {p: l for l, p in enumerate(phys)}
    segments: List[Segment] = []
    n_ex = 0

    def swap_phys(seg: Segment, pa: int, pb: int) -> None:
        if pa == pb:
            return
        seg.gates.append(LGate(SWAP, (pa, pb), X_MAT.reshape(4).copy(), pat_a=0b01, pat_b=0b10, name="bitswap"))
        la, lb = inv[pa], inv[pb]
        phys[la], phys[lb] = pb, pa
        inv[pa], inv[pb] = lb, la

    def do_exchange(seg: Segment) -> None:
        seg.exchange_after = True
        for i in range(g):
            a, b = n_local - g + i, n_local + i
            la, lb = inv[a], inv[b]
            phys[la], phys[lb] = b, a
            inv[a], inv[b] = lb, la

    targets = [n_local + i for i in range(g)]
    if g and any(phys[t] != t for t in targets):
        if any(phys[t] >= n_local for t in targets):
            # a target sits on a rank bit: the all-to-all swaps ALL g rank bits, so first bring them all home,
            # sending g local non-target bits away
            seg = Segment([])
            fillers = [l for l in range(n) if phys[l] < n_local and l not in targets][:g]
            if len(fillers) < g:
                raise RuntimeError("not enough local qubits to restore the layout")
            for i, f in enumerate(fillers):
                swap_phys(seg, phys[f], n_local - g + i)
            do_exchange(seg)
            segments.append(seg)
            n_ex += 1
        seg = Segment([])
        for i, t in enumerate(targets):
            swap_phys(seg, phys[t], n_local - g + i)
        do_exchange(seg)
        segments.append(seg)
        n_ex += 1
    seg = Segment([])
    for l in range(n_local):
        swap_phys(seg, phys[l], l)
    if seg.gates:
        segments.append(seg)
    assert phys == list(range(n)), phys
    return ShardPlan(n, g, segments, phys, n_ex)


# ------------------------------------------------------------------------------------------
# the exchange
# ------------------------------------------------------------------------------------------
_LANES: dict = {}   # group -> extra NCCL communicators over the same ranks (parallel lanes of one exchange)


def _exchange_lanes(group: Any, world: int) -> list:
    """With few ranks one all-to-all is one send/recv pair per peer and NCCL leaves most of NVLink idle (2 ranks: 0.52 of
    the peer-copy peak).  Extra communicators over the same ranks let slices of the exchange run as concurrent collectives.
    TQB_EXCHANGE_LANES overrides the count (1 = a single all-to-all)."""
    import os
    key = id(group)
    if key not in _LANES:
        want = int(os.environ.get("TQB_EXCHANGE_LANES", "0")) or (4 if world <= 2 else (2 if world <= 4 else 1))
        ranks = dist.get_process_group_ranks(group) if group is not None else list(range(dist.get_world_size()))
        _LANES[key] = [dist.new_group(ranks=ranks, backend="nccl") for _ in range(max(0, want - 1))]
    return _LANES[key]


def exchange(out: torch.Tensor, inp: torch.Tensor, group: Any = None) -> None:
    """out[j] <- chunk `rank` of rank j's inp, for inp/out viewed as [G, chunk] (the global<->local swap)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    inp2 = inp.view(world, -1)
    out2 = out.view(world, -1)
    if dist.get_backend(group) == "nccl":
        # complex tensors go over the wire as reals
        o = torch.view_as_real(out2).reshape(world, -1)
        i = torch.view_as_real(inp2).reshape(world, -1)
        lanes = _exchange_lanes(group, world)
        k = len(lanes) + 1
        cols = o.shape[1]
        if k == 1 or cols % k or cols < (1 << 20):
            dist.all_to_all_single(o, i, group=group)
            return
        # lane l moves columns [l, l+1) * cols / k of every chunk: k concurrent all-to-alls on k communicators
        w = cols // k
        works = list(enumerate([group] + lanes))
        # slices of a [world, cols] view are strided: exchange them through split lists (no staging copies)
        hs = []
        for l, g_l in works:
            out_list = [o[j, l * w:(l + 1) * w] for j in range(world)]
            in_list = [i[j, l * w:(l + 1) * w] for j in range(world)]
            hs.append(dist.all_to_all(out_list, in_list, group=g_l, async_op=True))
        for h in hs:
            h.wait()
        return
    ops = []
    for j in range(world):
        if j == rank:
            out2[j].copy_(inp2[j])
            continue
        ops.append(dist.P2POp(dist.isend, torch.view_as_real(inp2[j]).contiguous(), j, group))
        ops.append(dist.P2POp(dist.irecv, torch.view_as_real(out2[j]), j, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


# ------------------------------------------------------------------------------------------
# execution
# ------------------------------------------------------------------------------------------
class ShardedState:
    """One rank's shard plus the scratch buffer of the exchange (ping-pong: no copy back)."""

    def __init__(self, n: int, dtype: torch.dtype, device: torch.device, group: Any = None,
                 backend: Optional[Any] = None) -> None:
        """``backend``: object with init_local / run_local / reduce_local (default: the CUDA kernels of this
        package).  The CPU test tier injects an oracle-backed one to exercise the planner and the exchange
        over gloo; the package itself has no CPU compute path."""
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.group = group
        self.g = int(self.world).bit_length() - 1
        if (1 << self.g) != self.world:
            raise ValueError("world size must be a power of two")
        self.n = n
        self.n_local = n - self.g
        self.dtype = dtype
        self.device = torch.device(device)
        self.state = torch.empty(1 << self.n_local, dtype=dtype, device=self.device)
        self.scratch = torch.empty_like(self.state) if self.g else None
        self.global_base = self.rank << self.n_local
        self.phys = list(range(n))
        self.programs: List[Any] = []
        self.backend = backend or _CudaLocal(self)
        self.exchange_s = 0.0
        self.pauli_exchanges = 0           # exchanges the last expect_pauli_sum needed

    def init_zero(self) -> None:
        self.backend.init_local(self.state, self.n_local, self.global_base)
        self.phys = list(range(self.n))

    def run(self, plan: ShardPlan, cache: bool = True) -> None:
        """``cache``: keep the compiled per-segment programs (slot = segment index) for the next run of the SAME plan."""
        for si, seg in enumerate(plan.segments):
            if seg.gates:
                self.backend.run_local(self.state, seg.gates, self.n_local, self.global_base, si if cache else None)
            if seg.exchange_after:
                t0 = time.perf_counter()
                exchange(self.scratch, self.state, self.group)
                self.state, self.scratch = self.scratch, self.state
                self.exchange_s += time.perf_counter() - t0
        self.phys = list(plan.final_phys)

    def apply(self, gates: Sequence[LGate]) -> int:
        """Apply more gates (logical bits) to the state in WHATEVER layout it is in now; returns the exchanges used."""
        plan = plan_sharded(list(gates), self.n, self.g, phys0=self.phys)
        self.run(plan, cache=False)
        return plan.n_exchanges

    # -- sampling -------------------------------------------------------------------------
    def restore_layout(self) -> int:
        """Bring the shard layout back to the identity map (logical bit l = physical bit l); returns the number
        of exchanges it took (0-2)."""
        if self.phys == list(range(self.n)):
            return 0
        plan = plan_restore(self.phys, self.n, self.g)
        self.run(plan, cache=False)
        return plan.n_exchanges

    def sample(self, uniforms: torch.Tensor) -> torch.Tensor:
        """Basis-state indices (int64 [shots], logical index, identical on every rank) drawn with the given host
        uniforms: the blocked-CDF contract of the single-GPU sampler (oracle/sv_oracle.py sample_indices; reference
        engine.py:377-418), bit for bit.  Every rank sums its own chunks, the chunk totals are all-gathered in rank
        order, every rank runs the same sequential prefix over all chunks and resolves the samples that fall into
        its shard; a max-reduction assembles the result."""
        self.restore_layout()
        be = self.backend
        u = uniforms.to(device=self.device, dtype=torch.float64).contiguous().reshape(-1)
        totals = be.chunk_totals(self.state, self.n_local)
        ncl = int(totals.numel())
        if self.world > 1:
            parts = [torch.empty_like(totals) for _ in range(self.world)]
            dist.all_gather(parts, totals, group=self.group)
            totals_all = torch.cat(parts)
        else:
            totals_all = totals
        prefix = be.chunk_prefix(totals_all)
        last = self.rank == self.world - 1
        idx = be.sample_shard(self.state, self.n_local, prefix, self.rank * ncl, (1 << self.n) if last else -1, u)
        if self.world > 1:
            dist.all_reduce(idx, op=dist.ReduceOp.MAX, group=self.group)
        return idx

    def counts(self, uniforms: torch.Tensor) -> dict:
        """{bitstring: count} over all n qubits (big-endian, engine.py:418-463) from ``sample``."""
        idx = self.sample(uniforms).cpu().numpy()
        idx = np.minimum(idx, (1 << self.n) - 1)
        vals, cnt = np.unique(idx, return_counts=True)
        return {format(int(v), f"0{self.n}b"): int(c) for v, c in zip(vals, cnt)}

    # -- reductions -----------------------------------------------------------------------
    def expect_z_all(self) -> torch.Tensor:
        """<Z> for every LOGICAL index bit (float64 [n]); one local read + one all-reduce."""
        zl, nrm = self.backend.reduce_local(self.state, self.n_local)
        zphys = torch.empty(self.n, dtype=torch.float64, device=self.device)
        zphys[: self.n_local] = zl
        for i in range(self.g):
            zphys[self.n_local + i] = nrm if not (self.rank >> i) & 1 else -nrm
        if self.world > 1:
            dist.all_reduce(zphys, group=self.group)
        return zphys[torch.tensor(self.phys, device=self.device)]


    def expect_zmasks(self, masks: Sequence[int]) -> torch.Tensor:
        """<prod_{b in mask} Z_b> for masks over LOGICAL index bits (qubit q = bit n-1-q), float64 [len(masks)] on every
        rank: the masks are remapped to physical bits (rank bits included: their parity comes from global_base), one
        local read + one all-reduce.  Covers diagonal Hamiltonians (ZZ couplings, fields) without any exchange."""
        pm = [_remap_mask(int(m), self.phys) for m in masks]
        out = self.backend.zmasks_local(self.state, pm, self.global_base)
        if self.world > 1:
            dist.all_reduce(out, group=self.group)
        return out

    def expect_pauli_sum(self, ps: Any) -> torch.Tensor:
        """<psi|H|psi> for a ``pauli.PauliSum`` over LOGICAL qubits (complex128 scalar tensor, same on every rank) -- the
        sharded form of ``PauliSum.expectation`` (reference: kernels/pauli.py:74-87 + dynamics.py:117-126 on one device).
        A term flips the bits of its xmask, so its partner amplitude lives on this rank iff the xmask is local in the
        current layout; Z factors on rank bits only need ``global_base``.  Groups are evaluated layout by layout: all
        groups whose xmask is local now go through one local kernel call, then ONE exchange (``plan_localize``) brings
        the rank bits home and sends away the g local bits that the fewest pending groups flip, until none is pending.
        An xmask on more than n - g qubits can never be local: its terms are measured after rotating g of its X/Y
        factors into Z (H, or Sdg then H), and the rotation is undone afterwards.  One all-reduce at the end.
        The layout (``self.phys``) may differ afterwards; the state itself is unchanged."""
        from .pauli import PauliSum
        if int(ps.n) != self.n:
            raise ValueError(f"PauliSum on {ps.n} qubits, state on {self.n}")
        n, n_local, g = self.n, self.n_local, self.g
        gx = [int(x) for x in ps.group_x]
        pending = list(range(len(gx)))
        acc = torch.zeros(2, dtype=torch.float64, device=self.device)
        self.pauli_exchanges = 0

        def terms_of(gi: int) -> List[Tuple[int, int, complex]]:
            a, b = int(ps.group_ptr[gi]), int(ps.group_ptr[gi + 1])
            return [(gx[gi], int(ps.term_z[t]), complex(ps.term_coef[t])) for t in range(a, b)]

        def eval_local(terms: List[Tuple[int, int, complex]]) -> None:
            nonlocal acc
            sub = PauliSum(n_local, [(_remap_mask(x, self.phys), _remap_mask(z, self.phys), c) for x, z, c in terms])
            acc = acc + self.backend.pauli_local(self.state, sub, self.global_base)

        def localize(mask: int, groups: Sequence[int]) -> None:
            cost = [sum((gx[gi] >> l) & 1 for gi in groups) for l in range(n)]
            plan = plan_localize(self.phys, mask, n, g, cost)
            if plan.segments:
                self.run(plan, cache=False)
                self.pauli_exchanges += plan.n_exchanges

        wide = [gi for gi in pending if bin(gx[gi]).count("1") > n_local]
        pending = [gi for gi in pending if gi not in set(wide)]
        while pending:
            now = [gi for gi in pending if _remap_mask(gx[gi], self.phys) < (1 << n_local)]
            if now:
                eval_local([t for gi in now for t in terms_of(gi)])
                done = set(now)
                pending = [gi for gi in pending if gi not in done]
                continue
            localize(gx[pending[0]], pending)
        for gi in wide:
            by_basis: dict = {}
            rot = _bits(gx[gi])[-g:]                     # g of the flipped bits: these factors are rotated to Z
            rmask = sum(1 << b for b in rot)
            for x, z, c in terms_of(gi):
                by_basis.setdefault(z & rmask, []).append((x, z, c))
            for ymask, terms in by_basis.items():
                fwd, bwd = [], []
                for b in rot:
                    q = n - 1 - b
                    if (ymask >> b) & 1:
                        fwd += [("sdg", q), ("h", q)]
                        bwd += [("h", q), ("s", q)]
                    else:
                        fwd += [("h", q)]
                        bwd += [("h", q)]
                ny = bin(ymask).count("1")
                self.pauli_exchanges += self.apply(lower_and_fuse(fwd, n))
                rterms = [(x & ~rmask, (z & ~rmask) | rmask, c / (1j ** ny)) for x, z, c in terms]
                localize(gx[gi] & ~rmask, [])
                eval_local(rterms)
                self.pauli_exchanges += self.apply(lower_and_fuse(bwd, n))
        if self.world > 1:
            dist.all_reduce(acc, group=self.group)
        return torch.view_as_complex(acc.reshape(1, 2))[0]


class _CudaLocal:
    """Local work of one rank on its GPU: the fused passes and reductions of libtyxonq_b200.so."""

    def pauli_local(self, state: torch.Tensor, sub: Any, global_base: int) -> torch.Tensor:
        """(re, im) float64 [2] of sum_j conj(psi_{j^x}) phase(global_base | j) psi_j over this shard (tqb_expect_pauli_sum)."""
        return torch.view_as_real(sub.expectation(state, global_base=global_base)).reshape(-1)[:2]

    def zmasks_local(self, state: torch.Tensor, masks: Sequence[int], global_base: int) -> torch.Tensor:
        from . import program as P
        return P.expect_zmasks(state, list(masks), global_base=global_base)[0]

    def __init__(self, owner: "ShardedState") -> None:
        self.owner = owner

    def init_local(self, state: torch.Tensor, n_local: int, global_base: int) -> None:
        from . import _lib
        from . import program as P
        ptr, n_, b_, dt, stream = P._prep(state)
        _lib.check(_lib.load().tqb_init_basis(ptr, n_, 1, dt, global_base, 0, stream))

    def run_local(self, state: torch.Tensor, gates: List[LGate], n_local: int, global_base: int, cache_slot: int) -> None:
        from . import program as P
        from .planner import compile_program, default_tile
        if cache_slot is None:
            prog = compile_program(gates, n_local, default_tile(n_local, state.element_size(), 1), itemsize=state.element_size(), chain=True)
            P.DeviceProgram(prog, state.device, state.dtype).run(state, global_base=global_base)
            return
        progs = self.owner.programs
        while len(progs) <= cache_slot:
            progs.append(None)
        if progs[cache_slot] is None:
            prog = compile_program(gates, n_local, default_tile(n_local, state.element_size(), 1), itemsize=state.element_size(), chain=True)
            progs[cache_slot] = P.DeviceProgram(prog, state.device, state.dtype)
        progs[cache_slot].run(state, global_base=global_base)

    def reduce_local(self, state: torch.Tensor, n_local: int) -> Tuple[torch.Tensor, torch.Tensor]:
        from . import program as P
        return P.expect_z_bits(state)[0], P.norm2(state)[0]

    def chunk_totals(self, state: torch.Tensor, n_local: int) -> torch.Tensor:
        from . import _lib
        from . import program as P
        if n_local < 12:
            raise _lib.TqbError("sharded sampling needs shards of at least one 4096-amplitude chunk (n_local >= 12)")
        ptr, n_, _, dt, stream = P._prep(state)
        out = torch.empty(1 << (n_local - 12), dtype=torch.float64, device=state.device)
        with torch.cuda.device(state.device):
            _lib.check(_lib.load().tqb_chunk_totals(ptr, n_, 1, dt, out.data_ptr(), stream))
        return out

    def chunk_prefix(self, totals_all: torch.Tensor) -> torch.Tensor:
        from . import _lib
        nc = int(totals_all.numel())
        out = torch.empty(nc + 1, dtype=torch.float64, device=totals_all.device)
        with torch.cuda.device(totals_all.device):
            _lib.check(_lib.load().tqb_chunk_prefix(totals_all.data_ptr(), nc, 1, out.data_ptr(), _lib.current_stream_ptr(totals_all.device)))
        return out

    def sample_shard(self, state: torch.Tensor, n_local: int, prefix: torch.Tensor, chunk_first: int, tail_index: int,
                     uniforms: torch.Tensor) -> torch.Tensor:
        from . import _lib
        from . import program as P
        ptr, n_, _, dt, stream = P._prep(state)
        idx = torch.empty(uniforms.numel(), dtype=torch.int64, device=state.device)
        with torch.cuda.device(state.device):
            _lib.check(_lib.load().tqb_sample_shard(ptr, n_, dt, prefix.data_ptr(), int(prefix.numel()) - 1, int(chunk_first),
                                                    int(tail_index), uniforms.data_ptr(), int(uniforms.numel()), idx.data_ptr(), stream))
        return idx


def lower_and_fuse(ops: Sequence[tuple], n: int, mode: str = "run") -> List[LGate]:
    from .fuse import fuse
    from .gates import lower_op
    return fuse([g for g in (lower_op(o, n, mode=mode) for o in ops) if g is not None])


class ShardedBench:
    """bench.py's N > 1 arm: the HEA workload on a state sharded over all ranks of the default group."""

    def __init__(self, n: int, ops: Sequence[tuple], dtype: torch.dtype, device: torch.device) -> None:
        self.ops = list(ops)
        self.n = n
        t0 = time.perf_counter()
        gates = lower_and_fuse(self.ops, n)
        self.st = ShardedState(n, dtype, device)
        self.plan = plan_sharded(gates, n, self.st.g)
        self.plan_ms = 1e3 * (time.perf_counter() - t0)
        self.n_local = self.st.n_local
        self._ev: List[Tuple[Any, Any]] = []
        self.st.init_zero()
        self.st.run(self.plan)          # compiles + uploads the per-segment programs
        torch.cuda.synchronize()
        passes = sum(p.prog.n_passes for p in self.st.programs if p is not None)
        B = 16 if dtype == torch.complex128 else 8
        self.info = {"passes": passes, "gates_per_pass": len([o for o in self.ops if o[0] != "measure_z"]) / max(passes, 1),
                     "plan_ms": self.plan_ms, "swaps": self.plan.n_exchanges, "n_local": self.n_local,
                     "exchange_bytes_per_gpu": self.plan.n_exchanges * ((1 << self.n_local) * B * (self.st.world - 1)) // self.st.world}
        self._local_ms = 0.0
        self._local_n = 0

    def step(self, timed: bool = False) -> torch.Tensor:
        self.st.init_zero()
        self.st.run(self.plan)
        return self.st.expect_z_all()

    def pass_ms_per_launch(self) -> float:
        """Device time of the local passes alone (one extra untimed run with events around each segment);
        also fills ``self.breakdown`` with the exchange time (NCCL kernels are on the same stream)."""
        self.st.init_zero()
        tot = 0.0
        xch = 0.0
        for si, seg in enumerate(self.plan.segments):
            if seg.gates:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                self.st.backend.run_local(self.st.state, seg.gates, self.st.n_local, self.st.global_base, si)
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            if seg.exchange_after:
                dist.barrier()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                exchange(self.st.scratch, self.st.state, self.st.group)
                e1.record()
                torch.cuda.synchronize()
                xch += e0.elapsed_time(e1)
                self.st.state, self.st.scratch = self.st.scratch, self.st.state
        t = torch.tensor([tot, xch], dtype=torch.float64, device=self.st.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot, xch = float(t[0]), float(t[1])
        nbytes = self.info["exchange_bytes_per_gpu"]
        self.breakdown = {"pass_ms": tot, "exchange_ms": xch, "exchanges": self.plan.n_exchanges,
                          "exchange_bytes_per_gpu_each_way": nbytes,
                          "nvlink_gbps_per_gpu_each_way": (nbytes / (xch * 1e-3) / 1e9) if xch > 0 else None,
                          "nvlink_peak_gbps": 770.0, "nvlink_peak_source": "measured peer copy, B200_PROFILING.md"}
        return tot / max(self.info["passes"], 1)

    def e2e(self, args: Any) -> dict:
        """Host op list -> plan -> upload -> passes/exchanges -> <Z_q> on the host, all ranks."""
        n_gates = len([o for o in self.ops if o[0] != "measure_z"])
        ts = []
        for _ in range(2):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            gates = lower_and_fuse(self.ops, self.n)
            st = self.st
            st.programs = []
            plan = plan_sharded(gates, self.n, st.g)
            st.init_zero()
            st.run(plan)
            z = st.expect_z_all().cpu()
            torch.cuda.synchronize()
            dist.barrier()
            ts.append(time.perf_counter() - t0)
        h2d = sum(p.h2d_bytes for p in self.st.programs if p is not None)
        return {"value": n_gates * 2.0 ** (self.n - 30) / float(np.mean(ts)), "unit": "gates/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(z.numel() * 8), "ms_per_step": 1e3 * float(np.mean(ts)),
                "api": "lower+fuse+plan_sharded -> ShardedState.run -> expect_z_all().cpu() on every rank"}
