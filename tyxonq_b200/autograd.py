"""Adjoint differentiation of circuit states on the device (torch.autograd.Function).

The reference's torch path keeps one 2^n tensor per gate on the autograd tape
(numerics/backends/pytorch_backend.py:446-564 via einsum nodes); its numpy path is forward
finite differences (numpy_backend.py:386-454).  Here the forward pass runs the fused passes
once and the backward pass is an adjoint sweep with O(1) state copies (model: the CI-space
sweep of applications/chem/chem_libs/quantum_chem_library/civector_ops.py:141-200):

    bra <- dL/dpsi (torch cotangent),  ket <- psi_N
    for gate j = N..1:   if parametrised:  dL/dtheta_j = Re <bra| D_j |ket>,  D_j = (dU_j/dtheta) U_j^dagger
                         ket <- U_j^dagger ket ;  bra <- U_j^dagger bra        (one batched pass)

torch's convention for a real loss is grad = dL/dRe + i dL/dIm, hence the plain Re<bra|D|ket>.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import program as P
from .gates import CHAIN, DENSE, DIAG, GEN, MUX, PAIR, SWAP, LGate, lower_op, _to_np
from .planner import compile_program, default_tile


LAYERED_MIN_QUBITS = 12   # from here on the backward pass sweeps layer by layer (layered_sweep)


def has_grad_params(circuit: Any) -> bool:
    for op in getattr(circuit, "ops", []):
        if isinstance(op, (list, tuple)):
            for a in op[1:]:
                if isinstance(a, torch.Tensor) and a.requires_grad:
                    return True
    return False


def dagger(g: LGate) -> LGate:
    if g.kind == DENSE:
        d = g.data.conj().T.copy()
    elif g.kind == DIAG:
        d = g.data.conj().copy()
    elif g.kind == SWAP:
        d = g.data.copy()
    else:
        d = np.concatenate([b.reshape(2, 2).conj().T.reshape(4) for b in g.data.reshape(-1, 4)])
    # MUX / CHAIN carry a structure tag in pat_b that does not survive the dagger ((U, X.U)^+ = (U^+, U^+.X))
    pat_b = 0 if g.kind in (MUX, CHAIN) else g.pat_b
    assert g.kind != CHAIN or g.pat_b == 0, "daggering a rotation-form CHAIN is not supported: dagger before fusion"
    return LGate(g.kind, g.bits, d, pat_a=g.pat_a, pat_b=pat_b, zmask=g.zmask, name=g.name + "^")


def grad_dense(bra: torch.Tensor, ket: torch.Tensor, bits: Sequence[int], gen: np.ndarray, scale: float,
               out: torch.Tensor, slot: int) -> None:
    """out[slot] += scale * Re <bra| gen_bits |ket>; bits[j] = index bit of matrix-index bit j."""
    ptr_b, n, _, dt, stream = P._prep(bra)
    ptr_k = ket.data_ptr()
    k = len(bits)
    b_arr = (C.c_int * k)(*[int(b) for b in bits])
    g = np.ascontiguousarray(np.asarray(gen, dtype=np.complex128).reshape(-1)).view(np.float64)
    with torch.cuda.device(bra.device):
        _lib.check(_lib.load().tqb_grad_dense(ptr_b, ptr_k, n, dt, k, C.cast(b_arr, C.c_void_p), g.ctypes.data,
                                              float(scale), out.data_ptr(), int(slot), stream))


# ---- layered reverse sweep ---------------------------------------------------------------------------------------
def _mat1q(g: LGate) -> np.ndarray:
    d = np.asarray(g.data, dtype=np.complex128)
    if g.kind == DENSE:
        return d.reshape(2, 2)
    if g.kind == DIAG:
        return np.diag(d.reshape(2))
    if g.kind == SWAP:
        return np.array([[0, 1], [1, 0]], dtype=np.complex128)
    raise NotImplementedError("single-qubit gate kind")


def transition_1q(kb: torch.Tensor, bits: Sequence[int]) -> dict:
    """{index bit: T (2x2 complex)} with T[a][b] = sum_rest conj(bra[a, rest]) ket[b, rest] for kb = [ket, bra]:
    tqb_transition_1q, up to 6 bits per read of both states (tile = low bits + the bits of interest)."""
    ptr, n, _, dt, stream = P._prep(kb[0])
    itemsize = kb.element_size()
    m = min(n, 11 if itemsize == 16 else 12)
    L = m if m == n else min(m, 5 if itemsize == 16 else 6)
    lib = _lib.load()
    todo = sorted(set(int(b) for b in bits))
    low = [b for b in todo if b < L]
    high = [b for b in todo if b >= L]
    calls = []
    while low or high:
        take_h = high[: min(6, m - L)]
        high = high[len(take_h):]
        take_l = low[: 6 - len(take_h)]
        low = low[len(take_l):]
        hb = list(take_h)
        f = L
        while len(hb) < m - L:        # fill the tile with the lowest unused high bits
            if f not in hb:
                hb.append(f)
            f += 1
        hb.sort()
        pos = {b: L + j for j, b in enumerate(hb)}
        calls.append((hb, take_l + take_h, [b if b < L else pos[b] for b in take_l + take_h]))
    res = torch.zeros((max(len(calls), 1), 48), dtype=torch.float64, device=kb.device)
    with torch.cuda.device(kb.device):
        for ci, (hb, _, tpos) in enumerate(calls):
            hb_a = (C.c_int8 * 16)(*(hb + [0] * (16 - len(hb))))
            tb_a = (C.c_int8 * 8)(*(tpos + [0] * (8 - len(tpos))))
            _lib.check(lib.tqb_transition_1q(kb[1].data_ptr(), kb[0].data_ptr(), n, dt, m, L, C.cast(hb_a, C.c_void_p),
                                             C.cast(tb_a, C.c_void_p), len(tpos), res[ci].data_ptr(), stream))
    host = res.cpu().numpy()
    out: dict = {}
    for ci, (_, bs, _) in enumerate(calls):
        for j, b in enumerate(bs):
            v = host[ci, 8 * j:8 * j + 8]
            out[b] = np.array([[v[0] + 1j * v[1], v[2] + 1j * v[3]], [v[4] + 1j * v[5], v[6] + 1j * v[7]]])
    return out


def layered_sweep(kb: torch.Tensor, gates: Sequence[LGate], slots: Sequence[Optional[Tuple[int, float]]], grad: np.ndarray) -> None:
    """Reverse sweep over kb = [ket, bra] (the states AFTER all gates), layer by layer:
    grad[slot] += weight * Re <bra| D_k |ket_k> for every gate k with slots[k] = (slot, weight), D_k = (dU_k/dtheta) U_k^+.

    The gate list is cut, from the end, into blocks in which any two gates act on different qubits or are both single-qubit
    gates on the same qubit.  Inside a block the generator of a gate, conjugated through the LATER single-qubit gates of its
    own qubit, is still a single-qubit operator A and commutes with everything else in the block, so
    <bra| D |ket> = sum_ab A[a][b] T_q[a][b] with the transition matrices T_q of the pair AFTER the block (transition_1q).
    Between blocks both states are un-applied by fused passes on the 2-member batch.  Multi-qubit parametrised gates keep
    the per-gate reduction (grad_dense)."""
    from .fuse import fuse
    ptr, n, _, dt, stream = P._prep(kb[0])
    itemsize = kb.element_size()
    blocks: List[List[int]] = []
    cur: List[int] = []
    one_q_bits = 0     # bits of single-qubit gates in the current block
    other_bits = 0     # bits of everything else in it
    for idx in range(len(gates) - 1, -1, -1):
        g = gates[idx]
        is1q = len(g.bits) == 1
        ok = not (g.mask & other_bits) if is1q else not (g.mask & (other_bits | one_q_bits))
        if not ok:
            blocks.append(cur)
            cur, one_q_bits, other_bits = [], 0, 0
        cur.append(idx)
        if is1q:
            one_q_bits |= g.mask
        else:
            other_bits |= g.mask
    if cur:
        blocks.append(cur)
    first_param = min((i for i, r in enumerate(slots) if r is not None), default=len(gates))
    pending: List[LGate] = []
    tile2 = default_tile(n, itemsize, 2)
    for blk in blocks:                      # blk: gate indices in REVERSE order
        par = [i for i in blk if slots[i] is not None]
        if par:
            if pending:                     # un-apply everything after this block: both states are then "after the block"
                prog2 = compile_program(fuse(list(pending)), n, tile2, itemsize=itemsize)
                P.DeviceProgram(prog2, kb.device, kb.dtype).run(kb)
                pending = []
            need = sorted({int(gates[i].bits[0]) for i in par if len(gates[i].bits) == 1})
            T = transition_1q(kb, need) if need else {}
            later: dict = {}                # bit -> product of the later single-qubit gates of that bit
            for i in blk:                   # reverse order: later gates first
                g = gates[i]
                if len(g.bits) == 1:
                    b = int(g.bits[0])
                    W = later.get(b, np.eye(2, dtype=np.complex128))
                    if slots[i] is not None:
                        D = np.asarray(GEN[g.name], dtype=np.complex128).reshape(2, 2)
                        A = W @ D @ W.conj().T
                        grad[slots[i][0]] += slots[i][1] * float(np.real(np.sum(A * T[b])))
                    later[b] = W @ _mat1q(g)
                elif slots[i] is not None:  # a parametrised multi-qubit gate, alone on its qubits in the block
                    out1 = torch.zeros(1, dtype=torch.float64, device=kb.device)
                    grad_dense(kb[1], kb[0], g.bits, GEN[g.name], slots[i][1], out1, 0)
                    grad[slots[i][0]] += float(out1.cpu()[0])
        if min(blk) <= first_param:
            break                           # nothing left to differentiate
        pending.extend(dagger(gates[i]) for i in blk)


def lower_circuit(circuit: Any, mode: str) -> Tuple[List[LGate], List[torch.Tensor]]:
    """Lower all gate ops; parametrised ops whose angle is a grad-requiring tensor get a slot."""
    n = int(getattr(circuit, "num_qubits", 0))
    ucache = getattr(circuit, "_unitary_cache", {}) or {}
    gates: List[LGate] = []
    thetas: List[torch.Tensor] = []
    for op in getattr(circuit, "ops", []):
        if not isinstance(op, (list, tuple)) or not op:
            continue
        nm = op[0]
        if nm in ("project_z", "reset", "kraus", "pulse", "pulse_inline"):
            raise NotImplementedError(f"gradients through {nm!r} are not supported")
        slot = None
        fixed = []
        for a in op:
            if isinstance(a, torch.Tensor):
                if a.requires_grad and slot is None:
                    slot = len(thetas)
                    thetas.append(a)
                fixed.append(float(a.detach().cpu()))
            else:
                fixed.append(a)
        g = lower_op(tuple(fixed), n, mode=mode, unitary_cache=ucache, param=slot)
        if g is None:
            if slot is not None:
                thetas.pop()
            continue
        if slot is not None and g.name not in GEN:
            raise NotImplementedError(f"no generator for parametrised op {g.name!r}")
        gates.append(g)
    return gates, thetas


def _gen_bits(g: LGate) -> Tuple[int, ...]:
    return g.bits


class _CircuitState(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, circuit, gates, *thetas):  # type: ignore[override]
        n = int(circuit.num_qubits)
        init = getattr(circuit, "_initial_state", None)
        if init is not None:
            arr = np.ascontiguousarray(_to_np(init).reshape(-1))
            state = torch.from_numpy(arr).to(engine.device).to(engine.dtype).contiguous()
        else:
            state = P.new_state(n, dtype=engine.dtype, device=engine.device)
        tile = engine.tile or default_tile(n, state.element_size(), 1)
        if gates:
            P.DeviceProgram(compile_program(gates, n, tile, itemsize=state.element_size()), state.device, state.dtype).run(state)
        ctx.engine = engine
        ctx.gates = gates
        ctx.n = n
        ctx.final = state
        ctx.theta_meta = [(t.dtype, t.device, t.shape) for t in thetas]
        out_dev = thetas[0].device if thetas else state.device
        ctx.out_cuda = out_dev.type == "cuda"
        return state.clone() if ctx.out_cuda else state.to(torch.complex128).cpu()

    @staticmethod
    def backward(ctx, grad_out):  # type: ignore[override]
        engine, gates, n = ctx.engine, ctx.gates, ctx.n
        dev, dtype = engine.device, engine.dtype
        kb = torch.empty((2, 1 << n), dtype=dtype, device=dev)
        kb[0].copy_(ctx.final)
        kb[1].copy_(grad_out.to(dev).to(dtype).reshape(-1))
        nparam = len(ctx.theta_meta)
        gout = torch.zeros(max(nparam, 1), dtype=torch.float64, device=dev)
        tile = engine.tile or default_tile(n, kb.element_size(), 2)
        pending: List[LGate] = []

        def flush() -> None:
            if pending:
                P.DeviceProgram(compile_program(pending, n, tile, itemsize=kb.element_size()), dev, dtype).run(kb)
                pending.clear()

        if n >= LAYERED_MIN_QUBITS:
            # layer by layer: all single-qubit gradients of a layer from one pair of states (layered_sweep)
            acc = np.zeros(max(nparam, 1))
            layered_sweep(kb, gates, [None if g.param is None else (g.param, 1.0) for g in gates], acc)
            gout += torch.from_numpy(acc).to(dev)
        else:
            for g in reversed(gates):
                if g.param is not None:
                    flush()
                    grad_dense(kb[1], kb[0], _gen_bits(g), GEN[g.name], 1.0, gout, g.param)
                pending.append(dagger(g))
        # the remaining un-applies are not needed for the gradient
        res = gout.cpu()
        grads = []
        for i, (dt, dv, shp) in enumerate(ctx.theta_meta):
            grads.append(res[i].to(dt).reshape(shp).to(dv))
        return (None, None, None, *grads)


def circuit_state_autograd(engine: Any, circuit: Any) -> torch.Tensor:
    gates, thetas = lower_circuit(circuit, "state")
    return _CircuitState.apply(engine, circuit, gates, *thetas)
