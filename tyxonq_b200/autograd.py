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
