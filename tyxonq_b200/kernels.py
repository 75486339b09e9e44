"""Kernel functions with the reference's signatures (seam B2 of SURVEY.md section 8b).

Same names, argument order and out-of-place semantics as
``tyxonq.libs.quantum_library.kernels.statevector`` (statevector.py:19-218); the work is
done by the CUDA kernels.  Inputs may be numpy arrays, CPU torch tensors or CUDA tensors;
the result comes back in the same container (CUDA tensors stay on the device).  The
``backend`` argument is accepted for signature compatibility and otherwise ignored.
"""
from __future__ import annotations

from typing import Any, Sequence

import numpy as np
import torch

from . import _lib
from . import program as P
from .gates import classify_unitary, dense_gate, _to_np

__all__ = [
    "init_statevector", "apply_1q_statevector", "apply_2q_statevector", "apply_kqubit_unitary",
    "expect_z_statevector", "apply_kraus_statevector",
]


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.TqbError("tyxonq_b200 kernels need a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(state: Any) -> tuple[torch.Tensor, str]:
    """Fresh device copy of the input state (kernels are out-of-place) + how to hand it back."""
    if isinstance(state, torch.Tensor):
        if state.is_cuda:
            t = state.detach().clone().contiguous().reshape(-1)
            if not t.is_complex():
                t = t.to(torch.complex128)
            return t, "cuda"
        return state.detach().to(torch.complex128).reshape(-1).to(_device()).contiguous(), "torch"
    arr = np.ascontiguousarray(np.asarray(state, dtype=np.complex128).reshape(-1))
    return torch.from_numpy(arr).to(_device()), "numpy"


def _back(t: torch.Tensor, how: str) -> Any:
    if how == "cuda":
        return t
    c = t.cpu()
    return c.numpy() if how == "numpy" else c


def _needs_grad(*xs: Any) -> bool:
    return torch.is_grad_enabled() and any(isinstance(x, torch.Tensor) and x.requires_grad for x in xs)


class _ApplyUnitary(torch.autograd.Function):
    """y = U_qubits x as one device pass, differentiable like the reference's einsum kernels (statevector.py:28-129):
    backward is grad_x = U^H grad_y (one more pass) and, when the gate itself requires grad, grad_U[a, b] =
    sum_rest grad_y[a, rest] conj(x[b, rest]) from the device reduction tqb_grad_dense (k <= 2)."""

    @staticmethod
    def forward(ctx, state: torch.Tensor, gate: torch.Tensor, qubits: tuple, n: int):  # type: ignore[override]
        t, how = _to_dev(state)
        U = _to_np(gate).reshape(1 << len(qubits), 1 << len(qubits))
        P.apply_gates(t, [classify_unitary(U, list(qubits), n)])
        ctx.qubits, ctx.n, ctx.U = tuple(qubits), int(n), U
        ctx.gate_meta = (gate.dtype, gate.device, gate.shape) if isinstance(gate, torch.Tensor) else None
        ctx.state_meta = (state.dtype, state.device, state.shape)
        ctx.need_gate = bool(isinstance(gate, torch.Tensor) and gate.requires_grad)
        if ctx.need_gate:
            ctx.save_for_backward(state)
        out = _back(t, how)
        return out.reshape(state.shape) if isinstance(out, torch.Tensor) else out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):  # type: ignore[override]
        qubits, n, U = ctx.qubits, ctx.n, ctx.U
        gy, _ = _to_dev(grad_out)
        g_state = None
        if ctx.needs_input_grad[0]:
            gx = gy.clone()
            P.apply_gates(gx, [classify_unitary(U.conj().T.copy(), list(qubits), n)])
            dt, dv, shp = ctx.state_meta
            g_state = gx.to(dv).to(dt).reshape(shp)
        g_gate = None
        if ctx.need_gate:
            k = len(qubits)
            if k > 2:
                raise NotImplementedError("gradient with respect to a k > 2 gate matrix is not supported on the device path")
            from .autograd import grad_dense
            (x,) = ctx.saved_tensors
            xd, _ = _to_dev(x)
            d = 1 << k
            out = torch.zeros(2 * d * d, dtype=torch.float64, device=gy.device)
            bits = [n - 1 - int(q) for q in reversed(qubits)]   # matrix-index bit j (LSB first) -> index bit
            for a in range(d):
                for b in range(d):
                    E = np.zeros((d, d), dtype=np.complex128)
                    E[b, a] = 1.0
                    grad_dense(xd, gy, bits, E, 1.0, out, 2 * (a * d + b))            # Re G[a, b]
                    grad_dense(xd, gy, bits, -1j * E, 1.0, out, 2 * (a * d + b) + 1)   # Im G[a, b]
            G = torch.view_as_complex(out.reshape(d, d, 2).contiguous())
            dt, dv, shp = ctx.gate_meta
            g_gate = G.to(dv).to(dt if dt.is_complex else torch.complex128).reshape(shp)
        return g_state, g_gate, None, None


def init_statevector(num_qubits: int, backend: Any | None = None, *, device: Any = None, dtype: torch.dtype = torch.complex128) -> torch.Tensor:
    """statevector.py:19-25.  Returns a CUDA tensor (the reference builds a Python list of 2^n complex)."""
    return P.new_state(max(int(num_qubits), 0), dtype=dtype, device=device or _device())


def apply_1q_statevector(backend: Any, state: Any, gate2: Any, qubit: int, num_qubits: int) -> Any:
    """statevector.py:28-42."""
    if _needs_grad(state, gate2):
        return _ApplyUnitary.apply(state, gate2 if isinstance(gate2, torch.Tensor) else torch.as_tensor(_to_np(gate2)), (int(qubit),), int(num_qubits))
    t, how = _to_dev(state)
    P.apply_gates(t, [dense_gate(_to_np(gate2).reshape(2, 2), [int(qubit)], int(num_qubits))])
    return _back(t, how)


def apply_2q_statevector(backend: Any, state: Any, gate4: Any, q0: int, q1: int, num_qubits: int) -> Any:
    """statevector.py:45-59 (returns the input unchanged when q0 == q1)."""
    if q0 == q1:
        return state
    if _needs_grad(state, gate4):
        return _ApplyUnitary.apply(state, gate4 if isinstance(gate4, torch.Tensor) else torch.as_tensor(_to_np(gate4)), (int(q0), int(q1)), int(num_qubits))
    t, how = _to_dev(state)
    P.apply_gates(t, [classify_unitary(_to_np(gate4).reshape(4, 4), [int(q0), int(q1)], int(num_qubits))])
    return _back(t, how)


def apply_kqubit_unitary(state: Any, unitary: Any, qubit_indices: Sequence[int], num_qubits: int, backend: Any | None = None) -> Any:
    """statevector.py:71-129 (first listed qubit = most significant bit of the matrix index).  k <= 4: one fused device
    pass; k > 4: one device GEMM (_apply_dense_large)."""
    k = len(qubit_indices)
    if k == 0:
        return state
    if k > 4:
        return _apply_dense_large(state, unitary, [int(q) for q in qubit_indices], int(num_qubits))
    if _needs_grad(state, unitary):
        return _ApplyUnitary.apply(state, unitary if isinstance(unitary, torch.Tensor) else torch.as_tensor(_to_np(unitary)),
                                   tuple(int(q) for q in qubit_indices), int(num_qubits))
    t, how = _to_dev(state)
    P.apply_gates(t, [classify_unitary(_to_np(unitary), [int(q) for q in qubit_indices], int(num_qubits))])
    return _back(t, how)


def _apply_dense_large(state: Any, unitary: Any, qubits: Sequence[int], n: int) -> Any:
    """k > 4 target qubits (the reference has no limit, statevector.py:71-129): the 2^k x 2^k block is GEMM-shaped, so it
    runs as ONE device matmul on the state viewed as [2^(n-k), 2^k] with the target axes moved last (a plain library
    GEMM through torch, differentiable by torch itself); the fused pass kernels stop at 4-qubit blocks."""
    k = len(qubits)
    if len(set(qubits)) != k or any(q < 0 or q >= n for q in qubits):
        raise ValueError("apply_kqubit_unitary: bad qubit indices")
    dev = _device()
    if isinstance(state, torch.Tensor):
        how = "cuda" if state.is_cuda else "torch"
        t = state.to(dev).to(torch.complex128).reshape(-1)
    else:
        how = "numpy"
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(state, dtype=np.complex128).reshape(-1))).to(dev)
    U = unitary if isinstance(unitary, torch.Tensor) else torch.from_numpy(_to_np(unitary))
    U = U.to(dev).to(torch.complex128).reshape(1 << k, 1 << k)
    rest = [q for q in range(n) if q not in qubits]
    perm = rest + list(qubits)                     # tensor axis q = qubit q (big-endian), first listed target = MSB of the block index
    psi = t.reshape([2] * n).permute(perm).reshape(1 << (n - k), 1 << k)
    out = (psi @ U.transpose(0, 1)).reshape([2] * n)
    inv = [0] * n
    for pos, q in enumerate(perm):
        inv[q] = pos
    out = out.permute(inv).reshape(-1)
    if how == "cuda":
        return out.reshape(state.shape)
    c = out.cpu()
    return c.numpy() if how == "numpy" else c.reshape(state.shape)


def expect_z_statevector(state: Any, qubit: int, num_qubits: int, backend: Any | None = None) -> Any:
    """statevector.py:62-68."""
    if isinstance(state, torch.Tensor) and state.is_cuda:
        t = state.detach().contiguous().reshape(-1)
        return P.expect_z_bits(t)[0, int(num_qubits) - 1 - int(qubit)]
    t, how = _to_dev(state)
    v = P.expect_z_bits(t)[0, int(num_qubits) - 1 - int(qubit)].cpu()
    return v if how == "torch" else np.float64(v.item())


def apply_kraus_statevector(state: Any, kraus_operators: Sequence[Any], qubit: int, num_qubits: int,
                            status: float | None = None, backend: Any | None = None) -> Any:
    """statevector.py:132-218."""
    from .engine import StatevectorEngine
    t, how = _to_dev(state)
    eng = StatevectorEngine("b200", device=t.device, dtype=t.dtype)
    eng._apply_kraus(t, [_to_np(k) for k in kraus_operators], int(qubit), int(num_qubits), status)
    return _back(t, how)
