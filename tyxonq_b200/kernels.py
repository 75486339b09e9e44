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


def init_statevector(num_qubits: int, backend: Any | None = None, *, device: Any = None, dtype: torch.dtype = torch.complex128) -> torch.Tensor:
    """statevector.py:19-25.  Returns a CUDA tensor (the reference builds a Python list of 2^n complex)."""
    return P.new_state(max(int(num_qubits), 0), dtype=dtype, device=device or _device())


def apply_1q_statevector(backend: Any, state: Any, gate2: Any, qubit: int, num_qubits: int) -> Any:
    """statevector.py:28-42."""
    t, how = _to_dev(state)
    P.apply_gates(t, [dense_gate(_to_np(gate2).reshape(2, 2), [int(qubit)], int(num_qubits))])
    return _back(t, how)


def apply_2q_statevector(backend: Any, state: Any, gate4: Any, q0: int, q1: int, num_qubits: int) -> Any:
    """statevector.py:45-59 (returns the input unchanged when q0 == q1)."""
    if q0 == q1:
        return state
    t, how = _to_dev(state)
    P.apply_gates(t, [classify_unitary(_to_np(gate4).reshape(4, 4), [int(q0), int(q1)], int(num_qubits))])
    return _back(t, how)


def apply_kqubit_unitary(state: Any, unitary: Any, qubit_indices: Sequence[int], num_qubits: int, backend: Any | None = None) -> Any:
    """statevector.py:71-129 (first listed qubit = most significant bit of the matrix index); k <= 4."""
    k = len(qubit_indices)
    if k == 0:
        return state
    if k > 4:
        raise NotImplementedError("apply_kqubit_unitary: k > 4 dense blocks are not supported on the device path")
    t, how = _to_dev(state)
    P.apply_gates(t, [classify_unitary(_to_np(unitary), [int(q) for q in qubit_indices], int(num_qubits))])
    return _back(t, how)


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
