"""Variational energies with adjoint gradients on the device.

``AdjointEnergy`` evaluates E(theta) = <psi(theta)| H |psi(theta)> for a circuit template and a
Pauli-sum Hamiltonian, and its gradient by one adjoint sweep (2 state buffers, no tape):
it replaces ``value_and_grad`` of the reference's numerics backends on this path
(numerics/backends/numpy_backend.py:386-454 = P+1 finite-difference evaluations,
pytorch_backend.py:446-564 = autograd tape of einsum nodes).

``TFIMVqe`` is the workload of examples/vqetfim_benchmark.py (ansatz :21-34, energy :70-103).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import program as P
from .autograd import dagger, grad_dense
from .gates import GEN, LGate, lower_op
from .pauli import PauliSum
from .planner import TileConfig, compile_program, default_tile


@dataclass(frozen=True)
class Param:
    """Placeholder for an angle: scale * params[index]."""
    index: int
    scale: float = 1.0


class AdjointEnergy:
    def __init__(self, n: int, template: Sequence[Sequence[Any]], ham: PauliSum, *, device: str | torch.device = "cuda",
                 dtype: torch.dtype = torch.complex128, tile: Optional[TileConfig] = None, mode: str = "state") -> None:
        self.n = int(n)
        self.template = [tuple(op) for op in template]
        self.ham = ham
        self.device = torch.device(device)
        self.dtype = dtype
        self.mode = mode
        self.tile = tile or default_tile(self.n, 16 if dtype == torch.complex128 else 8, 2)
        self.n_params = 1 + max([a.index for op in self.template for a in op if isinstance(a, Param)], default=-1)
        _lib.ensure_device(self.device.index or 0)
        self._kb = torch.empty((2, 1 << self.n), dtype=dtype, device=self.device)
        self._gout = torch.zeros(max(self.n_params, 1), dtype=torch.float64, device=self.device)

    def _lower(self, params: np.ndarray) -> Tuple[List[LGate], List[Optional[Param]]]:
        gates: List[LGate] = []
        refs: List[Optional[Param]] = []
        for op in self.template:
            ref = next((a for a in op if isinstance(a, Param)), None)
            fixed = tuple(a.scale * float(params[a.index]) if isinstance(a, Param) else a for a in op)
            g = lower_op(fixed, self.n, mode=self.mode, param=(len(gates) if ref is not None else None))
            if g is None:
                continue
            if ref is not None and g.name not in GEN:
                raise NotImplementedError(f"no generator for parametrised op {g.name!r}")
            gates.append(g)
            refs.append(ref)
        return gates, refs

    def statevector(self, params: Sequence[float]) -> torch.Tensor:
        gates, _ = self._lower(np.asarray(params, dtype=np.float64).reshape(-1))
        st = P.new_state(self.n, dtype=self.dtype, device=self.device)
        P.apply_gates(st, gates, tile=self.tile)
        return st

    def energy(self, params: Sequence[float]) -> float:
        return float(self.ham.expectation(self.statevector(params))[0].real.cpu())

    def energy_and_grad(self, params: Sequence[float]) -> Tuple[float, np.ndarray]:
        p = np.asarray(params, dtype=np.float64)
        shape = p.shape
        p = p.reshape(-1)
        gates, refs = self._lower(p)
        kb = self._kb
        with torch.cuda.device(self.device):
            ptr, n, _, dt, stream = P._prep(kb[0])
            _lib.check(_lib.load().tqb_init_basis(ptr, n, 1, dt, 0, 0, stream))
            P.apply_gates(kb[0], gates, tile=self.tile)
            self.ham.apply(kb[0], kb[1])
            e = P.inner(kb[0], kb[1])
            self._gout.zero_()
            pending: List[LGate] = []
            for g, ref in zip(reversed(gates), reversed(refs)):
                if ref is not None:
                    if pending:
                        P.apply_gates(kb, pending, tile=self.tile)
                        pending = []
                    grad_dense(kb[1], kb[0], g.bits, GEN[g.name], 2.0 * ref.scale, self._gout, ref.index)
                pending.append(dagger(g))
            out = torch.cat([torch.view_as_real(e).reshape(-1)[:1], self._gout[: max(self.n_params, 1)]]).cpu().numpy()
        return float(out[0]), out[1:1 + self.n_params].reshape(shape).copy()


def tfim_hamiltonian(n: int, Jx: float = 1.0, h: float = -1.0) -> PauliSum:
    """H = h * sum Z_i + Jx * sum X_i X_{i+1}  (examples/vqetfim_benchmark.py:88-103)."""
    ham = [(h, [("Z", i)]) for i in range(n)] + [(Jx, [("X", i), ("X", i + 1)]) for i in range(n - 1)]
    return PauliSum.from_pauli_list(n, ham)


def tfim_template(n: int, nlayers: int) -> List[tuple]:
    """ansatz_ops_xx_rz (examples/vqetfim_benchmark.py:21-34); param layout [2*nlayers, n] flattened."""
    ops: List[tuple] = []
    t = 0
    for _ in range(nlayers):
        ops += [("rxx", i, i + 1, Param(t * n + i)) for i in range(n - 1)]
        t += 1
        ops += [("rz", i, Param(t * n + i)) for i in range(n)]
        t += 1
    return ops


class TFIMVqe(AdjointEnergy):
    def __init__(self, n: int = 10, nlayers: int = 1, *, Jx: float = 1.0, h: float = -1.0, **kw: Any) -> None:
        super().__init__(n, tfim_template(n, nlayers), tfim_hamiltonian(n, Jx, h), **kw)
        self.nlayers = nlayers
