"""Variational energies with adjoint gradients on the device.

``AdjointEnergy`` evaluates E(theta) = <psi(theta)| H |psi(theta)> for a circuit template and a
Pauli-sum Hamiltonian, and its gradient by one adjoint sweep (2 state buffers, no tape):
it replaces ``value_and_grad`` of the reference's numerics backends on this path
(numerics/backends/numpy_backend.py:386-454 = P+1 finite-difference evaluations,
pytorch_backend.py:446-564 = autograd tape of einsum nodes).

The circuit structure is compiled once (forward program + one un-apply pass per gate); an
evaluation only rewrites the matrices in a pinned buffer and replays one CUDA graph.

``TFIMVqe`` is the workload of examples/vqetfim_benchmark.py (ansatz :21-34, energy :70-103).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import program as P
from .autograd import dagger
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
        self.itemsize = 16 if dtype == torch.complex128 else 8
        self.tile = tile or default_tile(self.n, self.itemsize, 2)
        self.n_params = 1 + max([a.index for op in self.template for a in op if isinstance(a, Param)], default=-1)
        _lib.ensure_device(self.device.index or 0)
        self._kb = torch.empty((2, 1 << self.n), dtype=dtype, device=self.device)
        self._gout = torch.zeros(max(self.n_params, 1), dtype=torch.float64, device=self.device)
        self._e = torch.zeros(2, dtype=torch.float64, device=self.device)
        self._out = torch.zeros(1 + max(self.n_params, 1), dtype=torch.float64, device=self.device)
        self._out_host = torch.zeros(1 + max(self.n_params, 1), dtype=torch.float64).pin_memory()
        self._graph = None
        self._calls = 0
        # structure: lower once (theta = 0), compile the forward program and one un-apply pass per gate
        gates, self._refs = self._lower(np.zeros(max(self.n_params, 1)))
        self._ng = len(gates)
        self._fwd = compile_program(gates, self.n, self.tile, itemsize=self.itemsize)
        one = TileConfig(m=self.tile.m, L=self.tile.L, threads=self.tile.threads, ctas_per_sm=self.tile.ctas_per_sm, max_gates=1)
        self._rev = compile_program([dagger(g) for g in reversed(gates)], self.n, one, itemsize=self.itemsize)
        assert self._rev.order == list(range(self._ng))
        self._fwd_dev = P.DeviceProgram(self._fwd, self.device, dtype)
        self._rev_dev = P.DeviceProgram(self._rev, self.device, dtype)
        self._gen = []  # per gate: None or (bits array, generator as float64 pairs, scale, slot)
        for g, ref in zip(gates, self._refs):
            if ref is None:
                self._gen.append(None)
                continue
            bits = (C.c_int * len(g.bits))(*[int(b) for b in g.bits])
            gen = np.ascontiguousarray(np.asarray(GEN[g.name], dtype=np.complex128).reshape(-1)).view(np.float64).copy()
            self._gen.append((bits, gen, 2.0 * ref.scale, ref.index, len(g.bits)))

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

    def _fill(self, params: np.ndarray) -> None:
        gates, _ = self._lower(params)
        for prog, dev, lst in ((self._fwd, self._fwd_dev, gates), (self._rev, self._rev_dev, [dagger(g) for g in reversed(gates)])):
            mats = np.empty(prog.mats.size, dtype=np.complex128)
            off = prog.gates["mat_off"]
            for i, gi in enumerate(prog.order):
                d = np.asarray(lst[gi].data, dtype=np.complex128).reshape(-1)
                mats[off[i]:off[i] + d.size] = d
            dev.fill_host(mats)

    def statevector(self, params: Sequence[float]) -> torch.Tensor:
        p = np.asarray(params, dtype=np.float64).reshape(-1)
        self._fill(p)
        st = P.new_state(self.n, dtype=self.dtype, device=self.device)
        self._fwd_dev.upload()
        self._fwd_dev.run(st)
        return st

    def energy(self, params: Sequence[float]) -> float:
        return float(self.ham.expectation(self.statevector(params))[0].real.cpu())

    def _enqueue(self) -> None:
        lib = _lib.load()
        kb = self._kb
        ptr, n, _, dt, stream = P._prep(kb[0])
        _lib.check(lib.tqb_init_basis(ptr, n, 1, dt, 0, 0, stream))
        self._fwd_dev.upload()
        self._rev_dev.upload()
        self._fwd_dev.run(kb[0])
        self.ham.apply(kb[0], kb[1])
        _lib.check(lib.tqb_inner(kb[0].data_ptr(), kb[1].data_ptr(), n, 1, dt, self._e.data_ptr(), stream))
        self._gout.zero_()
        t = self.tile
        passes = self._rev_dev._passes
        psz = passes.dtype.itemsize
        last_param = max([i for i, g in enumerate(self._gen) if g is not None], default=-1)
        for i in range(self._ng):          # i-th un-applied gate = gate ng-1-i of the circuit
            j = self._ng - 1 - i
            if self._gen[j] is not None:
                bits, gen, scale, slot, k = self._gen[j]
                _lib.check(lib.tqb_grad_dense(kb[1].data_ptr(), kb[0].data_ptr(), n, dt, k, C.cast(bits, C.c_void_p),
                                              gen.ctypes.data, float(scale), self._gout.data_ptr(), int(slot), stream))
            if j == 0 or all(g is None for g in self._gen[:j]):
                break                      # nothing left to differentiate: skip the remaining un-applies
            _lib.check(lib.tqb_run_passes(kb.data_ptr(), n, 2, dt, 0, passes.ctypes.data + i * psz, 1,
                                          self._rev_dev.gates_dev.data_ptr(), self._rev_dev.mats_dev.data_ptr(),
                                          t.threads, t.ctas_per_sm, stream))
        self._out[0:1].copy_(self._e[0:1])
        self._out[1:1 + self.n_params].copy_(self._gout[: self.n_params])

    def energy_and_grad(self, params: Sequence[float], *, graph: bool = True) -> Tuple[float, np.ndarray]:
        p = np.asarray(params, dtype=np.float64)
        shape = p.shape
        self._fill(p.reshape(-1))
        with torch.cuda.device(self.device):
            if graph and self._graph is None and self._calls >= 1:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue()
                self._graph = g
            if graph and self._graph is not None:
                self._graph.replay()
            else:
                self._enqueue()
            self._calls += 1
            self._out_host.copy_(self._out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        out = self._out_host.numpy()
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
