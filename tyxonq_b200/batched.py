"""Batched parameter sets: B states of n qubits as one [B, 2^n] device array.

Workload of examples/vqe_batched_architecture_search.py (config 5 of BASELINE.json): the hardware-
efficient RY ansatz of libs/circuits_library/blocks.py:60-85 for B parameter sets at once, then a
matrix-free Pauli-sum expectation per state (instead of the dense 2^n x 2^n Hamiltonian of
kernels/pauli.py:74-87) and shots per state from host-supplied uniforms.  The batch index is just more
tile-index bits for the pass kernel; gates carry one matrix per batch member (tqb_gate.mat_bstride).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import _lib
from . import program as P
from .fuse import fuse
from .gates import LGate, dense_gate, swap_gate
from .pauli import PauliSum
from .planner import TileConfig, compile_program, default_tile


def hwe_ry_gates(n: int, layers: int, params: np.ndarray) -> List[LGate]:
    """build_hwe_ry_ops (blocks.py:60-85) for params [B, (layers+1)*n]: ry layers carry B matrices each."""
    p = np.asarray(params, dtype=np.float64).reshape(params.shape[0], layers + 1, n)
    c, s = np.cos(0.5 * p), np.sin(0.5 * p)
    ry = np.stack([c, -s, s, c], axis=-1).astype(np.complex128)  # [B, layer, qubit, 4]
    gates: List[LGate] = [dense_gate(ry[:, 0, q], [q], n, name="ry") for q in range(n)]
    for l in range(layers):
        gates += [swap_gate([q, q + 1], n, 0b10, 0b11, name="cx") for q in range(n - 1)]
        gates += [dense_gate(ry[:, l + 1, q], [q], n, name="ry") for q in range(n)]
    return gates


class BatchedAnsatz:
    def __init__(self, n: int, layers: int, batch: int, *, device: str | torch.device = "cuda",
                 dtype: torch.dtype = torch.complex64, tile: Optional[TileConfig] = None) -> None:
        self.n, self.layers, self.batch = int(n), int(layers), int(batch)
        self.device = torch.device(device)
        self.dtype = dtype
        self.itemsize = 16 if dtype == torch.complex128 else 8
        self.tile = tile or default_tile(self.n, self.itemsize, self.batch)
        _lib.ensure_device(self.device.index or 0)
        self.state = torch.empty((self.batch, 1 << self.n), dtype=dtype, device=self.device)
        self.h2d_bytes = 0
        self.passes = 0

    def run(self, params: np.ndarray) -> torch.Tensor:
        """|psi_b> = ansatz(params[b]) |0..0> for every batch member, in place on the device."""
        gates = fuse(hwe_ry_gates(self.n, self.layers, np.asarray(params)))
        prog = compile_program(gates, self.n, self.tile, batch_mats=self.batch, itemsize=self.itemsize)
        dp = P.DeviceProgram(prog, self.device, self.dtype)
        ptr, n, b, dt, stream = P._prep(self.state)
        _lib.check(_lib.load().tqb_init_basis(ptr, n, b, dt, 0, 0, stream))
        dp.run(self.state)
        self.h2d_bytes = dp.h2d_bytes
        self.passes = prog.n_passes
        return self.state

    def expvals(self, ham: PauliSum) -> torch.Tensor:
        return ham.expectation(self.state).real

    def sample(self, uniforms: torch.Tensor) -> torch.Tensor:
        return P.sample(self.state, uniforms)
