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
    LAYERED_MIN_QUBITS = 20   # from here on energy_and_grad takes the layer-by-layer sweep (energy_and_grad_layered)

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
        self._resident = None   # set at the end of __init__ (ResidentVQE is defined below)
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
        if mode == "state" and ResidentVQE.supports(self.n, self.template, dtype):
            self._resident = ResidentVQE(self.n, self.template, ham, device=self.device)

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

    # ---- layered adjoint sweep (large registers) ----------------------------------------------------------------
    def energy_and_grad_layered(self, params: Sequence[float]) -> Tuple[float, np.ndarray]:
        """Energy and adjoint gradient with the LAYERED reverse sweep (autograd.layered_sweep): every single-qubit gradient
        of a layer from one pair of states (tqb_transition_1q), fused un-apply passes between layers -- instead of one read of
        both states and one pass PER PARAMETER (civector_ops.py:141-200 gate by gate)."""
        from .autograd import layered_sweep
        from .fuse import fuse
        p = np.asarray(params, dtype=np.float64)
        shape = p.shape
        gates, refs = self._lower(p.reshape(-1))
        n = self.n
        kb = self._kb
        lib = _lib.load()
        grad = np.zeros(max(self.n_params, 1))
        with torch.cuda.device(self.device):
            ptr, _, _, dt, stream = P._prep(kb[0])
            _lib.check(lib.tqb_init_basis(ptr, n, 1, dt, 0, 0, stream))
            if gates:
                prog = compile_program(fuse(list(gates)), n, default_tile(n, self.itemsize, 1), itemsize=self.itemsize)
                P.DeviceProgram(prog, self.device, self.dtype).run(kb[0])
            self.ham.apply(kb[0], kb[1])
            e = float(P.inner(kb[0], kb[1])[0].real.cpu())
            slots = [None if r is None else (r.index, 2.0 * r.scale) for r in refs]
            layered_sweep(kb, gates, slots, grad)
        return e, grad[: self.n_params].reshape(shape).copy()

    def energy_and_grad_batch(self, params: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Many parameter vectors at once ([B, n_params] -> energies [B], gradients [B, n_params]): one launch of the
        CTA-resident kernel when the template qualifies, else one evaluation after the other."""
        p = np.asarray(params, dtype=np.float64).reshape(-1, max(self.n_params, 1))
        if self._resident is not None:
            return self._resident.energy_and_grad_batch(p)
        es, gs = zip(*(self.energy_and_grad(row) for row in p))
        return np.array(es), np.stack([np.asarray(g).reshape(-1) for g in gs])

    def energy_and_grad(self, params: Sequence[float], *, graph: bool = True, resident: Optional[bool] = None) -> Tuple[float, np.ndarray]:
        """``resident``: None = the CTA-resident kernel when the template qualifies (n <= 12, complex128, Pauli rotations +
        fixed Clifford-type gates), False = the fused-pass / CUDA-graph path below."""
        if resident is not False and self._resident is not None:
            return self._resident.energy_and_grad(params)
        if resident:
            raise NotImplementedError("template outside the resident kernel's op set")
        if self.n >= self.LAYERED_MIN_QUBITS and resident is None:
            return self.energy_and_grad_layered(params)
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


# ---- CTA-resident evaluation (csrc/tqb_vqe.cu) ---------------------------------------------------------------------
_PAULI_ROT = {"rx": "X", "ry": "Y", "rz": "Z", "rxx": "XX", "ryy": "YY", "rzz": "ZZ"}
_SQ = 2 ** -0.5
_FIXED_1Q = {
    "h": np.array([[_SQ, _SQ], [_SQ, -_SQ]], dtype=np.complex128),
    "x": np.array([[0, 1], [1, 0]], dtype=np.complex128),
    "s": np.array([[1, 0], [0, 1j]], dtype=np.complex128),
    "sdg": np.array([[1, 0], [0, -1j]], dtype=np.complex128),
}   # exactly the fixed 1-qubit ops the reference engine executes (engine.py:52-374; y / z / t are silently skipped there:
    # templates with other names keep the AdjointEnergy path, which mirrors that)
_FIXED_2Q = {   # matrix index = 2 * (first qubit) + (second qubit), as the reference's 4x4 gates (kernels/gates.py)
    "cx": np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128),
    "cz": np.diag([1, 1, 1, -1]).astype(np.complex128),
    "swap": np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128),
}


class ResidentVQE:
    """Energy + adjoint gradient of a template circuit for MANY parameter vectors in one launch: one shared-memory-resident
    CTA per vector walks forward circuit, H|psi>, <psi|H|psi> and the reverse sweep (csrc/tqb_vqe.cu).  n <= 12 qubits,
    complex128; ops: rx ry rz rxx ryy rzz with ``Param`` or fixed angles, h x s sdg cx cz swap.  Same numbers as
    ``AdjointEnergy`` (1e-10), which routes here when the template qualifies."""

    MAX_QUBITS = 12

    @classmethod
    def supports(cls, n: int, template: Sequence[Sequence[Any]], dtype: torch.dtype = torch.complex128) -> bool:
        if n > cls.MAX_QUBITS or n < 1 or dtype != torch.complex128:
            return False
        for op in template:
            nm = op[0]
            if nm == "measure_z":
                continue
            if nm in _PAULI_ROT:
                continue
            if nm in _FIXED_1Q or nm in _FIXED_2Q:
                if any(isinstance(a, Param) for a in op):
                    return False
                continue
            return False
        return True

    def __init__(self, n: int, template: Sequence[Sequence[Any]], ham: PauliSum, *, device: str | torch.device = "cuda") -> None:
        if not self.supports(n, template):
            raise NotImplementedError("template outside the resident kernel's op set (or n > 12)")
        self.n = int(n)
        self.device = torch.device(device)
        _lib.ensure_device(self.device.index or 0)
        self.n_params = 1 + max([a.index for op in template for a in op if isinstance(a, Param)], default=-1)
        ops = []
        mats: List[np.ndarray] = []
        moff = 0
        bit = lambda q: self.n - 1 - int(q)   # noqa: E731  (qubit q = index bit n-1-q)
        for op in template:
            nm = op[0]
            if nm == "measure_z":
                continue
            rec = np.zeros((), dtype=_lib.VQE_OP_DTYPE)
            if nm in _PAULI_ROT:
                ps = _PAULI_ROT[nm]
                x = z = 0
                for c, q in zip(ps, op[1:1 + len(ps)]):
                    if c in "XY":
                        x |= 1 << bit(q)
                    if c in "YZ":
                        z |= 1 << bit(q)
                a = op[1 + len(ps)]
                rec["kind"], rec["xmask"], rec["zmask"] = 0, x, z
                if isinstance(a, Param):
                    rec["param"], rec["scale"] = a.index, a.scale
                else:
                    rec["param"], rec["scale"] = -1, float(a)
            elif nm in _FIXED_1Q:
                rec["kind"], rec["param"], rec["bit0"], rec["mat_off"] = 1, -1, bit(op[1]), moff
                mats.append(_FIXED_1Q[nm].reshape(-1))
                moff += 4
            else:
                if int(op[1]) == int(op[2]):
                    continue   # apply_2q_statevector returns its input for q0 == q1 (statevector.py:46-47)
                rec["kind"], rec["param"], rec["bit0"], rec["bit1"], rec["mat_off"] = 2, -1, bit(op[1]), bit(op[2]), moff
                mats.append(_FIXED_2Q[nm].reshape(-1))
                moff += 16
            ops.append(rec)
        self._ops = torch.from_numpy(np.array(ops, dtype=_lib.VQE_OP_DTYPE).view(np.uint8).reshape(-1).copy()).to(self.device) if ops else \
            torch.zeros(40, dtype=torch.uint8, device=self.device)
        self._n_ops = len(ops)
        m = np.concatenate(mats) if mats else np.zeros(1, dtype=np.complex128)
        self._mats = torch.from_numpy(m.view(np.float64).copy()).to(self.device)
        self._hx = torch.from_numpy(ham.group_x.astype(np.uint32).view(np.int32).copy()).to(self.device)
        self._hp = torch.from_numpy(ham.group_ptr.astype(np.int32)).to(self.device)
        self._hz = torch.from_numpy(ham.term_z.astype(np.uint32).view(np.int32).copy()).to(self.device)
        self._hc = torch.from_numpy(ham.term_coef.view(np.float64).copy()).to(self.device)
        self._ng = ham.n_groups
        self._host = None

    def energy_and_grad_batch(self, params: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """params [B, n_params] -> (energies [B], gradients [B, n_params]); one kernel launch."""
        p = np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(-1, max(self.n_params, 1)))
        B = p.shape[0]
        if self._host is None or self._host[0].shape[0] < B:
            self._host = (torch.empty((B, max(self.n_params, 1)), dtype=torch.float64).pin_memory(),
                          torch.empty((B, 1 + self.n_params), dtype=torch.float64).pin_memory(),
                          torch.empty((B, max(self.n_params, 1)), dtype=torch.float64, device=self.device),
                          torch.empty((B, 1 + self.n_params), dtype=torch.float64, device=self.device))
        hp, ho, dp, do = self._host
        hp.numpy()[:B] = p
        with torch.cuda.device(self.device):
            # few vectors: the kernel reads the parameters from and writes the result to PINNED HOST memory directly
            # (unified addressing; a few hundred bytes over PCIe) -- no copy calls on the latency path of an optimiser loop
            zero_copy = B <= 64
            if not zero_copy:
                dp[:B].copy_(hp[:B], non_blocking=True)
            _lib.check(_lib.load().tqb_vqe_resident(self.n, self._ops.data_ptr(), self._n_ops, self._mats.data_ptr(), self._hx.data_ptr(),
                                                    self._hp.data_ptr(), self._ng, self._hz.data_ptr(), self._hc.data_ptr(),
                                                    (hp if zero_copy else dp).data_ptr(), self.n_params, B,
                                                    (ho if zero_copy else do).data_ptr(), _lib.current_stream_ptr(self.device)))
            if not zero_copy:
                ho[:B].copy_(do[:B], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        out = ho[:B].numpy()
        return out[:, 0].copy(), out[:, 1:1 + self.n_params].copy()

    def energy_and_grad(self, params: Sequence[float]) -> Tuple[float, np.ndarray]:
        p = np.asarray(params, dtype=np.float64)
        e, g = self.energy_and_grad_batch(p.reshape(1, -1))
        return float(e[0]), g[0].reshape(p.shape)


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
