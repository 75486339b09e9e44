"""Shot-based Pauli-sum energies: all measurement groups (and all parameter-shift variants) in ONE batched launch.

The reference's shots > 0 VQE path (applications/chem/runtimes/hea_device_runtime.py:120-262,
ucc_device_runtime.py:207-396) builds one circuit per measurement group -- the whole ansatz again plus a basis
rotation -- and for a gradient repeats that for 2P shifted parameter vectors: (1 + 2P) * G full simulations, each
followed by a Python loop over the counts dict (postprocessing/counts_expval.py:7-20).  Here:

    states[V, 2^n]  = the ansatz for the V = 1 + 2P parameter vectors       (one batched program)
    work[V*G, 2^n]  = every state copied once per group                      (member b = v*G + g)
    basis rotation  = one batched program, matrix of member b on wire q = I / H / H.Sdg by bases[g][q]
    idx[V*G, shots] = tqb_sample with host uniforms (same blocked-CDF contract as StatevectorEngine.run)
    energy[V*G]     = tqb_expval_from_samples: integer parity counts per term, coefficients in term order

so the ansatz is simulated once per parameter vector instead of once per group, and nothing but V*G doubles comes
back.  Grouping follows libs/hamiltonian_encoding/hamiltonian_grouping.py:120-138 (same dict order = same circuit
order).  ``bases[q]`` and term qubits are WIRE indices (bitstring position q = index bit n-1-q).
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import program as P
from .fuse import fuse
from .gates import C128, DENSE, DIAG, LGate, dense_gate, lower_op
from .planner import TileConfig, compile_program, default_tile
from .vqe import Param

Term = Tuple[Tuple[int, str], ...]
Groups = Dict[Tuple[str, ...], List[Tuple[Term, float]]]

_I = np.eye(2, dtype=C128)
_H = np.array([[1, 1], [1, -1]], dtype=C128) / math.sqrt(2.0)
_SDG = np.diag([1.0, -1.0j]).astype(C128)
_RZM = np.diag([np.exp(0.25j * math.pi), np.exp(-0.25j * math.pi)]).astype(C128)  # rz(-pi/2), gates.py:19-35


def group_pauli_terms(hamiltonian: Sequence[Tuple[float, Sequence[Tuple[str, int]]]], n: int) -> Tuple[float, Groups]:
    """[(coeff, [(P, q), ...]), ...] -> (identity_const, {bases: [(term, coeff), ...]}) -- hamiltonian_grouping.py:120-138."""
    identity = 0.0
    groups: Groups = {}
    for coeff, ops in hamiltonian:
        if not ops:
            identity += float(coeff)
            continue
        bases = ["I"] * n
        term = tuple((int(q), str(p).upper()) for (p, q) in ops)
        for q, p in term:
            bases[q] = p
        groups.setdefault(tuple(bases), []).append((term, float(coeff)))
    return identity, groups


def lower_batched(template: Sequence[tuple], params: np.ndarray, n: int, mode: str = "run") -> List[LGate]:
    """Lower an op template for B parameter vectors at once: ops that carry a ``Param`` become gates with one matrix
    per batch member (same kind and bits for every member), the others are shared."""
    params = np.asarray(params, dtype=np.float64)
    B = params.shape[0]
    out: List[LGate] = []
    for op in template:
        if not any(isinstance(a, Param) for a in op):
            g = lower_op(tuple(op), n, mode=mode)
            if g is not None:
                out.append(g)
            continue
        gs = [lower_op(tuple(a.scale * float(params[b, a.index]) if isinstance(a, Param) else a for a in op), n, mode=mode)
              for b in range(B)]
        g0 = gs[0]
        if g0 is None:
            continue
        if g0.kind not in (DENSE, DIAG):
            raise NotImplementedError(f"batched parameters on op {op[0]!r} (kind {g0.kind}) are not supported")
        data = np.stack([np.asarray(g.data, dtype=C128).reshape(-1) for g in gs])
        out.append(LGate(g0.kind, g0.bits, data, batched=True, name=g0.name))
    return out


class GroupedMeasurement:
    def __init__(self, n: int, groups: Groups, identity_const: float = 0.0, *, y_rotation: str = "sdg_h",
                 device: str | torch.device = "cuda", dtype: torch.dtype = torch.complex128,
                 tile: Optional[TileConfig] = None) -> None:
        if y_rotation not in ("sdg_h", "rz_h"):
            raise ValueError("y_rotation is 'sdg_h' (hea_device_runtime.py:286-289) or 'rz_h' (ucc_device_runtime.py:84-87)")
        self.n = int(n)
        self.identity = float(identity_const)
        self.bases = [tuple(b) for b in groups.keys()]
        self.items = [list(v) for v in groups.values()]
        self.G = len(self.bases)
        if self.G == 0:
            raise ValueError("no measurement groups")
        self.device = torch.device(device)
        self.dtype = dtype
        self.itemsize = 16 if dtype == torch.complex128 else 8
        self.tile = tile
        _lib.ensure_device(self.device.index or 0)
        rot_y = _H @ (_SDG if y_rotation == "sdg_h" else _RZM)
        self._rot = {"I": _I, "Z": _I, "X": _H, "Y": rot_y}
        ptr, zs, cs = [0], [], []
        for items in self.items:
            for term, coeff in items:
                z = 0
                for q, _p in term:
                    z |= 1 << (self.n - 1 - int(q))
                zs.append(z)
                cs.append(float(coeff))
            ptr.append(len(zs))
        self.max_terms = max(b - a for a, b in zip(ptr[:-1], ptr[1:]))
        self._term_ptr = torch.tensor(ptr, dtype=torch.int32, device=self.device)
        self._term_z = torch.from_numpy(np.array(zs, dtype=np.uint64).view(np.int64)).to(self.device)
        self._term_c = torch.tensor(cs, dtype=torch.float64, device=self.device)
        self._programs: Dict[int, P.DeviceProgram] = {}
        self.launches = 0

    @classmethod
    def from_pauli_list(cls, n: int, hamiltonian: Sequence[Tuple[float, Sequence[Tuple[str, int]]]], **kw: Any) -> "GroupedMeasurement":
        identity, groups = group_pauli_terms(hamiltonian, n)
        return cls(n, groups, identity, **kw)

    # ------------------------------------------------------------------------------------------
    def _rotation_program(self, V: int) -> Optional[P.DeviceProgram]:
        if V in self._programs:
            return self._programs[V]
        B = V * self.G
        gates: List[LGate] = []
        for q in range(self.n):
            letters = [b[q] for b in self.bases]
            if all(l in ("I", "Z") for l in letters):
                continue
            per_group = np.stack([self._rot[l] for l in letters])          # [G, 2, 2]
            gates.append(dense_gate(np.tile(per_group, (V, 1, 1)), [q], self.n, name="basis"))
        prog = None
        if gates:
            tile = self.tile or default_tile(self.n, self.itemsize, B)
            prog = P.DeviceProgram(compile_program(fuse(gates), self.n, tile, batch_mats=B, itemsize=self.itemsize), self.device, self.dtype)
        self._programs[V] = prog
        return prog

    def group_energies(self, states: torch.Tensor, uniforms: Any, *, want_expvals: bool = False) -> Any:
        """states: [V, 2^n] (or [2^n]) on the device, left untouched.  uniforms: [V*G, shots] float64 (host or device),
        row v*G + g drives the sampler of variant v in group g.  Returns float64 [V, G] (device) of
        sum_t coeff_t <Z_S_t> per group, optionally also the per-term expectation values [V, G, max_terms]."""
        st = states.reshape(-1, 1 << self.n)
        V = int(st.shape[0])
        B = V * self.G
        u = torch.as_tensor(uniforms, dtype=torch.float64).reshape(B, -1).to(self.device)
        shots = int(u.shape[1])
        work = st.to(self.dtype).repeat_interleave(self.G, dim=0).contiguous()
        prog = self._rotation_program(V)
        if prog is not None:
            prog.run(work)
        idx = P.sample(work, u)
        energy = torch.empty(B, dtype=torch.float64, device=self.device)
        evs = torch.zeros((B, self.max_terms), dtype=torch.float64, device=self.device) if want_expvals else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().tqb_expval_from_samples(
                idx.data_ptr(), B, shots, self.G, self._term_ptr.data_ptr(), self._term_z.data_ptr(), self._term_c.data_ptr(),
                energy.data_ptr(), evs.data_ptr() if evs is not None else None, self.max_terms, _lib.current_stream_ptr(self.device)))
        energy = energy.reshape(V, self.G)
        return (energy, evs.reshape(V, self.G, self.max_terms)) if want_expvals else energy

    def energies(self, states: torch.Tensor, uniforms: Any) -> np.ndarray:
        """identity + sum over groups (in group order, like hea_device_runtime.py:173-177), one value per state."""
        e = self.group_energies(states, uniforms).cpu().numpy()
        out = np.empty(e.shape[0], dtype=np.float64)
        for v in range(e.shape[0]):
            acc = self.identity
            for g in range(self.G):
                acc += float(e[v, g])
            out[v] = acc
        return out


class ShotEnergy:
    """E(theta) and its parameter-shift gradient from shots (hea_device_runtime.py:120-262) for an op template with
    ``Param`` placeholders: g_i = (E[theta_i + pi/2] - E[theta_i - pi/2]) / 2."""

    def __init__(self, n: int, template: Sequence[tuple], measurement: GroupedMeasurement, *, mode: str = "run",
                 tile: Optional[TileConfig] = None) -> None:
        self.n = int(n)
        self.template = [tuple(op) for op in template]
        self.meas = measurement
        self.mode = mode
        self.tile = tile
        self.n_params = 1 + max([a.index for op in self.template for a in op if isinstance(a, Param)], default=-1)
        self.passes = 0

    def _check_shift_rule(self) -> None:
        """The reference's rule (hea_device_runtime.py:180-262) shifts the PARAMETER by +-pi/2: exact only when every
        parameter drives exactly one rotation gate exp(-i theta/2 P) with unit scale -- anything else is refused instead
        of returning a silently wrong gradient (AdjointEnergy / ShardedStatevectorEngine.energy_and_grad differentiate
        per gate occurrence and take such templates)."""
        seen: dict = {}
        for op in self.template:
            for a in op:
                if isinstance(a, Param):
                    if op[0] not in ("rx", "ry", "rz", "rxx", "ryy", "rzz"):
                        raise NotImplementedError(f"parameter shift: no two-term rule for parametrised op {op[0]!r}")
                    if a.scale != 1.0:
                        raise NotImplementedError("parameter shift on the parameter needs Param.scale == 1 (scaled angles: use AdjointEnergy)")
                    if a.index in seen:
                        raise NotImplementedError("parameter shift on the parameter needs every parameter in exactly one gate (shared parameters: use AdjointEnergy)")
                    seen[a.index] = True

    def states(self, params: np.ndarray) -> torch.Tensor:
        params = np.asarray(params, dtype=np.float64).reshape(-1, max(self.n_params, 1))
        B = params.shape[0]
        m = self.meas
        st = P.new_state(self.n, batch=B, dtype=m.dtype, device=m.device)
        gates = fuse(lower_batched(self.template, params, self.n, self.mode))
        if gates:
            tile = self.tile or default_tile(self.n, m.itemsize, B)
            prog = compile_program(gates, self.n, tile, batch_mats=B, itemsize=m.itemsize)
            P.DeviceProgram(prog, m.device, m.dtype).run(st)
            self.passes = prog.n_passes
        return st

    def energy(self, params: Sequence[float], uniforms: Any) -> float:
        return float(self.meas.energies(self.states(np.asarray(params)[None, :]), uniforms)[0])

    def energy_and_grad(self, params: Sequence[float], uniforms: Any) -> Tuple[float, np.ndarray]:
        """uniforms: [(1 + 2P) * G, shots], circuit order base, (+0, -0), (+1, -1), ... each over all groups."""
        self._check_shift_rule()
        base = np.asarray(params, dtype=np.float64).reshape(-1)
        Pn = base.size
        variants = np.tile(base, (1 + 2 * Pn, 1))
        for i in range(Pn):
            variants[1 + 2 * i, i] += 0.5 * math.pi
            variants[2 + 2 * i, i] -= 0.5 * math.pi
        e = self.meas.group_energies(self.states(variants), uniforms).cpu().numpy()
        tot = np.empty(e.shape[0], dtype=np.float64)
        for v in range(e.shape[0]):
            acc = self.meas.identity if v == 0 else 0.0   # the reference adds the constant to the base energy only
            for g in range(self.meas.G):
                acc += float(e[v, g])
            tot[v] = acc
        grad = np.array([0.5 * (tot[1 + 2 * i] - tot[2 + 2 * i]) for i in range(Pn)])
        return float(tot[0]), grad
