"""Seam B1 over several GPUs: the ``run(circuit, shots)`` contract of the reference's statevector engine
(devices/simulators/statevector/engine.py:43-473) for states that do not fit one B200.

Launched under ``torchrun`` (one rank per GPU, ``torch.distributed`` initialised): EVERY rank constructs the engine and
calls ``run`` / ``expval`` with the same circuit -- the calls are collective -- and every rank gets the same result
dict.  The state is sharded by its g = log2(world) highest index bits (``sharded.ShardedState``); gates on global
qubits cost amplitude exchanges planned by ``sharded.plan_sharded``; counts come from the blocked-CDF sampler across
ranks and equal the single-GPU (and the reference-order) counts bit for bit given the same uniforms; <Z_q> and Pauli
sums are local reductions plus one all-reduce.  Ops outside the sharded path (kraus, project_z / reset, pulses,
noise models) raise: there is no fallback.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .gates import lower_op
from .sharded import ShardedState, plan_sharded

_NOT_SHARDED = ("kraus", "project_z", "reset", "pulse", "pulse_inline")


class ShardedStatevectorEngine:
    name = "statevector"
    capabilities = {"supports_shots": True, "sharded": True}

    def __init__(self, backend_name: str | None = None, *, device: str | torch.device | None = None,
                 dtype: torch.dtype = torch.complex128, group: Any = None, local_backend: Optional[Any] = None) -> None:
        """``local_backend``: the per-rank executor handed to ShardedState (default: this package's CUDA kernels)."""
        self.backend_name = backend_name or "b200"
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("ShardedStatevectorEngine needs a CUDA device (no CPU fallback)")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.dtype = dtype
        self.group = group
        self.local_backend = local_backend
        self.last_state: Optional[ShardedState] = None
        self.last_exchanges = 0

    @property
    def backend_label(self) -> str:
        return f"{self.backend_name}-sharded"

    # ------------------------------------------------------------------------------------
    def _evolve(self, circuit: Any, mode: str = "run"):
        n = int(getattr(circuit, "num_qubits", 0))
        if getattr(circuit, "_initial_state", None) is not None:
            raise NotImplementedError("a host-side initial state cannot seed a sharded run (it would not fit one device)")
        ucache = getattr(circuit, "_unitary_cache", {}) or {}
        measures: List[int] = []
        gates = []
        for op in getattr(circuit, "ops", []):
            if not isinstance(op, (list, tuple)) or not op:
                continue
            nm = op[0]
            if nm in _NOT_SHARDED:
                raise NotImplementedError(f"op {nm!r} is outside the sharded path; run it on one GPU (StatevectorEngine)")
            if nm == "measure_z":
                measures.append(int(op[1]))
                continue
            fixed = tuple(float(a.detach().cpu()) if isinstance(a, torch.Tensor) else a for a in op)
            g = lower_op(fixed, n, mode=mode, unitary_cache=ucache)   # unknown ops: silently skipped (engine.py:372-374)
            if g is not None:
                gates.append(g)
        from .fuse import fuse
        # one shard + one exchange buffer per engine, reused by every call (a fresh pair per call would hold four
        # shard-sized buffers from the second call on: half the largest state the devices can run)
        st = self.last_state
        if st is None or st.n != n:
            self.last_state = st = None   # release the old buffers BEFORE allocating the new ones
            st = ShardedState(n, self.dtype, self.device, self.group, backend=self.local_backend)
        plan = plan_sharded(fuse(gates), n, st.g)
        st.init_zero()
        st.run(plan, cache=False)
        self.last_state = st
        self.last_exchanges = plan.n_exchanges
        return st, measures

    def _uniforms(self, shots: int, kwargs: Dict[str, Any]) -> torch.Tensor:
        """The same uniforms on every rank: host-supplied ones as they are; else rank 0 draws them (fresh generator like
        nb.rng(None), engine.py:381, or ``seed``) and broadcasts."""
        u = kwargs.get("uniforms")
        if u is not None:
            return torch.from_numpy(np.ascontiguousarray(np.asarray(u, dtype=np.float64).reshape(-1)))
        t = torch.from_numpy(np.random.default_rng(kwargs.get("seed")).random(shots)).to(self.device)
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.broadcast(t, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        return t

    def run(self, circuit: Any, shots: int | None = None, **kwargs: Any) -> Dict[str, Any]:
        shots = int(shots or 0)
        if kwargs.get("three_level"):
            raise NotImplementedError("three_level mode is outside the B200 hot path")
        if kwargs.get("use_noise"):
            raise NotImplementedError("noise models are not part of the sharded path; run them on one GPU")
        n = int(getattr(circuit, "num_qubits", 0))
        st, measures = self._evolve(circuit)
        if shots > 0 and measures:
            counts = st.counts(self._uniforms(shots, kwargs))
            return {"result": counts, "metadata": {"shots": shots, "backend": self.backend_label, "three_level": False}}
        expectations: Dict[str, float] = {}
        if measures:
            z = st.expect_z_all().cpu().numpy()
            for q in measures:
                expectations[f"Z{q}"] = float(z[n - 1 - q])
        return {"expectations": expectations, "metadata": {"shots": shots, "backend": self.backend_label}}

    def state_shard(self, circuit: Any):
        """(this rank's slice of psi in LOGICAL index order, global_base): amplitudes global_base .. global_base + 2^n_local
        of the final state (engine.py:897-1039 for the part of the state this rank owns; the layout is restored first)."""
        st, _ = self._evolve(circuit, "state")
        st.restore_layout()
        return st.state, st.global_base

    def state(self, circuit: Any) -> Any:
        """The reference's driver asks for the full state after a shots == 0 run (driver.py:115-126).  A sharded state is
        never gathered: with one rank this is the device tensor; otherwise it raises (the driver then reports no
        ``probabilities`` / ``statevector``) -- use ``state_shard``."""
        st, _ = self._evolve(circuit, "state")
        if st.world == 1:
            return st.state
        raise NotImplementedError("the state is sharded over ranks and is not gathered; use state_shard(circuit)")

    def amplitude(self, circuit: Any, bitstring: str) -> complex:
        """psi[int(bitstring, 2)] (engine.py:1052-1060): the owner rank reads one amplitude, everyone gets it."""
        n = int(getattr(circuit, "num_qubits", 0))
        if len(bitstring) != n:
            raise ValueError(f"bitstring of length {len(bitstring)} for {n} qubits")
        st, _ = self._evolve(circuit, "state")
        idx = int(bitstring, 2) if n else 0
        pidx = 0
        for l in range(n):
            pidx |= ((idx >> l) & 1) << st.phys[l]
        out = torch.zeros(2, dtype=torch.float64, device=self.device)
        if (pidx >> st.n_local) == st.rank:
            a = st.state[pidx & ((1 << st.n_local) - 1)]
            out[0], out[1] = a.real.to(torch.float64), a.imag.to(torch.float64)
        if st.world > 1:
            dist.all_reduce(out, group=self.group)
        v = out.cpu().numpy()
        return complex(float(v[0]), float(v[1]))

    def expval(self, circuit: Any, obs: Any, **kwargs: Any) -> float:
        """<psi|H|psi> for a PauliSum, a [(coeff, [(P, q), ...]), ...] list or an OpenFermion-style operator
        (engine.py:475-484 without the sparse matrix)."""
        from .pauli import PauliSum
        n = int(getattr(circuit, "num_qubits", 0))
        if isinstance(obs, PauliSum):
            ham = obs
        elif hasattr(obs, "terms"):
            ham = PauliSum.from_qubit_operator(n, obs)
        else:
            ham = PauliSum.from_pauli_list(n, obs)
        st, _ = self._evolve(circuit, "state")
        return float(st.expect_pauli_sum(ham).real)

    # ------------------------------------------------------------------------------------
    _SHIFT_OPS = ("rx", "ry", "rz", "rxx", "ryy", "rzz")     # exp(-i theta/2 P), P^2 = 1: the two-term shift rule is exact

    def energy_and_grad(self, n: int, template: Sequence[Sequence[Any]], obs: Any, params: Sequence[float]) -> Tuple[float, np.ndarray]:
        """E(theta) and dE/dtheta by the parameter-shift rule (kernels/common.py:11-23, compiler/gradients/
        parameter_shift.py:9-36) on the sharded state; ``template`` holds ``vqe.Param(index, scale)`` placeholders.  The
        shift is applied per gate OCCURRENCE (angle +- pi/2, weighted by the placeholder's scale), so parameters that
        are shared between gates or scaled differentiate exactly; 1 + 2 * occurrences collective evaluations."""
        from .circuits import Circuit
        from .vqe import Param
        theta = np.asarray(params, dtype=np.float64).reshape(-1)
        occ = []
        for k, op in enumerate(template):
            refs = [a for a in op if isinstance(a, Param)]
            if refs:
                if op[0] not in self._SHIFT_OPS or len(refs) != 1:
                    raise NotImplementedError(f"no two-term shift rule for parametrised op {op[0]!r}")
                occ.append((k, refs[0]))

        def energy(shift_at: int = -1, delta: float = 0.0) -> float:
            ops = []
            for k, op in enumerate(template):
                ops.append(tuple((a.scale * float(theta[a.index]) + (delta if k == shift_at else 0.0)) if isinstance(a, Param) else a
                                 for a in op))
            return self.expval(Circuit(n, ops), obs)

        e0 = energy()
        grad = np.zeros_like(theta)
        for k, ref in occ:
            grad[ref.index] += ref.scale * 0.5 * (energy(k, 0.5 * np.pi) - energy(k, -0.5 * np.pi))
        return e0, grad
