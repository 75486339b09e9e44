"""Row f3 of SURVEY.md section 8: noise on the statevector path without leaving the device.

  * ``TrajectoryBatch`` -- B Monte-Carlo trajectories of a circuit with ``kraus`` ops as ONE [B, 2^n] array.  The
    reference unravels one trajectory at a time and applies every Kraus operator to a copy of the state to get the
    Born probabilities (libs/quantum_library/kernels/statevector.py:132-218).  Here p_i = tr(K_i^+ K_i rho_q) comes
    from the 2x2 reduced density matrix of the target qubit (tqb_reduced_1q: one read of the batch), the operator of
    every trajectory is picked on the host from its own ``status`` draw (status <= cumsum(p / sum p), :198-208) and
    K_sel / sqrt(p_sel) rides in the next fused pass as a gate with one matrix per batch member.
  * ``mix_depolarizing`` / ``apply_readout`` -- the probability-vector noise of StatevectorEngine.run
    (devices/simulators/statevector/engine.py:389-410) on the device: the readout calibration A = kron(A_0, A_1, ...)
    is applied qubit by qubit with the gate kernel (never forming the 2^n x 2^n matrix), and ``sample_probabilities``
    draws from the resulting float64 vector with the same blocked-CDF contract as the state sampler (TQB_F64).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from . import program as P
from .fuse import fuse
from .gates import LGate, _to_np, dense_gate, lower_op
from .planner import TileConfig, compile_program, default_tile


def reduced_1q(state: torch.Tensor, bit: int) -> torch.Tensor:
    """[batch, 4] float64: rho00, rho11, Re rho01, Im rho01 of index bit ``bit`` (rho01 = sum psi_0 conj(psi_1))."""
    ptr, n, batch, dt, stream = P._prep(state)
    out = torch.empty((batch, 4), dtype=torch.float64, device=state.device)
    with torch.cuda.device(state.device):
        _lib.check(_lib.load().tqb_reduced_1q(ptr, n, batch, dt, int(bit), out.data_ptr(), stream))
    return out


def kraus_probabilities(rho: np.ndarray, kraus: Sequence[np.ndarray]) -> np.ndarray:
    """p[b, i] = tr(K_i^+ K_i rho_b) from the reduced density matrices [B, 4]."""
    r00, r11, re01, im01 = rho[:, 0], rho[:, 1], rho[:, 2], rho[:, 3]
    out = np.empty((rho.shape[0], len(kraus)), dtype=np.float64)
    for i, k in enumerate(kraus):
        k = np.asarray(k, dtype=np.complex128).reshape(2, 2)
        e = k.conj().T @ k                     # E = K^+ K (Hermitian): p = E00 r00 + E11 r11 + 2 Re(E10 rho01)
        out[:, i] = e[0, 0].real * r00 + e[1, 1].real * r11 + 2.0 * (e[1, 0].real * re01 - e[1, 0].imag * im01)
    return out


class TrajectoryBatch:
    def __init__(self, n: int, batch: int, *, device: str | torch.device = "cuda", dtype: torch.dtype = torch.complex128,
                 tile: Optional[TileConfig] = None) -> None:
        self.n, self.batch = int(n), int(batch)
        self.device = torch.device(device)
        self.dtype = dtype
        self.itemsize = 16 if dtype == torch.complex128 else 8
        self.tile = tile or default_tile(self.n, self.itemsize, self.batch)
        _lib.ensure_device(self.device.index or 0)
        self.state = torch.empty((self.batch, 1 << self.n), dtype=dtype, device=self.device)
        self.selected: List[np.ndarray] = []   # per kraus op: the operator index every trajectory took
        self.passes = 0

    def _flush(self, pending: List[LGate]) -> None:
        if pending:
            prog = compile_program(fuse(pending), self.n, self.tile, batch_mats=self.batch, itemsize=self.itemsize)
            P.DeviceProgram(prog, self.device, self.dtype).run(self.state)
            self.passes += prog.n_passes
            pending.clear()

    def run(self, circuit: Any, status: Any = None, *, mode: str = "state") -> torch.Tensor:
        """``status``: [n_kraus_ops, batch] uniforms in [0, 1) (one draw per Kraus op and trajectory, in op order), a
        numpy Generator, or None (fresh Generator).  Ops carrying their own status (``("kraus", q, key, s)``) use it
        for every trajectory, like the reference's single-trajectory engine."""
        n = self.n
        ptr, _, b, dt, stream = P._prep(self.state)
        _lib.check(_lib.load().tqb_init_basis(ptr, n, b, dt, 0, 0, stream))
        ucache = getattr(circuit, "_unitary_cache", {}) or {}
        kcache = getattr(circuit, "_kraus_cache", {}) or {}
        rng = status if isinstance(status, np.random.Generator) else (np.random.default_rng() if status is None else None)
        table = None if rng is not None else np.asarray(status, dtype=np.float64).reshape(-1, self.batch)
        self.selected = []
        self.passes = 0
        pending: List[LGate] = []
        ki = 0
        for op in getattr(circuit, "ops", []):
            if not isinstance(op, (list, tuple)) or not op:
                continue
            nm = op[0]
            if nm in ("measure_z", "barrier"):
                continue
            if nm in ("project_z", "reset"):
                self._flush(pending)
                P.project_z(self.state, n - 1 - int(op[1]), (0 if int(op[2]) == 0 else 1) if nm == "project_z" else 0)
                continue
            if nm == "kraus":
                ks = kcache.get(str(op[2]))
                if ks is None:
                    continue
                ks = [np.asarray(_to_np(k), dtype=np.complex128).reshape(2, 2) for k in ks]
                self._flush(pending)
                q = int(op[1])
                p = kraus_probabilities(reduced_1q(self.state, n - 1 - q).cpu().numpy(), ks)
                if len(op) > 3:
                    s = np.full(self.batch, float(op[3]))
                elif rng is not None:
                    s = rng.random(self.batch)
                else:
                    s = table[ki]
                cum = np.cumsum(p / p.sum(axis=1, keepdims=True), axis=1)
                hit = s[:, None] <= cum
                sel = np.where(hit.any(axis=1), hit.argmax(axis=1), 0)      # first i with status <= cum_i, else 0 (:203-208)
                self.selected.append(sel)
                mats = np.stack(ks)[sel] / np.sqrt(p[np.arange(self.batch), sel])[:, None, None]
                pending.append(dense_gate(mats, [q], n, name="kraus"))
                ki += 1
                continue
            g = lower_op(tuple(float(a.detach().cpu()) if isinstance(a, torch.Tensor) else a for a in op), n, mode=mode,
                         unitary_cache=ucache)
            if g is not None:
                pending.append(g)
        self._flush(pending)
        return self.state


# ---- probability-vector noise of StatevectorEngine.run (engine.py:389-410) ----------------------------
def mix_depolarizing(probs: torch.Tensor, p: float) -> torch.Tensor:
    """(1 - alpha) * probs + alpha / dim with alpha = clamp(4p/3, 0, 1), then clip to [0, 1] (engine.py:404-408; the
    two renormalisations that follow in the reference cancel in the sampler's cdf / cdf[-1])."""
    alpha = max(0.0, min(1.0, 4.0 * float(p) / 3.0))
    dim = probs.shape[-1]
    return torch.clamp((1.0 - alpha) * probs + alpha * (1.0 / dim), 0.0, 1.0)


def apply_readout(probs: torch.Tensor, cals: Dict[int, Any], n: int, *, tile: Optional[TileConfig] = None) -> torch.Tensor:
    """p' = kron(A_0, ..., A_{n-1}) p without the 2^n x 2^n matrix (engine.py:393-403): every A_q is a 1-qubit 'gate'
    on the probability vector, run through the fused pass kernel on a complex128 view with zero imaginary parts."""
    mats = {int(q): np.real(np.asarray(_to_np(m))).astype(np.float64).reshape(2, 2) for q, m in (cals or {}).items() if m is not None}
    if not mats:
        return torch.clamp(probs, 0.0, 1.0)
    buf = torch.complex(probs.to(torch.float64), torch.zeros_like(probs, dtype=torch.float64)).contiguous()
    gates = [dense_gate(m.astype(np.complex128), [q], n, name="readout") for q, m in sorted(mats.items()) if 0 <= q < n]
    batch = buf.numel() >> n
    prog = compile_program(fuse(gates), n, tile or default_tile(n, 16, batch), itemsize=16)
    P.DeviceProgram(prog, buf.device, torch.complex128).run(buf)
    return torch.clamp(buf.real, 0.0, 1.0).contiguous()


def sample_probabilities(probs: torch.Tensor, uniforms: Any) -> torch.Tensor:
    """Indices drawn from a float64 probability vector [2^n] or [batch, 2^n] with host uniforms: the blocked-CDF
    contract of ``program.sample`` (cdf / cdf[-1], searchsorted 'right') on probabilities instead of amplitudes."""
    p = probs.to(torch.float64).contiguous()
    n = int(p.shape[-1]).bit_length() - 1
    batch = p.numel() >> n
    u = torch.as_tensor(uniforms, dtype=torch.float64).to(p.device).contiguous()
    shots = int(u.shape[-1])
    if u.numel() != batch * shots:
        raise _lib.TqbError("uniforms must be [shots] or [batch, shots]")
    nc = 1 << (n - min(n, 12))
    prefix = torch.empty((batch, nc + 1), dtype=torch.float64, device=p.device)
    idx = torch.empty(u.shape, dtype=torch.int64, device=p.device)
    with torch.cuda.device(p.device):
        lib = _lib.load()
        stream = _lib.current_stream_ptr(p.device)
        _lib.check(lib.tqb_cdf_chunks(p.data_ptr(), n, batch, _lib.TQB_F64, prefix.data_ptr(), stream))
        _lib.check(lib.tqb_sample(p.data_ptr(), n, batch, _lib.TQB_F64, prefix.data_ptr(), u.data_ptr(), shots, idx.data_ptr(), stream))
    return idx
