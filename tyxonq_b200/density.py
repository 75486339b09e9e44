"""Row f4 of SURVEY.md section 8: the density-matrix simulator on the statevector kernels.

The reference's ``DensityMatrixEngine`` (devices/simulators/density_matrix/engine.py:40-148) keeps rho as a dense
2^n x 2^n matrix and applies every gate with a three-operand einsum (libs/quantum_library/kernels/
density_matrix.py:20-140).  Here rho is a 2n-bit vector (rho[r, c] at index r * 2^n + c: the row qubits are wires
0..n-1 of a 2n-wire register, the column qubits wires n..2n-1):

    U rho U^+            = U on wire q, conj(U) on wire n + q                  (two gates of the fused passes)
    sum_i K_i rho K_i^+  = the 4x4 matrix sum_i K_i (x) conj(K_i) on wires (q, n + q)   (one dense 2-qubit gate)
    P rho P / tr         = the same with P (x) P, then a scale by 1 / trace

so the gate list goes through the same host fusion and the same tile-pass kernel as a pure state.  The
probabilities are the diagonal (tqb_dm_diag), <Z_q> and the sampler run on that float64 vector (TQB_F64).
Same class surface and result dicts as the reference (name, capabilities, run; ``expval`` needs OpenFermion there
and takes a ``PauliSum`` here).  Op set of run(): h rz rx ry cx cz cry x s sdg measure_z barrier project_z reset
kraus; anything else is skipped, like the reference.  Noise (``use_noise=True``): depolarizing / amplitude_damping /
phase_damping / pauli after every gate on its wires (engine.py:183-209), readout / depolarizing mixing of the
sampled distribution (engine.py:118-134).
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from . import program as P
from .fuse import fuse
from .gates import C128, LGate, _to_np, dense_gate, lower_op
from .planner import TileConfig, compile_program, default_tile

_GATES = ("h", "rz", "rx", "ry", "cx", "cz", "cry", "x", "s", "sdg")
_I = np.eye(2, dtype=C128)
_X = np.array([[0, 1], [1, 0]], dtype=C128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=C128)
_Z = np.array([[1, 0], [0, -1]], dtype=C128)


def noise_kraus(noise: Optional[Dict[str, Any]]) -> Optional[List[np.ndarray]]:
    """Kraus operators of the per-gate noise model (libs/quantum_library/noise.py:29-205 via engine.py:183-209)."""
    if not noise:
        return None
    t = str(noise.get("type", "")).lower()
    if t == "depolarizing":
        p = float(noise.get("p", 0.0))
        return [math.sqrt(1 - p) * _I, math.sqrt(p / 3) * _X, math.sqrt(p / 3) * _Y, math.sqrt(p / 3) * _Z]
    if t == "amplitude_damping":
        g = float(noise.get("gamma", noise.get("g", 0.0)))
        return [np.array([[1, 0], [0, math.sqrt(1 - g)]], dtype=C128), np.array([[0, math.sqrt(g)], [0, 0]], dtype=C128)]
    if t == "phase_damping":
        l = float(noise.get("lambda", noise.get("l", 0.0)))
        return [np.array([[1, 0], [0, math.sqrt(1 - l)]], dtype=C128), np.array([[0, 0], [0, math.sqrt(l)]], dtype=C128)]
    if t == "pauli":
        px, py, pz = (float(noise.get(k, 0.0)) for k in ("px", "py", "pz"))
        return [math.sqrt(1 - px - py - pz) * _I, math.sqrt(px) * _X, math.sqrt(py) * _Y, math.sqrt(pz) * _Z]
    return None


def superoperator(kraus: Sequence[Any]) -> np.ndarray:
    """sum_i K_i (x) conj(K_i): the channel as a 4x4 matrix on (row bit, column bit)."""
    s = np.zeros((4, 4), dtype=C128)
    for k in kraus:
        k = np.asarray(_to_np(k), dtype=C128).reshape(2, 2)
        s += np.kron(k, k.conj())
    return s


class DensityMatrixEngine:
    name = "density_matrix"
    capabilities = {"supports_shots": True}

    def __init__(self, backend_name: str | None = None, *, device: str | torch.device | None = None,
                 dtype: torch.dtype = torch.complex128, tile: Optional[TileConfig] = None) -> None:
        self.backend_name = backend_name or "numpy"
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self.dtype = dtype
        self.tile = tile
        self.last_passes = 0

    # ------------------------------------------------------------------------------------------
    def _column_gate(self, g: LGate, n: int) -> LGate:
        return LGate(g.kind, tuple(b - n for b in g.bits), np.conj(g.data), pat_a=g.pat_a, pat_b=g.pat_b, name=g.name + "*")

    def _flush(self, rho: torch.Tensor, pending: List[LGate], n: int) -> None:
        if pending:
            item = rho.element_size()
            prog = compile_program(fuse(pending), 2 * n, self.tile or default_tile(2 * n, item, 1), itemsize=item)
            P.DeviceProgram(prog, rho.device, rho.dtype).run(rho)
            self.last_passes += prog.n_passes
            pending.clear()

    def density(self, circuit: Any, **kwargs: Any) -> torch.Tensor:
        """rho as a device tensor [2^n, 2^n] after the circuit (gates, per-gate noise, kraus, project_z / reset)."""
        _lib.ensure_device(self.device.index or 0)
        n = int(getattr(circuit, "num_qubits", 0))
        if 2 * n > 34:
            raise _lib.TqbError("density matrices beyond 17 qubits do not fit one GPU")
        self.last_passes = 0
        rho = P.new_state(2 * n, dtype=self.dtype, device=self.device)      # |0..0><0..0|
        ks_noise = noise_kraus(kwargs.get("noise") if kwargs.get("use_noise") else None)
        s_noise = superoperator(ks_noise) if ks_noise is not None else None
        kcache = getattr(circuit, "_kraus_cache", {}) or {}
        pending: List[LGate] = []

        def channel(s: np.ndarray, q: int) -> None:
            pending.append(dense_gate(s, [q, n + q], 2 * n, name="channel"))

        for op in getattr(circuit, "ops", []):
            if not isinstance(op, (list, tuple)) or not op:
                continue
            nm = op[0]
            if nm in _GATES:
                fixed = tuple(float(a.detach().cpu()) if isinstance(a, torch.Tensor) else a for a in op)
                g = lower_op(fixed, 2 * n, mode="run")
                if g is None:
                    continue
                pending.append(g)
                pending.append(self._column_gate(g, n))
                if s_noise is not None:
                    for q in [int(a) for a in fixed[1:3] if isinstance(a, int)][: 2 if nm in ("cx", "cz", "cry") else 1]:
                        channel(s_noise, q)
            elif nm in ("project_z", "reset"):
                keep = int(op[2]) if nm == "project_z" else 0
                pr = np.diag([1.0, 0.0] if keep == 0 else [0.0, 1.0]).astype(C128)
                channel(np.kron(pr, pr), int(op[1]))
                self._flush(rho, pending, n)
                tr = float(self._trace(rho, n))
                if abs(tr) > 0:
                    P.scale(rho, 1.0 / tr)
            elif nm == "kraus":
                ks = kcache.get(str(op[2]))
                if ks is not None:
                    channel(superoperator(ks), int(op[1]))
            # measure_z, barrier and unknown names: nothing to apply
        self._flush(rho, pending, n)
        return rho.view(1 << n, 1 << n)

    def _diag(self, rho: torch.Tensor, n: int) -> torch.Tensor:
        out = torch.empty(1 << n, dtype=torch.float64, device=rho.device)
        with torch.cuda.device(rho.device):
            _lib.check(_lib.load().tqb_dm_diag(rho.data_ptr(), n, 1, _lib.dtype_code(rho.dtype), out.data_ptr(),
                                               _lib.current_stream_ptr(rho.device)))
        return out

    def _trace(self, rho: torch.Tensor, n: int) -> torch.Tensor:
        d = self._diag(rho, n)
        out = torch.empty(1, dtype=torch.float64, device=rho.device)
        with torch.cuda.device(rho.device):
            _lib.check(_lib.load().tqb_norm2(d.data_ptr(), n, 1, _lib.TQB_F64, out.data_ptr(), _lib.current_stream_ptr(rho.device)))
        return out[0]

    # ------------------------------------------------------------------------------------------
    def run(self, circuit: Any, shots: int | None = None, **kwargs: Any) -> Dict[str, Any]:
        shots = int(shots or 0)
        n = int(getattr(circuit, "num_qubits", 0))
        rho = self.density(circuit, **kwargs)
        measures = [int(op[1]) for op in getattr(circuit, "ops", []) if isinstance(op, (list, tuple)) and op and op[0] == "measure_z"]
        diag = self._diag(rho, n)
        if shots > 0 and measures:
            from . import noise as NZ
            probs = torch.clamp(diag, min=0.0)                       # engine.py:112 p[p < 0] = 0
            if bool(kwargs.get("use_noise", False)):
                nz = kwargs.get("noise", {}) or {}
                t = str(nz.get("type", "")).lower()
                if t == "readout":
                    probs = NZ.apply_readout(probs, nz.get("cals", {}) or {}, n, tile=None)
                elif t == "depolarizing":
                    probs = NZ.mix_depolarizing(probs, float(nz.get("p", 0.0)))
                else:
                    probs = torch.clamp(probs, 0.0, 1.0)
            u = kwargs.get("uniforms")
            if u is None:
                u = np.random.default_rng(kwargs.get("seed")).random(shots)
            idx = NZ.sample_probabilities(probs, np.asarray(u, dtype=np.float64).reshape(-1)).cpu().numpy()
            vals, cnts = np.unique(idx, return_counts=True)
            results = {format(int(v), f"0{n}b") if n else "": int(c) for v, c in zip(vals, cnts)}
            return {"result": results, "metadata": {"shots": shots, "backend": "b200"}}
        expectations: Dict[str, float] = {}
        if measures:
            z = torch.empty(n, dtype=torch.float64, device=diag.device)
            with torch.cuda.device(diag.device):
                _lib.check(_lib.load().tqb_expect_z_bits(diag.data_ptr(), n, 1, _lib.TQB_F64, z.data_ptr(),
                                                         _lib.current_stream_ptr(diag.device)))
            zh = z.cpu().numpy()
            for q in measures:
                expectations[f"Z{q}"] = float(zh[n - 1 - q])
        return {"expectations": expectations, "metadata": {"shots": shots, "backend": "b200"}}

    def expval(self, circuit: Any, obs: Any, **kwargs: Any) -> float:
        """tr(rho H) for a ``PauliSum`` (the reference builds a dense H with OpenFermion, engine.py:150-181, and -- like
        there -- per-gate noise is not applied on this path unless passed explicitly)."""
        from .pauli import PauliSum
        if not isinstance(obs, PauliSum):
            raise TypeError("expval takes a tyxonq_b200.PauliSum")
        n = int(getattr(circuit, "num_qubits", 0))
        rho = self.density(circuit, **kwargs)
        # tr(rho H) = sum_c (H rho)[c, c]: H acts on the row wires of every column = a batch of 2^n column vectors of rho^T
        cols = rho.t().contiguous()                  # [column c][row r] : batch member c is column c of rho
        out = obs.apply(cols)                        # H applied to every column
        d = torch.diagonal(out)                      # (H rho)[c, c]
        return float(d.sum().real.cpu())
