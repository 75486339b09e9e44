"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's density-matrix simulator.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this module.

Restates (paths relative to the reference's src/tyxonq/) with full 2^n x 2^n operators instead of the reference's
einsum contractions (small n only):
  * apply_1q_density / apply_2q_density / apply_kraus_density   libs/quantum_library/kernels/density_matrix.py:20-140
  * DensityMatrixEngine.run: op loop, per-gate noise, project_z, sampling and <Z>
                                                                 devices/simulators/density_matrix/engine.py:40-148, 183-222
  * the noise channels                                           libs/quantum_library/noise.py:29-205

Pinned by tests/test_dm_oracle.py against tests/golden/reference_dm.json (expectations and seeded counts of the
reference's own DensityMatrixEngine).
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import sv_oracle as O

C128 = np.complex128
_I = np.eye(2, dtype=C128)
_X = np.array([[0, 1], [1, 0]], dtype=C128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=C128)
_Z = np.array([[1, 0], [0, -1]], dtype=C128)


def full_operator(M: np.ndarray, qubits: Sequence[int], n: int) -> np.ndarray:
    """The 2^n x 2^n matrix of M acting on ``qubits`` (first listed = most significant bit of M's index)."""
    dim = 1 << n
    out = np.empty((dim, dim), dtype=C128)
    for j in range(dim):
        e = np.zeros(dim, dtype=C128)
        e[j] = 1.0
        out[:, j] = O.apply_kq(e, np.asarray(M, dtype=C128), list(qubits), n)
    return out


def channel_kraus(noise: Optional[Dict[str, Any]]) -> Optional[List[np.ndarray]]:
    if not noise:
        return None
    t = str(noise.get("type", "")).lower()
    if t == "depolarizing":
        p = float(noise.get("p", 0.0))
        return [np.sqrt(1 - p) * _I, np.sqrt(p / 3) * _X, np.sqrt(p / 3) * _Y, np.sqrt(p / 3) * _Z]
    if t == "amplitude_damping":
        g = float(noise.get("gamma", noise.get("g", 0.0)))
        return [np.array([[1.0, 0.0], [0.0, np.sqrt(1 - g)]], dtype=C128), np.array([[0.0, np.sqrt(g)], [0.0, 0.0]], dtype=C128)]
    if t == "phase_damping":
        l = float(noise.get("lambda", noise.get("l", 0.0)))
        return [np.array([[1.0, 0.0], [0.0, np.sqrt(1 - l)]], dtype=C128), np.array([[0.0, 0.0], [0.0, np.sqrt(l)]], dtype=C128)]
    if t == "pauli":
        px, py, pz = (float(noise.get(k, 0.0)) for k in ("px", "py", "pz"))
        return [np.sqrt(1 - px - py - pz) * _I, np.sqrt(px) * _X, np.sqrt(py) * _Y, np.sqrt(pz) * _Z]
    return None


def apply_channel(rho: np.ndarray, kraus: Sequence[np.ndarray], q: int, n: int) -> np.ndarray:
    out = np.zeros_like(rho)
    for k in kraus:
        F = full_operator(np.asarray(k, dtype=C128).reshape(2, 2), [q], n)
        out += F @ rho @ F.conj().T
    return out


_ONE = {"h": O.gate_h, "x": O.gate_x, "s": O.gate_s, "sdg": O.gate_sd}
_ONE_P = {"rz": O.gate_rz, "rx": O.gate_rx, "ry": O.gate_ry}
_TWO = {"cx": O.gate_cx_4x4, "cz": O.gate_cz_4x4}


def evolve_density(n: int, ops: Sequence[tuple], noise: Optional[Dict[str, Any]] = None,
                   kraus_cache: Optional[Dict[str, Sequence[np.ndarray]]] = None) -> np.ndarray:
    dim = 1 << n
    rho = np.zeros((dim, dim), dtype=C128)
    rho[0, 0] = 1.0
    ks = channel_kraus(noise)
    kraus_cache = kraus_cache or {}
    for op in ops:
        nm = op[0]
        wires: List[int] = []
        if nm in _ONE:
            F, wires = full_operator(_ONE[nm](), [int(op[1])], n), [int(op[1])]
        elif nm in _ONE_P:
            F, wires = full_operator(_ONE_P[nm](float(op[2])), [int(op[1])], n), [int(op[1])]
        elif nm in _TWO:
            F, wires = full_operator(_TWO[nm](), [int(op[1]), int(op[2])], n), [int(op[1]), int(op[2])]
        elif nm == "cry":
            F, wires = full_operator(O.gate_cry_4x4(float(op[3])), [int(op[1]), int(op[2])], n), [int(op[1]), int(op[2])]
        elif nm in ("project_z", "reset"):
            keep = int(op[2]) if nm == "project_z" else 0
            Pm = np.diag([1.0, 0.0] if keep == 0 else [0.0, 1.0]).astype(C128)
            Fp = full_operator(Pm, [int(op[1])], n)
            r2 = Fp @ rho @ Fp.conj().T
            tr = np.trace(r2)
            rho = r2 / tr if abs(tr) > 0 else r2
            continue
        elif nm == "kraus":
            kk = kraus_cache.get(str(op[2]))
            if kk is not None:
                rho = apply_channel(rho, kk, int(op[1]), n)
            continue
        else:
            continue
        rho = F @ rho @ F.conj().T
        if ks is not None:
            for q in wires:
                rho = apply_channel(rho, ks, q, n)
    return rho


def run_density(n: int, ops: Sequence[tuple], shots: int = 0, *, use_noise: bool = False, noise: Optional[Dict[str, Any]] = None,
                uniforms: Optional[np.ndarray] = None, kraus_cache: Optional[Dict[str, Sequence[np.ndarray]]] = None) -> Dict[str, Any]:
    """engine.py:40-148 with the Generator's uniforms made explicit."""
    rho = evolve_density(n, ops, noise if use_noise else None, kraus_cache)
    measures = [int(op[1]) for op in ops if op[0] == "measure_z"]
    if shots > 0 and measures:
        p = np.real(np.diag(rho)).astype(float).copy()
        p[p < 0.0] = 0.0
        s = float(np.sum(p))
        dim = p.size
        if use_noise:
            nz = noise or {}
            t = str(nz.get("type", "")).lower()
            if t == "readout":
                A = None
                cals = nz.get("cals", {}) or {}
                for q in range(n):
                    m = cals.get(q)
                    m = np.eye(2) if m is None else np.asarray(m)
                    A = m if A is None else np.kron(A, m)
                p = np.asarray(A, dtype=float) @ p
            elif t == "depolarizing":
                alpha = max(0.0, min(1.0, 4.0 * float(nz.get("p", 0.0)) / 3.0))
                p = (1.0 - alpha) * p + alpha * (1.0 / dim)
            p = np.clip(p, 0.0, 1.0)
            s = float(np.sum(p))
        p = p / s if s > 0 else np.full((dim,), 1.0 / dim)
        idx = O.sample_indices_numpy_formula(p, uniforms)
        return {"result": O.counts_from_indices(idx, n)}
    diag = np.real(np.diag(rho))
    exps = {}
    for q in measures:
        bits = (np.arange(1 << n) >> (n - 1 - q)) & 1
        exps[f"Z{q}"] = float(np.sum(diag * (1.0 - 2.0 * bits)))
    return {"expectations": exps}
