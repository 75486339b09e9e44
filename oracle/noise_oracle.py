"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's noise on the statevector path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this module.

Restates (paths relative to the reference's src/tyxonq/):
  * the probability-vector noise of StatevectorEngine.run      devices/simulators/statevector/engine.py:389-414
    (readout calibration A = kron(A_0, .., A_{n-1}) applied to p; depolarizing mixing with alpha = 4p/3; clip;
    renormalise; p /= p.sum())  followed by Generator.choice   engine.py:415
  * Monte-Carlo Kraus trajectories                              libs/quantum_library/kernels/statevector.py:132-218
    (oracle/sv_oracle.py apply_kraus), looped over trajectories

Pinned by tests/test_noise_oracle.py against tests/golden/reference_noise.json (counts of the reference's engine with
a seeded Generator, trajectory states of the reference's Circuit.kraus + engine.state).
"""
from __future__ import annotations

from typing import Any, Dict, Sequence

import numpy as np

from . import sv_oracle as O


def noisy_probabilities(p: np.ndarray, n: int, noise: Dict[str, Any] | None) -> np.ndarray:
    """engine.py:389-414 verbatim (numpy backend)."""
    p = np.asarray(p, dtype=float).copy()
    dim = p.size
    if noise:
        ntype = str(noise.get("type", "")).lower()
        if ntype == "readout":
            A = None
            cals = noise.get("cals", {}) or {}
            for q in range(n):
                m = cals.get(q)
                if m is None:
                    m = np.eye(2)
                m = np.asarray(m)
                A = m if A is None else np.kron(A, m)
            p = np.asarray(A, dtype=float) @ p
        elif ntype == "depolarizing":
            pp = float(noise.get("p", 0.0))
            alpha = max(0.0, min(1.0, 4.0 * pp / 3.0))
            p = (1.0 - alpha) * p + alpha * (1.0 / dim)
        p = np.clip(p, 0.0, 1.0)
        s = float(np.sum(p))
        p = p / (s if s > 1e-12 else 1.0)
    if p.sum() > 0:
        p = p / float(p.sum())
    else:
        p = np.full((dim,), 1.0 / dim, dtype=float)
    return p


def noisy_counts(n: int, ops: Sequence[tuple], noise: Dict[str, Any] | None, uniforms: np.ndarray) -> Dict[str, int]:
    """engine.run(shots, use_noise=True, noise=...) with the Generator's uniforms made explicit."""
    psi, _ = O.evolve_ops(n, ops, mode="run")
    p = noisy_probabilities(O.probabilities(psi), n, noise)
    return O.counts_from_indices(O.sample_indices_numpy_formula(p, uniforms), n)


def trajectories(n: int, ops: Sequence[tuple], kraus_cache: Dict[str, Sequence[np.ndarray]], status: np.ndarray,
                 mode: str = "state") -> np.ndarray:
    """[B, 2^n]: trajectory b uses status[k, b] for the k-th kraus op of the circuit."""
    status = np.asarray(status, dtype=np.float64)
    B = status.shape[1]
    out = np.empty((B, 1 << n), dtype=np.complex128)
    for b in range(B):
        k = 0
        fixed = []
        for op in ops:
            if op[0] == "kraus":
                fixed.append(("kraus", op[1], op[2], float(status[k, b])))
                k += 1
            else:
                fixed.append(tuple(op))
        out[b], _ = O.evolve_ops(n, fixed, mode=mode, kraus_cache=kraus_cache)
    return out
