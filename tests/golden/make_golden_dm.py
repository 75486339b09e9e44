"""Golden fixture for the density-matrix simulator, generated from the LIVE reference (build container only).

    cd /tmp && PYTHONDONTWRITEBYTECODE=1 python /root/repo/tests/golden/make_golden_dm.py

Runs the reference's DensityMatrixEngine (devices/simulators/density_matrix/engine.py) on a seeded circuit without
noise and with each per-gate noise model, shots = 0 (expectations) and shots > 0 (counts; the unseeded Generator is
replaced by a seeded one), plus a Kraus / project_z / reset circuit -> tests/golden/reference_dm.json.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, "/root/reference/src")
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))

import tyxonq as tq  # noqa: E402
from tyxonq.devices.simulators.density_matrix.engine import DensityMatrixEngine  # noqa: E402


def main() -> None:
    tq.set_backend("numpy")
    rng = np.random.default_rng(2024)
    n, shots = 3, 500
    ops = [("h", 0), ("rx", 1, 0.7), ("cx", 0, 1), ("ry", 2, -1.1), ("cz", 1, 2), ("rz", 0, 0.45), ("s", 1), ("cry", 2, 0, 0.9),
           ("x", 2), ("sdg", 0), ("h", 1), ("cx", 2, 1), ("rx", 0, -0.3)] + [("measure_z", q) for q in range(n)]
    noises = [None, {"type": "depolarizing", "p": 0.03}, {"type": "amplitude_damping", "gamma": 0.08},
              {"type": "phase_damping", "lambda": 0.12}, {"type": "pauli", "px": 0.02, "py": 0.01, "pz": 0.04},
              {"type": "readout", "cals": {0: [[0.95, 0.08], [0.05, 0.92]], 1: [[0.9, 0.1], [0.1, 0.9]]}}]
    out = {"n": n, "shots": shots, "ops": [list(o) for o in ops], "runs": []}
    for nz in noises:
        kw = {} if nz is None else {"use_noise": True, "noise": nz}
        e = DensityMatrixEngine().run(tq.Circuit(n, ops=ops), shots=0, **kw)["expectations"]
        eng = DensityMatrixEngine()
        seed = 31
        eng.backend.rng = lambda s=None, seed=seed: np.random.default_rng(seed)
        c = eng.run(tq.Circuit(n, ops=ops), shots=shots, **kw)["result"]
        nzj = None
        if nz is not None:
            nzj = dict(nz)
            if "cals" in nzj:
                nzj["cals"] = {str(k): v for k, v in nzj["cals"].items()}
        out["runs"].append({"noise": nzj, "expectations": e, "seed": seed, "counts": c})
    # Kraus + project_z + reset
    g = 0.3
    ad = [np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=complex), np.array([[0, np.sqrt(g)], [0, 0]], dtype=complex)]
    c = tq.Circuit(n)
    c.h(0).cx(0, 1).ry(2, 0.7)
    c.kraus(1, ad)
    c.cx(1, 2).rx(0, 0.4)
    c.ops.append(("project_z", 2, 1))
    c.h(2)
    c.kraus(0, ad)
    c.ops.append(("reset", 1))
    c.h(1)
    for q in range(n):
        c.ops.append(("measure_z", q))
    e = DensityMatrixEngine().run(c, shots=0)["expectations"]
    out["kraus"] = {"expectations": e, "gamma": g, "ops": [[o[0], *[x for x in o[1:] if isinstance(x, (int, float, str))]] for o in c.ops]}
    (HERE / "reference_dm.json").write_text(json.dumps(out))
    print(len(out["runs"]), "runs;", out["kraus"]["ops"])


if __name__ == "__main__":
    main()
