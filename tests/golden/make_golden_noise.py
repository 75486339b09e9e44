"""Golden fixture for noise on the statevector path, generated from the LIVE reference (build container only).

    cd /tmp && PYTHONDONTWRITEBYTECODE=1 python /root/repo/tests/golden/make_golden_noise.py

  * counts of StatevectorEngine.run(shots, use_noise=True, noise={readout | depolarizing}) (engine.py:377-465) with the
    unseeded Generator replaced by a seeded one (uniforms = default_rng(seed).random(shots))
  * final states of Circuit.kraus trajectories (core/ir/circuit.py:1219, kernels/statevector.py:132-218) through
    StatevectorEngine.state for a table of status draws
-> tests/golden/reference_noise.json
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
from tyxonq.devices.simulators.statevector.engine import StatevectorEngine  # noqa: E402

from oracle import sv_oracle as O  # noqa: E402  (op-list builders only)


def main() -> None:
    tq.set_backend("numpy")
    rng = np.random.default_rng(99)
    n, shots = 4, 400
    ops = O.hea_ops(n, 2, rng.uniform(-3, 3, 4 * n)) + [("measure_z", q) for q in range(n)]
    cals = {0: [[0.97, 0.06], [0.03, 0.94]], 2: [[0.9, 0.12], [0.1, 0.88]]}
    out = {"n": n, "shots": shots, "ops": [list(o) for o in ops], "cals": {str(k): v for k, v in cals.items()}, "runs": []}
    for noise in ({"type": "readout", "cals": cals}, {"type": "depolarizing", "p": 0.05}, {"type": "depolarizing", "p": 0.9}):
        for seed in (5, 6):
            eng = StatevectorEngine()
            eng.backend.rng = lambda s=None, seed=seed: np.random.default_rng(seed)
            res = eng.run(tq.Circuit(n, ops=ops), shots=shots, use_noise=True, noise=noise)
            nz = dict(noise)
            if "cals" in nz:
                nz["cals"] = {str(k): v for k, v in nz["cals"].items()}
            out["runs"].append({"noise": nz, "seed": seed, "counts": res["result"]})
    # Kraus trajectories: amplitude damping on qubit 1, then gates, then a dephasing-like 3-operator channel on qubit 0
    nk = 3
    g1, g2 = 0.3, 0.2
    ad = [np.array([[1, 0], [0, np.sqrt(1 - g1)]], dtype=complex), np.array([[0, np.sqrt(g1)], [0, 0]], dtype=complex)]
    px, pz = 0.15, g2
    pc = [np.sqrt(1 - px - pz) * np.eye(2, dtype=complex), np.sqrt(px) * np.array([[0, 1], [1, 0]], dtype=complex),
          np.sqrt(pz) * np.array([[1, 0], [0, -1]], dtype=complex)]
    status = rng.random((2, 8))
    status[0, 0], status[1, 1] = 0.999, 0.0
    states = []
    for b in range(status.shape[1]):
        c = tq.Circuit(nk)
        c.h(0).cx(0, 1).ry(2, 0.7).cx(1, 2)
        c.kraus(1, ad, status=float(status[0, b]))
        c.rx(0, 0.4).cx(2, 0).rz(1, -0.9)
        c.kraus(0, pc, status=float(status[1, b]))
        c.h(2)
        psi = np.asarray(StatevectorEngine().state(c)).reshape(-1)
        states.append([[float(z.real), float(z.imag)] for z in psi])
    kops = [op for op in c.ops]
    out["kraus"] = {"n": nk, "status": status.tolist(), "states": states,
                    "ops": [[o[0], *[x if not hasattr(x, "shape") else None for x in o[1:]]] for o in kops],
                    "channels": {"ad": [[[float(z.real), float(z.imag)] for z in k.reshape(-1)] for k in ad],
                                 "pc": [[[float(z.real), float(z.imag)] for z in k.reshape(-1)] for k in pc]}}
    (HERE / "reference_noise.json").write_text(json.dumps(out))
    print("runs", len(out["runs"]), "kraus ops", [o for o in out["kraus"]["ops"] if o[0] == "kraus"])


if __name__ == "__main__":
    main()
