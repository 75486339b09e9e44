"""Golden fixture for the shots > 0 energy path, generated from the LIVE reference (build container only).

    cd /tmp && PYTHONDONTWRITEBYTECODE=1 python /root/repo/tests/golden/make_golden_measure.py

Runs the reference's own ``group_hamiltonian_pauli_terms`` (hamiltonian_grouping.py:120-138), its statevector
engine with shots (engine.py:377-465; the unseeded ``nb.rng(None)`` is replaced by a seeded Generator so that the
draw is reproducible) and ``expval_pauli_sum`` (counts_expval.py:87-112) on a seeded Hamiltonian and ansatz, and
stores inputs, the seeds (uniforms = default_rng(seed).random(shots), first 8 stored as a check), the counts and the energies in tests/golden/reference_measure.json.
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
from tyxonq.postprocessing.counts_expval import expval_pauli_sum, term_expectation_from_counts  # noqa: E402

# tyxonq.libs.hamiltonian_encoding/__init__.py imports the chem stack (renormalizer / openfermion, absent here):
# load the grouping module from its file.
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location(
    "_ref_hamiltonian_grouping", "/root/reference/src/tyxonq/libs/hamiltonian_encoding/hamiltonian_grouping.py")
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
group_hamiltonian_pauli_terms = _mod.group_hamiltonian_pauli_terms

from oracle import measure_oracle as MO  # noqa: E402  (only for the prefix ops, which live in a chem runtime that cannot be imported)
from oracle import sv_oracle as O  # noqa: E402


def main() -> None:
    tq.set_backend("numpy")
    rng = np.random.default_rng(77)
    n, shots = 5, 512
    # random Pauli-sum: [(coeff, [(P, q), ...])], a few terms share a measurement basis, one identity term
    ham = [(0.7, [])]
    letters = "XYZ"
    for _ in range(12):
        k = int(rng.integers(1, 4))
        qs = sorted(rng.choice(n, size=k, replace=False).tolist())
        ham.append((float(rng.normal()), [(letters[int(rng.integers(0, 3))], int(q)) for q in qs]))
    ham += [(0.25, [("Z", 0), ("Z", 1)]), (-0.5, [("Z", 0), ("Z", 1)]), (0.3, [("X", 2)]), (0.1, [("X", 2)])]
    identity, groups = group_hamiltonian_pauli_terms(ham, n)
    ansatz = O.hea_ops(n, 2, rng.uniform(-3, 3, 4 * n))
    out = {"n": n, "shots": shots, "hamiltonian": ham, "identity": identity, "ansatz": [list(o) for o in ansatz],
           "groups": [{"bases": list(b), "items": [[[list(t) for t in term], c] for term, c in items]} for b, items in groups.items()],
           "runs": {}}
    for y_rot in ("sdg_h", "rz_h"):
        energy = float(identity)
        runs = []
        for g, (bases, items) in enumerate(groups.items()):
            ops = list(ansatz) + MO.prefix_ops_for_bases(bases, n, y_rot)
            eng = StatevectorEngine()
            seed = 1000 + g
            eng.backend.rng = lambda s=None, seed=seed: np.random.default_rng(seed)
            res = eng.run(tq.Circuit(n, ops=ops), shots=shots)
            counts = res["result"]
            r = expval_pauli_sum(counts, items)
            energy += float(r["energy"])
            runs.append({"seed": seed, "uniforms_head": np.random.default_rng(seed).random(shots)[:8].tolist(), "counts": counts,
                         "energy": float(r["energy"]), "expvals": [float(x) for x in r["expvals"]],
                         "first_term_ev": float(term_expectation_from_counts(counts, [q for q, _ in items[0][0]]))})
        out["runs"][y_rot] = {"energy": energy, "groups": runs}
    (HERE / "reference_measure.json").write_text(json.dumps(out))
    print("groups", len(groups), "energy", {k: v["energy"] for k, v in out["runs"].items()})


if __name__ == "__main__":
    main()
