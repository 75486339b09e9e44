"""The shots > 0 energy oracle (oracle/measure_oracle.py) against the fixture produced by the live reference
(tests/golden/make_golden_measure.py: its grouping, its engine's sampled counts, its expval_pauli_sum)."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest

from oracle import measure_oracle as MO
from oracle import sv_oracle as O

FIX = Path(__file__).resolve().parent / "golden" / "reference_measure.json"


@pytest.fixture(scope="module")
def ref():
    return json.loads(FIX.read_text())


def _groups(ref):
    return {tuple(g["bases"]): [(tuple((int(q), p) for q, p in term), float(c)) for term, c in g["items"]] for g in ref["groups"]}


def test_grouping_matches_reference(ref):
    ham = [(c, [(p, q) for p, q in ops]) for c, ops in ref["hamiltonian"]]
    identity, groups = MO.group_hamiltonian_pauli_terms(ham, ref["n"])
    assert identity == ref["identity"]
    want = _groups(ref)
    assert list(groups.keys()) == list(want.keys())          # insertion order = submission order of the circuits
    assert groups == want


@pytest.mark.parametrize("y_rot", ["sdg_h", "rz_h"])
def test_counts_and_energies_match_reference(ref, y_rot):
    n, shots = ref["n"], ref["shots"]
    groups = _groups(ref)
    ansatz = [tuple(o) for o in ref["ansatz"]]
    runs = ref["runs"][y_rot]["groups"]
    uniforms = np.stack([np.random.default_rng(r["seed"]).random(shots) for r in runs])
    for r, u in zip(runs, uniforms):
        assert np.array_equal(u[:8], np.array(r["uniforms_head"]))
    # per group: identical counts dict, expvals and energy contribution
    for (bases, items), r, u in zip(groups.items(), runs, uniforms):
        psi, _ = O.evolve_ops(n, ansatz + MO.prefix_ops_for_bases(bases, n, y_rot), mode="run")
        counts = O.counts_from_indices(O.sample_indices(O.probabilities(psi), u), n)
        assert counts == r["counts"]
        e, evs = MO.expval_pauli_sum(counts, items)
        assert e == r["energy"] and evs == r["expvals"]
        assert O.term_expectation_from_counts(counts, [q for q, _ in items[0][0]]) == r["first_term_ev"]
    total, contribs = MO.grouped_shot_energy(n, ansatz, ref["identity"], groups, uniforms, y_rot)
    assert total == ref["runs"][y_rot]["energy"]
    assert contribs == [r["energy"] for r in runs]


def test_parameter_shift_on_shots_converges_to_the_exact_gradient():
    """hea_device_runtime.py:180-262 restated: with many shots the estimate approaches d<H>/dtheta."""
    n = 3
    ham = [(0.4, [("Z", 0), ("Z", 1)]), (-0.7, [("X", 1)]), (0.3, [("Y", 0), ("Y", 2)]), (0.2, [])]
    identity, groups = MO.group_hamiltonian_pauli_terms(ham, n)

    def build(p):
        return [("ry", 0, p[0]), ("ry", 1, p[1]), ("cx", 0, 1), ("ry", 2, p[2]), ("cx", 1, 2)]

    codes = {"I": 0, "X": 1, "Y": 2, "Z": 3}
    terms, weights = [], []
    for c, ops in ham:
        ps = [0] * n
        for p, q in ops:
            ps[q] = codes[p]
        terms.append(ps)
        weights.append(c)

    def exact(p):
        psi, _ = O.evolve_ops(n, build(p), mode="run")
        return O.expect_pauli_sum(psi, terms, weights)

    params = np.array([0.3, -0.8, 1.1])
    shots = 40000
    u = np.random.default_rng(0).random(((1 + 2 * 3) * len(groups), shots))
    e, g = MO.grouped_shot_energy_and_grad(n, build, params, identity, groups, u)
    assert abs(e - exact(params)) < 0.02
    assert np.abs(g - O.parameter_shift_gradient(exact, params)).max() < 0.02
