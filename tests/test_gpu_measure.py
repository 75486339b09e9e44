"""Row f1 of SURVEY.md section 8: shot-based Pauli-sum energies, all measurement groups and all parameter-shift
variants in one batched launch, against the reference fixture and the oracle (same uniforms => same counts =>
identical energies: the per-term expectation is an integer count divided by shots)."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest

from oracle import measure_oracle as MO
from oracle import sv_oracle as O

pytestmark = pytest.mark.gpu
FIX = Path(__file__).resolve().parent / "golden" / "reference_measure.json"


def _groups(ref):
    return {tuple(g["bases"]): [(tuple((int(q), p) for q, p in term), float(c)) for term, c in g["items"]] for g in ref["groups"]}


@pytest.mark.parametrize("y_rot", ["sdg_h", "rz_h"])
def test_grouped_energy_equals_reference_fixture(cuda_device, y_rot):
    import torch
    from tyxonq_b200 import StatevectorEngine
    from tyxonq_b200.measure import GroupedMeasurement, group_pauli_terms
    from tests.conftest import FakeCircuit
    ref = json.loads(FIX.read_text())
    n, shots = ref["n"], ref["shots"]
    ham = [(c, [(p, q) for p, q in ops]) for c, ops in ref["hamiltonian"]]
    identity, groups = group_pauli_terms(ham, n)
    assert identity == ref["identity"] and groups == _groups(ref) and list(groups) == list(_groups(ref))
    runs = ref["runs"][y_rot]["groups"]
    uniforms = np.stack([np.random.default_rng(r["seed"]).random(shots) for r in runs])
    eng = StatevectorEngine("b200", device=cuda_device)
    psi, _, _ = eng._evolve(FakeCircuit(n, [tuple(o) for o in ref["ansatz"]]), "run")
    gm = GroupedMeasurement(n, groups, identity, y_rotation=y_rot, device=cuda_device)
    e, evs = gm.group_energies(psi, uniforms, want_expvals=True)
    e, evs = e.cpu().numpy(), evs.cpu().numpy()
    for g, r in enumerate(runs):
        assert e[0, g] == r["energy"]
        assert list(evs[0, g, :len(r["expvals"])]) == r["expvals"]
    assert gm.energies(psi, uniforms)[0] == ref["runs"][y_rot]["energy"]
    # complex64 states: same counts unless a uniform sits within rounding of a CDF step; energies agree statistically
    gm32 = GroupedMeasurement(n, groups, identity, y_rotation=y_rot, device=cuda_device, dtype=torch.complex64)
    e32 = gm32.energies(psi.to(torch.complex64), uniforms)[0]
    assert abs(e32 - ref["runs"][y_rot]["energy"]) < 0.05


def test_parameter_shift_from_shots_matches_oracle(cuda_device):
    from tyxonq_b200.measure import GroupedMeasurement, ShotEnergy
    from tyxonq_b200.vqe import Param
    n = 4
    ham = [(0.4, [("Z", 0), ("Z", 1)]), (-0.7, [("X", 1)]), (0.3, [("Y", 0), ("Y", 2)]), (0.2, []), (0.15, [("Z", 2), ("X", 3)]),
           (-0.25, [("Z", 0)]), (0.35, [("X", 1), ("X", 3)])]
    template = [("h", 0), ("ry", 0, Param(0)), ("rx", 1, Param(1)), ("cx", 0, 1), ("rz", 2, Param(2)), ("h", 2), ("rzz", 1, 2, Param(3)),
                ("cx", 2, 3), ("ry", 3, Param(4)), ("rxx", 0, 3, Param(5))]

    def build(p):
        return [("h", 0), ("ry", 0, p[0]), ("rx", 1, p[1]), ("cx", 0, 1), ("rz", 2, p[2]), ("h", 2), ("rzz", 1, 2, p[3]),
                ("cx", 2, 3), ("ry", 3, p[4]), ("rxx", 0, 3, p[5])]

    gm = GroupedMeasurement.from_pauli_list(n, ham, device=cuda_device)
    identity, groups = MO.group_hamiltonian_pauli_terms(ham, n)
    assert gm.G == len(groups) and gm.identity == identity
    se = ShotEnergy(n, template, gm)
    assert se.n_params == 6
    params = np.array([0.3, -0.8, 1.1, 0.45, -0.6, 0.2])
    shots = 300
    u = np.random.default_rng(11).random(((1 + 2 * 6) * gm.G, shots))
    e, g = se.energy_and_grad(params, u)
    e_ref, g_ref = MO.grouped_shot_energy_and_grad(n, build, params, identity, groups, u)
    assert abs(e - e_ref) < 1e-12 and np.abs(g - g_ref).max() < 1e-12
    assert abs(se.energy(params, u[:gm.G]) - e_ref) < 1e-12
    # the batched states are the oracle's states
    st = se.states(np.stack([params, params * 0.5])).cpu().numpy()
    for row, p in zip(st, (params, params * 0.5)):
        assert np.abs(row - O.evolve_ops(n, build(p), mode="run")[0]).max() < 1e-12
    # scaled or shared parameters: energies and states work, the parameter-shift gradient is refused (its rule would be wrong)
    tied = ShotEnergy(n, [("ry", 0, Param(0)), ("rxx", 0, 3, Param(0, -1.0)), ("ry", 3, Param(1, 2.0))], gm)
    tied.energy(np.array([0.3, 0.1]), u[:gm.G])
    with pytest.raises(NotImplementedError):
        tied.energy_and_grad(np.array([0.3, 0.1]), u[:5 * gm.G])
