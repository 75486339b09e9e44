"""Row f4 of SURVEY.md section 8: the density-matrix simulator on the statevector kernels (rho as a 2n-bit vector),
against the live-reference fixture and the oracle."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import dm_oracle as D
from oracle import sv_oracle as O
from tests.conftest import FakeCircuit
from tests.test_dm_oracle import dm_noise, load_dm_fixture

pytestmark = pytest.mark.gpu


def test_density_engine_equals_reference_fixture(cuda_device):
    from tyxonq_b200.density import DensityMatrixEngine
    ref, kops, cache = load_dm_fixture()
    n, ops = ref["n"], [tuple(o) for o in ref["ops"]]
    eng = DensityMatrixEngine(device=cuda_device)
    assert eng.name == "density_matrix" and eng.capabilities == {"supports_shots": True}
    for run in ref["runs"]:
        kw = dm_noise(run)
        res = eng.run(FakeCircuit(n, ops), shots=0, **kw)
        for k, v in run["expectations"].items():
            assert abs(res["expectations"][k] - v) < 1e-10, (run["noise"], k)
        u = np.random.default_rng(run["seed"]).random(ref["shots"])
        assert eng.run(FakeCircuit(n, ops), shots=ref["shots"], uniforms=u, **kw)["result"] == run["counts"], run["noise"]
    res = eng.run(FakeCircuit(n, kops, kraus_cache=cache), shots=0)
    for k, v in ref["kraus"]["expectations"].items():
        assert abs(res["expectations"][k] - v) < 1e-10


@pytest.mark.parametrize("n,noise", [(5, None), (6, {"type": "depolarizing", "p": 0.02}), (6, {"type": "amplitude_damping", "gamma": 0.1}),
                                     (7, {"type": "pauli", "px": 0.01, "py": 0.02, "pz": 0.03})])
def test_density_matrix_matches_oracle(cuda_device, n, noise):
    import torch
    from tyxonq_b200.density import DensityMatrixEngine
    from tyxonq_b200.pauli import PauliSum
    rng = np.random.default_rng(n)
    ops = O.hea_ops(n, 2, rng.uniform(-3, 3, 4 * n)) + [("cry", 0, n - 1, 0.8), ("cz", 1, 2), ("s", 0), ("sdg", 3), ("x", 2), ("ry", 1, 0.3)]
    ops += [("measure_z", q) for q in range(n)]
    kw = {"use_noise": True, "noise": noise} if noise else {}
    want = D.evolve_density(n, ops, noise)
    for dt, tol in ((torch.complex128, 1e-10), (torch.complex64, 2e-6)):
        eng = DensityMatrixEngine(device=cuda_device, dtype=dt)
        rho = eng.density(FakeCircuit(n, ops), **kw).cpu().numpy()
        assert np.abs(rho - want).max() < tol
    eng = DensityMatrixEngine(device=cuda_device)
    e = eng.run(FakeCircuit(n, ops), shots=0, **kw)["expectations"]
    e_ref = D.run_density(n, ops, 0, **kw)["expectations"]
    assert max(abs(e[k] - e_ref[k]) for k in e_ref) < 1e-10
    # tr(rho H) through the matrix-free Pauli sum
    terms, weights = O.heisenberg_terms(n, [(i, i + 1) for i in range(n - 1)], hzz=1.0, hxx=0.5, hyy=0.25, hz=0.3)
    H = O.pauli_sum_dense(terms, weights)
    got = eng.expval(FakeCircuit(n, ops), PauliSum.from_codes(terms, weights), **kw)
    assert abs(got - np.real(np.trace(want @ H))) < 1e-9
