"""Row f3 of SURVEY.md section 8: noise without leaving the device -- readout / depolarizing sampling through
StatevectorEngine.run, batched Kraus trajectories -- against the live-reference fixture and the oracle."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import noise_oracle as NO
from oracle import sv_oracle as O
from tests.conftest import FakeCircuit
from tests.test_noise_oracle import KRAUS_OPS, fixture_noise, load_fixture

pytestmark = pytest.mark.gpu


def test_engine_noisy_counts_equal_reference_fixture(cuda_device):
    from tyxonq_b200 import StatevectorEngine
    ref, _, _ = load_fixture()
    ops = [tuple(o) for o in ref["ops"]]
    eng = StatevectorEngine(device=cuda_device)
    for run in ref["runs"]:
        u = np.random.default_rng(run["seed"]).random(ref["shots"])
        res = eng.run(FakeCircuit(ref["n"], ops), shots=ref["shots"], use_noise=True, noise=fixture_noise(run), uniforms=u)
        assert res["result"] == run["counts"], run["noise"]


@pytest.mark.parametrize("n,kind", [(9, "readout"), (13, "readout"), (13, "depolarizing")])
def test_noisy_sampling_matches_oracle(cuda_device, n, kind):
    from tyxonq_b200 import StatevectorEngine
    rng = np.random.default_rng(n)
    ops = O.hea_ops(n, 2, rng.uniform(-3, 3, 4 * n)) + [("measure_z", q) for q in range(n)]
    if kind == "readout":
        cals = {}
        for q in rng.choice(n, size=n // 2, replace=False):
            e0, e1 = rng.uniform(0, 0.1, 2)
            cals[int(q)] = np.array([[1 - e0, e1], [e0, 1 - e1]])
        noise = {"type": "readout", "cals": cals}
    else:
        noise = {"type": "depolarizing", "p": 0.08}
    u = rng.random(3000)
    res = StatevectorEngine(device=cuda_device).run(FakeCircuit(n, ops), shots=3000, use_noise=True, noise=noise, uniforms=u)
    assert res["result"] == NO.noisy_counts(n, ops, noise, u)


def test_trajectory_batch_equals_reference_fixture(cuda_device):
    from tyxonq_b200.noise import TrajectoryBatch
    ref, ch, states = load_fixture()
    status = np.array(ref["kraus"]["status"])
    c = FakeCircuit(ref["kraus"]["n"], KRAUS_OPS)
    c._kraus_cache = ch
    tb = TrajectoryBatch(ref["kraus"]["n"], status.shape[1], device=cuda_device)
    got = tb.run(c, status).cpu().numpy()
    assert np.abs(got - states).max() < 1e-10
    assert len(tb.selected) == 2 and len(set(tb.selected[1].tolist())) > 1


def test_trajectory_batch_matches_oracle_on_a_larger_register(cuda_device):
    import torch
    from tyxonq_b200.noise import TrajectoryBatch, reduced_1q
    n, B = 10, 32
    rng = np.random.default_rng(3)
    g = 0.25
    ad = [np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=complex), np.array([[0, np.sqrt(g)], [0, 0]], dtype=complex)]
    ops = []
    for layer in range(3):
        ops += O.hea_ops(n, 1, rng.uniform(-3, 3, 2 * n))[n if layer else 0:]
        ops += [("kraus", int(q), "ad") for q in rng.choice(n, size=3, replace=False)]
    ops += [("project_z", 4, 1), ("rx", 4, 0.3), ("reset", 7), ("h", 7)]
    nk = sum(1 for o in ops if o[0] == "kraus")
    status = rng.random((nk, B))
    c = FakeCircuit(n, ops)
    c._kraus_cache = {"ad": ad}
    want = NO.trajectories(n, ops, {"ad": ad}, status)
    for dt, tol in ((torch.complex128, 1e-10), (torch.complex64, 1e-5)):
        tb = TrajectoryBatch(n, B, device=cuda_device, dtype=dt)
        got = tb.run(c, status).cpu().numpy()
        assert np.abs(got - want).max() < tol
    # the reduced density matrix the selection uses
    rho = reduced_1q(tb.state, 3).cpu().numpy()
    psi = want.reshape(B, -1, 2, 8)
    r00 = (np.abs(psi[:, :, 0, :]) ** 2).sum(axis=(1, 2))
    r01 = (psi[:, :, 0, :] * psi[:, :, 1, :].conj()).sum(axis=(1, 2))
    assert np.abs(rho[:, 0] - r00).max() < 1e-5 and np.abs(rho[:, 2] + 1j * rho[:, 3] - r01).max() < 1e-5
