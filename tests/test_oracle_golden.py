"""The oracle against (a) fixtures generated from the live reference and (b) the known-answer
vectors held by the reference's own tests.  CPU only."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import sv_oracle as O


def test_gate_matrices_bit_exact(golden):
    ang = golden["gate_angles"]
    for nm in ["gate_rz", "gate_rx", "gate_ry", "gate_phase", "gate_rxx", "gate_ryy", "gate_rzz", "gate_cry_4x4"]:
        mine = np.stack([getattr(O, nm)(float(t)) for t in ang])
        assert np.array_equal(mine, golden[nm]), nm
    for nm in ["gate_h", "gate_x", "gate_s", "gate_sd", "gate_cx_4x4", "gate_cz_4x4", "gate_iswap_4x4", "gate_swap_4x4"]:
        assert np.array_equal(getattr(O, nm)(), golden[nm]), nm


@pytest.mark.parametrize("tag", ["rand5", "rand9", "rand12"])
def test_random_circuit_state_and_expz(golden, tag):
    c = golden["circuits"][tag]
    psi, _ = O.evolve_ops(c["n"], c["ops"], mode="state")
    assert np.abs(psi - golden[f"{tag}_state"]).max() < 1e-14
    ops_m = list(c["ops"]) + [("measure_z", q) for q in range(c["n"])]
    ez = O.run_expectations(c["n"], ops_m)
    assert np.abs(np.array([ez[f"Z{q}"] for q in range(c["n"])]) - golden[f"{tag}_expz"]).max() < 1e-13


def test_builders_match_reference(golden):
    psi, _ = O.evolve_ops(12, O.hea_ops(12, 3, golden["hea12_params"]))
    assert np.abs(psi - golden["hea12_state"]).max() < 1e-14
    psi, _ = O.evolve_ops(10, O.hwe_ry_ops(10, 4, golden["hwe10_params"]))
    assert np.abs(psi - golden["hwe10_state"]).max() < 1e-14
    psi, _ = O.evolve_ops(10, O.qaoa_ring_ops(10, 3, golden["qaoa10_params"]))
    assert np.abs(psi - golden["qaoa10_state"]).max() < 1e-14
    terms, w = O.tfim_terms(8, 1.0, 1.0)
    ops = O.trotter_ops(terms, w, 1.0, 3)
    psi, _ = O.evolve_ops(8, ops)
    assert np.abs(psi - golden["trot8_state"]).max() < 1e-14
    ez = O.run_expectations(8, ops)
    assert np.abs(np.array([ez[f"Z{q}"] for q in range(8)]) - golden["trot8_expz"]).max() < 1e-13


def test_kqubit_kraus_project(golden):
    psi = golden["k7_psi"]
    for k in (1, 2, 3, 4):
        out = O.apply_kq(psi, golden[f"k7_U{k}"], golden[f"k7_q{k}"].tolist(), 7)
        assert np.abs(out - golden[f"k7_out{k}"]).max() < 1e-15
    for s, ref in zip(golden["k7_kraus_status"], golden["k7_kraus_out"]):
        out = O.apply_kraus(psi, list(golden["k7_kraus_ops"]), 3, 7, float(s))
        assert np.abs(out - ref).max() < 1e-15
    assert np.abs(O.project_z(psi, 2, 0, 7) - golden["k7_proj_out"][0]).max() < 1e-15
    assert np.abs(O.project_z(psi, 5, 1, 7) - golden["k7_proj_out"][1]).max() < 1e-15


@pytest.mark.parametrize("tag", ["rand9", "rand12"])
def test_sampler_matches_generator_choice(golden, tag):
    """Blocked-CDF sampler == numpy Generator.choice on the reference's probabilities, same uniforms."""
    st = golden[f"{tag}_state"]
    u = np.random.default_rng(golden["meta"]["sample_seed"]).random(golden["meta"]["sample_shots"])
    assert np.array_equal(O.sample_indices_numpy_formula(np.abs(st) ** 2, u), golden[f"{tag}_sample_idx"])
    idx = O.sample_indices(O.probabilities(st), u)
    assert np.array_equal(idx, golden[f"{tag}_sample_idx"])
    # blocked order with tiny blocks is still the same draw on these inputs
    assert np.array_equal(O.sample_indices(O.probabilities(st), u, block=16), golden[f"{tag}_sample_idx"])


def test_pauli_sum(golden):
    terms = golden["pauli6_terms"].tolist()
    w = golden["pauli6_w"].tolist()
    psi = golden["pauli6_psi"]
    assert abs(O.expect_pauli_sum(psi, terms, w) - float(golden["pauli6_energy"])) < 1e-13
    assert np.abs(O.apply_pauli_sum(psi, terms, w) - golden["pauli6_hpsi"]).max() < 1e-13


# ---- known-answer vectors from the reference's own tests -------------------------------------
def test_kat_bell_and_big_endian():
    # tests_core_module/test_devices_simulators_gates.py:17, test_statevector_engine_probs.py:7-29
    psi, _ = O.evolve_ops(2, [("h", 0), ("cx", 0, 1)])
    assert np.allclose(psi, [2 ** -0.5, 0, 0, 2 ** -0.5], atol=1e-12)
    assert abs(O.expect_z(psi, 0, 2)) < 1e-12 and abs(O.expect_z(psi, 1, 2)) < 1e-12
    p = np.abs(O.evolve_ops(2, [("h", 0)])[0]) ** 2
    assert abs(p[0] + p[2] - 1.0) < 1e-12 and p[1] == 0 and p[3] == 0


def test_kat_unitary_vectors():
    # tests_core_module/test_circuit_unitary.py:24-26, 43-44, 98-101
    sx = 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]])
    psi, _ = O.evolve_ops(1, [("unitary", 0, "k")], unitary_cache={"k": sx})
    assert np.allclose(psi, [0.5 + 0.5j, 0.5 - 0.5j])
    psi, _ = O.evolve_ops(2, [("x", 0), ("unitary", 0, 1, "k")], unitary_cache={"k": O.gate_iswap_4x4()})
    assert np.allclose(psi, [0, 1j, 0, 0])  # iSWAP|10> = i|01>
    psi, _ = O.evolve_ops(2, [("h", 0), ("h", 1), ("unitary", 0, 1, "k")], unitary_cache={"k": O.gate_iswap_4x4()})
    assert np.allclose(psi, [0.5, 0.5j, 0.5j, 0.5])


def test_kat_expectations():
    # tests_core_module/test_circuit_expectation.py:16-69, test_noise_integration.py:13-20,
    # test_circuit_state_autograd.py (<Z> = cos(0.5) after rx(0.5))
    psi, _ = O.evolve_ops(1, [("h", 0)])
    assert abs(O.expect_pauli_sum(psi, [[1]], [1.0]) - 1.0) < 1e-10
    bell, _ = O.evolve_ops(2, [("h", 0), ("cx", 0, 1)])
    assert abs(O.expect_pauli_sum(bell, [[3, 3]], [1.0]) - 1.0) < 1e-10
    assert abs(O.expect_pauli_sum(bell, [[1, 1]], [1.0]) - 1.0) < 1e-10
    ez = O.run_expectations(1, [("rx", 0, np.pi), ("measure_z", 0)])
    assert abs(ez["Z0"] + 1.0) < 1e-8
    ez = O.run_expectations(1, [("rx", 0, 0.5), ("measure_z", 0)])
    assert abs(ez["Z0"] - np.cos(0.5)) < 1e-12
    # 3-qubit TFI energy of |000>: -sum Z_i = -3
    psi = O.init_statevector(3)
    assert abs(O.expect_pauli_sum(psi, [[3, 0, 0], [0, 3, 0], [0, 0, 3]], [-1, -1, -1]) + 3.0) < 1e-10


def test_quirks():
    # unknown ops are skipped in both interpreters; cry only in run(); run() ignores the initial state
    a, _ = O.evolve_ops(2, [("y", 0), ("z", 1), ("t", 0), ("cy", 0, 1)])
    assert np.array_equal(a, O.init_statevector(2))
    s, _ = O.evolve_ops(2, [("x", 0), ("cry", 0, 1, 1.0)], mode="state")
    r, _ = O.evolve_ops(2, [("x", 0), ("cry", 0, 1, 1.0)], mode="run")
    assert np.allclose(s, [0, 0, 1, 0]) and not np.allclose(r, s)
    init = np.array([0, 1, 0, 0], dtype=complex)
    assert np.allclose(O.evolve_ops(2, [], mode="state", initial=init)[0], init)
    assert np.allclose(O.evolve_ops(2, [], mode="run", initial=init)[0], [1, 0, 0, 0])
    # project_z / reset (tests_core_module/test_mid_measure_postselect.py:9)
    ez = O.run_expectations(2, [("h", 0), ("cx", 0, 1), ("project_z", 0, 0), ("reset", 1), ("measure_z", 0), ("measure_z", 1)])
    assert abs(ez["Z0"] - 1) < 1e-12 and abs(ez["Z1"] - 1) < 1e-12


def test_counts_helpers():
    counts = O.counts_from_indices(np.array([0, 3, 3, 1]), 2)
    assert counts == {"00": 1, "01": 1, "11": 2}
    assert abs(O.term_expectation_from_counts(counts, [0, 1]) - (1 - 1 + 2) / 4) < 1e-15
    p = np.array([0.25, 0.25, 0.0, 0.5])
    assert abs(O.zproduct_from_probabilities(p, [0], 2) - (0.5 - 0.5)) < 1e-15
    assert abs(O.zproduct_from_probabilities(p, [0, 1], 2) - (0.25 - 0.25 + 0.5)) < 1e-15


def test_tfim_vqe_energy_zero_params():
    # examples/vqetfim_benchmark.py with param0 = zeros: |0..0> -> h*n = -10
    assert abs(O.tfim_vqe_energy(10, 1, np.zeros((2, 10))) + 10.0) < 1e-12
