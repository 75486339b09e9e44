"""The OpenMP C restatement (cpu_baseline port) against the numpy oracle and the golden fixtures."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import sv_oracle as O
from tests.conftest import random_ops


@pytest.mark.parametrize("n", [1, 2, 5, 11, 16])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_c_oracle_matches_numpy_oracle(n, dtype):
    ops = random_ops(np.random.default_rng(n), n, 100)
    ref, _ = O.evolve_ops(n, ops, mode="run")
    psi = CO.new_state(n, dtype)
    CO.apply_ops(psi, n, ops, "run")
    assert np.abs(psi - ref).max() < (1e-12 if dtype == np.complex128 else 2e-5)
    for q in range(n):
        assert abs(CO.expect_z(psi, q, n) - O.expect_z(psi.astype(np.complex128), q, n)) < 1e-6


def test_c_oracle_golden(golden):
    c = golden["circuits"]["rand12"]
    psi = CO.new_state(12)
    CO.apply_ops(psi, 12, c["ops"], "state")
    assert np.abs(psi - golden["rand12_state"]).max() < 1e-12
    assert CO.max_threads() >= 1
