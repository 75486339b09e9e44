from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Fixtures generated from the live reference by tests/golden/make_golden.py."""
    data = dict(np.load(ROOT / "tests" / "golden" / "reference_vectors.npz"))
    meta = json.loads((ROOT / "tests" / "golden" / "reference_circuits.json").read_text())
    data["circuits"] = {k: {"n": v["n"], "ops": [tuple(o) for o in v["ops"]]} for k, v in meta["circuits"].items()}
    data["meta"] = meta["meta"]
    return data


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tyxonq_b200 import _lib
    _lib.ensure_device(0)
    return torch.device("cuda", 0)


def random_ops(rng, n, N, *, run_mode=True):
    names1 = ["h", "x", "s", "sdg", "y", "z", "t"]
    p1 = ["rz", "rx", "ry"]
    n2 = ["cx", "cz", "iswap", "swap"]
    p2 = ["rxx", "ryy", "rzz"] + (["cry"] if run_mode else [])
    ops = []
    for _ in range(N):
        r = rng.integers(4)
        if r == 0:
            ops.append((names1[rng.integers(len(names1))], int(rng.integers(n))))
        elif r == 1:
            ops.append((p1[rng.integers(3)], int(rng.integers(n)), float(rng.uniform(-3, 3))))
        elif r == 2 and n >= 2:
            a, b = rng.choice(n, 2, replace=False)
            ops.append((n2[rng.integers(4)], int(a), int(b)))
        elif n >= 2:
            a, b = rng.choice(n, 2, replace=False)
            ops.append((p2[rng.integers(len(p2))], int(a), int(b), float(rng.uniform(-3, 3))))
    return ops


class FakeCircuit:
    """Duck-typed stand-in for tyxonq.Circuit (num_qubits, ops, caches, initial state)."""

    def __init__(self, n, ops, inputs=None, unitary_cache=None, kraus_cache=None):
        self.num_qubits = n
        self.ops = list(ops)
        self._initial_state = inputs
        self._unitary_cache = unitary_cache or {}
        self._kraus_cache = kraus_cache or {}
