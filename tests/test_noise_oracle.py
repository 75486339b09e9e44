"""oracle/noise_oracle.py against the fixture produced by the live reference (tests/golden/make_golden_noise.py)."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest

from oracle import noise_oracle as NO

FIX = Path(__file__).resolve().parent / "golden" / "reference_noise.json"

KRAUS_OPS = [("h", 0), ("cx", 0, 1), ("ry", 2, 0.7), ("cx", 1, 2), ("kraus", 1, "ad"), ("rx", 0, 0.4), ("cx", 2, 0),
             ("rz", 1, -0.9), ("kraus", 0, "pc"), ("h", 2)]


def load_fixture():
    ref = json.loads(FIX.read_text())
    ch = {k: [np.array([complex(a, b) for a, b in m]).reshape(2, 2) for m in v] for k, v in ref["kraus"]["channels"].items()}
    states = np.array([[complex(a, b) for a, b in row] for row in ref["kraus"]["states"]])
    return ref, ch, states


def fixture_noise(run):
    nz = dict(run["noise"])
    if "cals" in nz:
        nz["cals"] = {int(k): np.array(v) for k, v in nz["cals"].items()}
    return nz


def test_noisy_counts_match_reference():
    ref, _, _ = load_fixture()
    ops = [tuple(o) for o in ref["ops"]]
    for run in ref["runs"]:
        u = np.random.default_rng(run["seed"]).random(ref["shots"])
        assert NO.noisy_counts(ref["n"], ops, fixture_noise(run), u) == run["counts"], run["noise"]


def test_kraus_trajectories_match_reference():
    ref, ch, states = load_fixture()
    # the reference circuit of the fixture really has the two kraus ops where KRAUS_OPS puts them
    assert [o[0] for o in ref["kraus"]["ops"]] == [o[0] for o in KRAUS_OPS]
    got = NO.trajectories(ref["kraus"]["n"], KRAUS_OPS, ch, np.array(ref["kraus"]["status"]))
    assert np.abs(got - states).max() < 1e-13
    # different draws select different operators
    assert len({tuple(np.round(r, 6)) for r in states}) > 2
