"""oracle/dm_oracle.py against the fixture produced by the live reference's DensityMatrixEngine
(tests/golden/make_golden_dm.py)."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from oracle import dm_oracle as D

FIX = Path(__file__).resolve().parent / "golden" / "reference_dm.json"


def load_dm_fixture():
    ref = json.loads(FIX.read_text())
    g = ref["kraus"]["gamma"]
    ad = [np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=complex), np.array([[0, np.sqrt(g)], [0, 0]], dtype=complex)]
    kops, cache = [], {}
    for o in ref["kraus"]["ops"]:
        kops.append(tuple(o))
        if o[0] == "kraus":
            cache[str(o[2])] = ad
    return ref, kops, cache


def dm_noise(run):
    nz = run["noise"]
    if nz is None:
        return {}
    nz = dict(nz)
    if "cals" in nz:
        nz["cals"] = {int(k): np.array(v) for k, v in nz["cals"].items()}
    return {"use_noise": True, "noise": nz}


def test_density_oracle_matches_reference_engine():
    ref, kops, cache = load_dm_fixture()
    n, ops = ref["n"], [tuple(o) for o in ref["ops"]]
    for run in ref["runs"]:
        kw = dm_noise(run)
        e = D.run_density(n, ops, 0, **kw)["expectations"]
        for k, v in run["expectations"].items():
            assert abs(e[k] - v) < 1e-12, (run["noise"], k)
        u = np.random.default_rng(run["seed"]).random(ref["shots"])
        assert D.run_density(n, ops, ref["shots"], uniforms=u, **kw)["result"] == run["counts"], run["noise"]
    e = D.run_density(n, kops, 0, kraus_cache=cache)["expectations"]
    for k, v in ref["kraus"]["expectations"].items():
        assert abs(e[k] - v) < 1e-12
    rho = D.evolve_density(n, kops, None, cache)
    assert abs(np.trace(rho) - 1) < 1e-12 and np.abs(rho - rho.conj().T).max() < 1e-12
