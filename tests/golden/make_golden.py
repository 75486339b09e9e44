"""Generate golden fixtures from the LIVE reference (run in the build container only).

    cd /tmp && PYTHONDONTWRITEBYTECODE=1 python /root/repo/tests/golden/make_golden.py

Imports TyxonQ from /root/reference/src (numpy backend), runs the reference's own engine and
kernels on seeded inputs and stores inputs + outputs in tests/golden/*.npz / *.json.  The GPU
box has no /root/reference; tests only read the committed fixtures.
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
from tyxonq.libs.quantum_library.kernels import gates as RG  # noqa: E402
from tyxonq.libs.quantum_library.kernels import statevector as RS  # noqa: E402
from tyxonq.libs.quantum_library.kernels.pauli import pauli_string_sum_dense  # noqa: E402

# tyxonq.libs.circuits_library/__init__.py pulls in the chem stack (openfermion/pyscf, absent here):
# register a bare package and load the three builder modules from their files.
import importlib.util  # noqa: E402
import types  # noqa: E402

_pkg = types.ModuleType("tyxonq.libs.circuits_library")
_pkg.__path__ = ["/root/reference/src/tyxonq/libs/circuits_library"]
sys.modules["tyxonq.libs.circuits_library"] = _pkg


def _load(name: str):
    spec = importlib.util.spec_from_file_location(
        f"tyxonq.libs.circuits_library.{name}", f"/root/reference/src/tyxonq/libs/circuits_library/{name}.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


example_block = _load("blocks").example_block
build_hwe_ry_ops = _load("blocks").build_hwe_ry_ops
qaoa_ising = _load("qaoa_ising").qaoa_ising
build_trotter_circuit = _load("trotter_circuit").build_trotter_circuit

tq.set_backend("numpy")
from oracle import sv_oracle as O  # noqa: E402  (generators only: op lists must equal the reference builders')


def rand_ops(rng, n, N, with_cry=True):
    names1 = ["h", "x", "s", "sdg", "y", "z", "t"]
    p1 = ["rz", "rx", "ry"]
    n2 = ["cx", "cz", "iswap", "swap"]
    p2 = ["rxx", "ryy", "rzz"] + (["cry"] if with_cry else [])
    ops = []
    for _ in range(N):
        r = rng.integers(4)
        if r == 0:
            ops.append([names1[rng.integers(len(names1))], int(rng.integers(n))])
        elif r == 1:
            ops.append([p1[rng.integers(3)], int(rng.integers(n)), float(rng.uniform(-3, 3))])
        elif r == 2:
            a, b = rng.choice(n, 2, replace=False)
            ops.append([n2[rng.integers(4)], int(a), int(b)])
        else:
            a, b = rng.choice(n, 2, replace=False)
            ops.append([p2[rng.integers(len(p2))], int(a), int(b), float(rng.uniform(-3, 3))])
    return ops


def main() -> None:
    rng = np.random.default_rng(20261017)
    eng = StatevectorEngine("numpy")
    out = {}
    meta = {}

    # 1. gate matrices
    angles = rng.uniform(-2 * np.pi, 2 * np.pi, 6)
    out["gate_angles"] = angles
    for nm in ["gate_rz", "gate_rx", "gate_ry", "gate_phase", "gate_rxx", "gate_ryy", "gate_rzz", "gate_cry_4x4"]:
        out[nm] = np.stack([np.asarray(getattr(RG, nm)(float(t))) for t in angles])
    for nm in ["gate_h", "gate_x", "gate_s", "gate_sd", "gate_cx_4x4", "gate_cz_4x4", "gate_iswap_4x4", "gate_swap_4x4"]:
        out[nm] = np.asarray(getattr(RG, nm)())

    # 2. random circuits through engine.state() and engine.run(shots=0)
    circuits = {}
    for tag, n, N in [("rand5", 5, 60), ("rand9", 9, 120), ("rand12", 12, 200)]:
        ops = rand_ops(rng, n, N)
        c = tq.Circuit(n, ops=[tuple(o) for o in ops])
        out[f"{tag}_state"] = np.asarray(eng.state(c))
        ops_m = ops + [["measure_z", q] for q in range(n)]
        r = eng.run(tq.Circuit(n, ops=[tuple(o) for o in ops_m]), shots=0)
        out[f"{tag}_expz"] = np.array([r["expectations"][f"Z{q}"] for q in range(n)])
        circuits[tag] = {"n": n, "ops": ops}

    # 3. reference circuit builders (op lists AND states)
    n = 12
    p = rng.uniform(-np.pi, np.pi, 2 * 3 * n)
    c = example_block(tq.Circuit(n), p, nlayers=3)
    assert [tuple(o) for o in c.ops] == O.hea_ops(n, 3, p)
    out["hea12_params"] = p
    out["hea12_state"] = np.asarray(eng.state(c))
    n = 10
    p = rng.random((4 + 1) * n)
    c = build_hwe_ry_ops(n, 4, p)
    assert [tuple(o) for o in c.ops if o[0] != "barrier"] == O.hwe_ry_ops(n, 4, p)
    out["hwe10_params"] = p
    out["hwe10_state"] = np.asarray(eng.state(c))
    n = 10
    p = rng.uniform(-np.pi, np.pi, 2 * 3)
    zz = [[1 if q in (i, (i + 1) % n) else 0 for q in range(n)] for i in range(n)]
    c = qaoa_ising(n, 3, zz, [1.0] * n, list(p))
    assert [tuple(o) for o in c.ops] == O.qaoa_ring_ops(n, 3, p), "qaoa op list mismatch"
    out["qaoa10_params"] = p
    out["qaoa10_state"] = np.asarray(eng.state(c))
    n = 8
    terms, w = O.tfim_terms(n, 1.0, 1.0)
    c = build_trotter_circuit(terms, weights=w, time=1.0, steps=3, num_qubits=n)
    assert [tuple(o) for o in c.ops] == O.trotter_ops(terms, w, 1.0, 3), "trotter op list mismatch"
    out["trot8_state"] = np.asarray(eng.state(c))
    r = eng.run(c, shots=0)
    out["trot8_expz"] = np.array([r["expectations"][f"Z{q}"] for q in range(n)])

    # 4. kernels: k-qubit unitary, kraus, project
    n = 7
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi /= np.linalg.norm(psi)
    out["k7_psi"] = psi
    for k, qs in [(1, [4]), (2, [5, 1]), (3, [6, 0, 3]), (4, [2, 5, 1, 6])]:
        U = np.linalg.qr(rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k)))[0]
        out[f"k7_U{k}"] = U
        out[f"k7_q{k}"] = np.array(qs)
        out[f"k7_out{k}"] = np.asarray(RS.apply_kqubit_unitary(psi, U, qs, n))
    g = 0.3
    K0 = np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=complex)
    K1 = np.array([[0, np.sqrt(g)], [0, 0]], dtype=complex)
    out["k7_kraus_ops"] = np.stack([K0, K1])
    out["k7_kraus_status"] = np.array([0.05, 0.5, 0.97])
    out["k7_kraus_out"] = np.stack([np.asarray(RS.apply_kraus_statevector(psi, [K0, K1], 3, n, float(s))) for s in out["k7_kraus_status"]])
    out["k7_proj_out"] = np.stack([eng._project_z(psi.copy(), 2, 0, n), eng._project_z(psi.copy(), 5, 1, n)])

    # 5. sampling: the reference formula with explicit uniforms (Generator.choice internals)
    for tag in ["rand9", "rand12"]:
        st = out[f"{tag}_state"]
        p_np = np.abs(st) ** 2
        p_np = p_np / float(p_np.sum())
        u = np.random.default_rng(99).random(4096)
        idx_formula = O.sample_indices_numpy_formula(np.abs(st) ** 2, u)
        idx_choice = np.random.default_rng(99).choice(p_np.size, size=4096, p=p_np)
        assert np.array_equal(idx_formula, idx_choice)
        out[f"{tag}_sample_idx"] = idx_choice.astype(np.int64)
    meta["sample_seed"] = 99
    meta["sample_shots"] = 4096

    # 6. dense Pauli-sum expectation (pauli.py:74-87 + psi^dagger H psi)
    n = 6
    terms, w = O.heisenberg_terms(n, [(i, i + 1) for i in range(n - 1)], hzz=1.0, hxx=0.7, hyy=-0.4, hz=0.3, hx=-0.2, hy=0.1)
    H = np.asarray(pauli_string_sum_dense(terms, w))
    psi6 = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi6 /= np.linalg.norm(psi6)
    out["pauli6_terms"] = np.array(terms)
    out["pauli6_w"] = np.array(w)
    out["pauli6_psi"] = psi6
    out["pauli6_energy"] = np.array(np.real(np.vdot(psi6, H @ psi6)))
    out["pauli6_hpsi"] = H @ psi6

    np.savez_compressed(HERE / "reference_vectors.npz", **out)
    (HERE / "reference_circuits.json").write_text(json.dumps({"circuits": circuits, "meta": meta}))
    print("wrote", HERE / "reference_vectors.npz", (HERE / "reference_vectors.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
