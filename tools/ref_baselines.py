"""CPU baselines for bench.py: the REAL reference (imported from baseline/_ref, the pip-installed unmodified package)
timed on the box's host cores, beside every GPU number.  Nothing here is on the product path.

Each function returns a ``cpu_baseline`` dict: {"value", "unit", "cores", "kind": "reference" | "port", "sample"}.
kind "reference" = the reference's own functions through its public API; kind "port" = the oracle's restatement
(used only where the reference cannot run: n >= 23, or the chem stack that is not installable here).
"""
from __future__ import annotations

import os
import sys
import time
import warnings
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "baseline" / "_ref"


def reference_available() -> bool:
    return (REF / "tyxonq" / "__init__.py").exists()


def import_reference():
    if not reference_available():
        raise ImportError("baseline/_ref is not present")
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    warnings.filterwarnings("ignore")
    import tyxonq
    return tyxonq


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def hea_gates_per_s_reference(n: int = 20, layers: int = 1, reps: int = 2) -> dict:
    """The reference's own engine on the bench circuit at the largest size it can run in seconds (its 2-qubit einsum
    stops at n = 22): Circuit(n, ops).state() through StatevectorEngine (engine.py:914-1038), numpy
    backend.  value = gates * 2^(n-30) / s, the bench's normalised unit (a gate sweep costs 2^n)."""
    tq = import_reference()
    tq.set_backend("numpy")
    from tyxonq.devices.simulators.statevector.engine import StatevectorEngine
    sys.path.insert(0, str(ROOT))
    from tyxonq_b200.circuits import hea_ops
    ops = hea_ops(n, layers, np.random.default_rng(1234).uniform(-np.pi, np.pi, 2 * layers * n))
    c = tq.Circuit(n, ops=list(ops))
    eng = StatevectorEngine()
    eng.state(c)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        eng.state(c)
        ts.append(time.perf_counter() - t0)
    dt = float(np.mean(ts))
    return {"value": len(ops) * 2.0 ** (n - 30) / dt, "unit": "gates/s", "cores": 1, "kind": "reference",
            "raw_gates_per_s_at_n": len(ops) / dt, "n": n,
            "sample": f"reference StatevectorEngine.state(Circuit({n}, hea depth {layers}: {len(ops)} gates)), numpy einsum "
                      f"(single-threaded), {reps} repetitions of {dt:.2f} s; normalised by 2^({n}-30)"}


def tfim10_energy_grad_reference(reps: int = 5) -> dict:
    """examples/vqetfim_benchmark.py:70-103 (exact_energy: the reference kernels on the pytorch backend) + torch
    autograd for the gradient, n = 10, one layer -- restated call for call, on the reference's own functions."""
    import torch
    tq = import_reference()
    tq.set_backend("pytorch")
    try:
        nb = tq.get_backend("pytorch")
        from tyxonq.libs.quantum_library.kernels.gates import gate_h, gate_rxx, gate_rz
        from tyxonq.libs.quantum_library.kernels.statevector import (apply_1q_statevector, apply_2q_statevector,
                                                                     expect_z_statevector, init_statevector)
        n, nlayers, Jx, h = 10, 1, 1.0, -1.0
        dim = 1 << n
        signs = [torch.as_tensor([1.0 if (((k >> (n - 1 - i)) & 1) == ((k >> (n - 2 - i)) & 1)) else -1.0 for k in range(dim)],
                                 dtype=torch.float64) for i in range(n - 1)]

        def exact_energy(param):
            psi = init_statevector(n, backend=nb)
            t = 0
            for _ in range(nlayers):
                for i in range(n - 1):
                    psi = apply_2q_statevector(nb, psi, gate_rxx(param[t, i]), i, i + 1, n)
                t += 1
                for i in range(n):
                    psi = apply_1q_statevector(nb, psi, gate_rz(param[t, i]), i, n)
                t += 1
            e = torch.zeros((), dtype=torch.float64)
            for i in range(n):
                e = e + h * expect_z_statevector(psi, i, n, backend=nb)
            psi_x = psi
            for q in range(n):
                psi_x = apply_1q_statevector(nb, psi_x, gate_h(), q, n)
            probs = nb.abs(psi_x) ** 2
            for i in range(n - 1):
                e = e + Jx * torch.sum(signs[i] * probs)   # (the example rebuilds the sign list per term: not timed here)
            return e

        p0 = np.random.default_rng(0).normal(size=(2 * nlayers, n))
        ts = []
        e = None
        for r in range(reps + 1):
            p = torch.tensor(p0, dtype=torch.float64, requires_grad=True)
            t0 = time.perf_counter()
            e = exact_energy(p)
            e.backward()
            if r:
                ts.append(time.perf_counter() - t0)
        dt = float(np.mean(ts))
        return {"value": 1.0 / dt, "unit": "energy+gradient evaluations/s", "cores": torch.get_num_threads(), "kind": "reference",
                "energy": float(e), "sample": f"exact_energy + torch autograd backward (examples/vqetfim_benchmark.py:70-103), "
                                              f"n = 10, 1 layer, {reps} repetitions, {1e3 * dt:.1f} ms each"}
    finally:
        tq.set_backend("numpy")


def ucc_h2o_energy_grad_port(reps: int = 3) -> dict:
    """The reference's chem stack (pyscf / openfermion) is not installable here; baseline = the oracle's numpy
    restatement of energy_and_grad_statevector (statevector_ops.py:203-244) on the same synthetic 14-qubit problem."""
    sys.path.insert(0, str(ROOT))
    from oracle import ucc_oracle as U
    i1, i2 = U.random_integral(7, 2077)
    H = U.hamiltonian_from_integral(i1, i2)
    ex_ops, pids = U.uccsd_ex_ops(5, 2)
    np.random.seed(2077)
    p = np.random.rand(max(pids) + 1) - 0.5
    U.energy_and_grad_adjoint(p, H, 14, (5, 5), ex_ops, pids)
    ts = []
    e = None
    for _ in range(reps):
        t0 = time.perf_counter()
        e, g = U.energy_and_grad_adjoint(p, H, 14, (5, 5), ex_ops, pids)
        ts.append(time.perf_counter() - t0)
    dt = float(np.mean(ts))
    return {"value": 1.0 / dt, "unit": "energy+gradient evaluations/s", "cores": 1, "kind": "port", "energy": float(e),
            "sample": f"oracle/ucc_oracle.energy_and_grad_adjoint (numpy + scipy.sparse H), 14 qubits, {len(ex_ops)} excitations, "
                      f"{reps} repetitions, {1e3 * dt:.0f} ms each"}


def batched_hwe20_reference(n: int = 20, layers: int = 4, shots: int = 8192, sets: int = 1) -> dict:
    """Config 5 with the reference: Circuit(n, build_hwe_ry_ops).state() (numpy engine), the Heisenberg energy term by
    term with the reference kernels on that state (circuit.py:1751-1796 does this per term, re-simulating each time:
    only ONE simulation is timed here), and Generator.choice + bincount for the shots (engine.py:377-418)."""
    tq = import_reference()
    tq.set_backend("numpy")
    nb = tq.get_backend("numpy")
    from tyxonq.devices.simulators.statevector.engine import StatevectorEngine
    from tyxonq.libs.quantum_library.kernels.gates import gate_x, gate_y, gate_z
    from tyxonq.libs.quantum_library.kernels.statevector import apply_1q_statevector
    sys.path.insert(0, str(ROOT))
    from tyxonq_b200.circuits import hwe_ry_ops
    eng = StatevectorEngine()
    t_state = t_ev = t_s = 0.0
    e = 0.0
    rng = np.random.default_rng(7)
    for _ in range(sets):
        ops = hwe_ry_ops(n, layers, rng.random((layers + 1) * n))
        c = tq.Circuit(n, ops=list(ops))
        t0 = time.perf_counter()
        psi = np.asarray(eng.state(c))
        t_state += time.perf_counter() - t0
        t0 = time.perf_counter()
        e = 0.0
        for i in range(n - 1):
            for g in (gate_z, gate_x, gate_y):
                phi = apply_1q_statevector(nb, psi, g(), i, n)
                phi = apply_1q_statevector(nb, phi, g(), i + 1, n)
                e += float(np.real(np.vdot(psi, phi)))
        t_ev += time.perf_counter() - t0
        t0 = time.perf_counter()
        p = np.abs(psi) ** 2
        p = p / p.sum()
        idx = np.random.default_rng(1).choice(len(p), size=shots, p=p)
        np.bincount(idx, minlength=len(p))
        t_s += time.perf_counter() - t0
    return {"states_per_s": {"value": sets / t_state, "unit": "states/s"},
            "expvals_per_s": {"value": sets / t_ev, "unit": "expectation values/s"},
            "shots_per_s": {"value": sets * shots / t_s, "unit": "shots/s"},
            "cores": 1, "kind": "reference", "energy": e,
            "sample": f"{sets} parameter set(s) of the {n}-qubit HWE-RY ansatz ({layers} layers): reference engine.state "
                      f"{t_state / sets:.1f} s, 57 Pauli terms with the reference kernels {t_ev / sets:.1f} s, "
                      f"Generator.choice + bincount of {shots} shots {1e3 * t_s / sets:.0f} ms"}
