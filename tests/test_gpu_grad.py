"""Gradients on the device: torch.autograd.Function (adjoint sweep) and the UCC energy+gradient path.

Energies within 1e-10 of the oracle; UCC energy + gradient within 1e-8 (BASELINE.json north_star)."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import sv_oracle as O
from oracle import ucc_oracle as U
from tests.conftest import FakeCircuit

pytestmark = pytest.mark.gpu


def _torch_circuit(n, thetas):
    """rx/ry/rz/rxx/ryy/rzz/cx circuit with torch angles (the gate set of the reference's
    tests_core_module/test_all_gates_gradient.py:16-127 and test_circuit_state_autograd.py)."""
    t = thetas
    ops = [("h", 0), ("rx", 0, t[0]), ("ry", 1, t[1]), ("rz", 2, t[2]), ("cx", 0, 1), ("rxx", 1, 2, t[3]),
           ("ryy", 0, 2, t[4]), ("rzz", 0, 1, t[5]), ("cz", 1, 2), ("rx", 2, t[0] * 2.0), ("s", 1), ("swap", 0, 2)]
    return FakeCircuit(n, ops)


def _loss_np(theta):
    t = [float(x) for x in theta]
    ops = [("h", 0), ("rx", 0, t[0]), ("ry", 1, t[1]), ("rz", 2, t[2]), ("cx", 0, 1), ("rxx", 1, 2, t[3]),
           ("ryy", 0, 2, t[4]), ("rzz", 0, 1, t[5]), ("cz", 1, 2), ("rx", 2, t[0] * 2.0), ("s", 1), ("swap", 0, 2)]
    psi, _ = O.evolve_ops(3, ops)
    w = np.arange(1, 9, dtype=np.float64)
    return float(np.sum(w * np.abs(psi) ** 2) + np.real(psi[1] * np.conj(psi[6])))


@pytest.mark.parametrize("pdtype", ["float64", "float32"])
def test_state_autograd_matches_finite_differences(cuda_device, pdtype):
    import torch
    from tyxonq_b200 import StatevectorEngine
    theta0 = np.array([0.3, -0.7, 1.1, 0.5, -0.2, 0.9])
    td = torch.float64 if pdtype == "float64" else torch.float32
    theta = torch.tensor(theta0, dtype=td, requires_grad=True)
    eng = StatevectorEngine("pytorch", device=cuda_device)
    psi = eng.state(_torch_circuit(3, theta))
    assert isinstance(psi, torch.Tensor) and psi.dtype == torch.complex128 and not psi.is_cuda and psi.requires_grad
    w = torch.arange(1, 9, dtype=torch.float64)
    loss = torch.sum(w * psi.abs() ** 2) + torch.real(psi[1] * torch.conj(psi[6]))
    loss.backward()
    th_used = theta.detach().double().numpy()
    assert abs(float(loss) - _loss_np(th_used)) < 1e-10
    fd = O.central_fd_gradient(_loss_np, th_used, 1e-6)
    tol = 1e-7 if pdtype == "float64" else 1e-5
    assert theta.grad.dtype == td
    assert np.abs(theta.grad.double().numpy() - fd).max() < tol


def test_state_autograd_kat(cuda_device):
    # test_circuit_state_autograd.py: <Z> = cos(0.5) after rx(0.5); d<Z>/dtheta = -sin(theta)
    import torch
    from tyxonq_b200 import StatevectorEngine
    th = torch.tensor(0.5, dtype=torch.float64, requires_grad=True)
    psi = StatevectorEngine("pytorch", device=cuda_device).state(FakeCircuit(1, [("rx", 0, th)]))
    z = psi[0].abs() ** 2 - psi[1].abs() ** 2
    z.backward()
    assert abs(float(z) - np.cos(0.5)) < 1e-12 and abs(float(th.grad) + np.sin(0.5)) < 1e-10


def test_state_autograd_layered_backward(cuda_device, monkeypatch):
    """torch backward of Circuit.state at n = 13: the layer-by-layer sweep (autograd.layered_sweep, default from 12 qubits
    on) gives the gradients of the per-gate sweep, and both agree with central differences of the oracle."""
    import torch
    from tyxonq_b200 import StatevectorEngine
    from tyxonq_b200 import autograd as A
    n = 13
    rng = np.random.default_rng(5)
    theta0 = rng.uniform(-1, 1, 8)
    w = rng.normal(size=1 << n)
    cvec = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)

    def ops_of(t):
        ops = [("h", q) for q in range(n)]
        ops += [("rz", q, t[q % 4]) for q in range(n)] + [("rx", q, t[4 + q % 3] * 0.5) for q in range(n)]
        ops += [("cx", q, q + 1) for q in range(n - 1)] + [("rzz", 0, n - 1, t[7])] + [("ry", q, t[(q + 1) % 8]) for q in range(0, n, 2)]
        ops += [("s", 3), ("rz", 3, t[2]), ("rxx", 5, 11, t[6] * -1.5), ("x", 6), ("rx", 6, t[0])]
        return ops

    def loss_np(t):
        psi, _ = O.evolve_ops(n, ops_of([float(x) for x in t]))
        return float(np.sum(w * np.abs(psi) ** 2) + np.real(np.vdot(cvec, psi)))

    grads = {}
    for name, thr in (("layered", 12), ("per_gate", 64)):
        monkeypatch.setattr(A, "LAYERED_MIN_QUBITS", thr)
        theta = torch.tensor(theta0, dtype=torch.float64, requires_grad=True)
        psi = StatevectorEngine("pytorch", device=cuda_device).state(FakeCircuit(n, ops_of(theta)))
        loss = torch.sum(torch.from_numpy(w) * psi.abs() ** 2) + torch.real(torch.sum(torch.from_numpy(cvec).conj() * psi))
        loss.backward()
        assert abs(float(loss) - loss_np(theta0)) < 1e-9
        grads[name] = theta.grad.numpy().copy()
    assert np.abs(grads["layered"] - grads["per_gate"]).max() < 1e-10
    fd = O.central_fd_gradient(loss_np, theta0, 1e-6)
    assert np.abs(grads["layered"] - fd).max() < 1e-6 * max(1.0, np.abs(fd).max())


def _ucc_problem(nao, ne):
    from tyxonq_b200 import ucc
    n = 2 * nao
    no, nv = ne // 2, nao - ne // 2
    ex_ops, param_ids = ucc.uccsd_ex_ops(no, nv)
    assert (ex_ops, param_ids) == U.uccsd_ex_ops(no, nv)
    i1, i2 = ucc.random_integral(nao, 2077)
    o1, o2 = U.random_integral(nao, 2077)
    assert np.array_equal(i1, o1) and np.array_equal(i2, o2)
    return n, (ne // 2, ne // 2), ex_ops, param_ids, i1, i2


def test_ucc_small_energy_grad(cuda_device):
    import torch
    from tyxonq_b200 import ucc
    n, nes, ex_ops, pids, i1, i2 = _ucc_problem(4, 4)
    ham = ucc.hamiltonian_from_integral(i1, i2)
    Hs = U.hamiltonian_from_integral(i1, i2)
    # the Pauli-sum H equals the fermionic sparse H on a random vector
    rng = np.random.default_rng(0)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    assert np.abs(ham.apply(torch.from_numpy(v).to(cuda_device)).cpu().numpy() - Hs @ v).max() < 1e-10
    np.random.seed(2077)
    params = np.random.rand(max(pids) + 1) - 0.5
    sv = ucc.UCCStatevector(n, nes, ex_ops, pids, ham, device=cuda_device)
    psi = sv.statevector(params).cpu().numpy()
    assert np.abs(psi - U.get_statevector(params, n, nes, ex_ops, pids)).max() < 1e-10
    e_ref, g_ref = U.energy_and_grad_adjoint(params, Hs, n, nes, ex_ops, pids)
    assert abs(sv.energy(params) - e_ref) < 1e-10
    e, g = sv.energy_and_grad(params)
    assert abs(e - e_ref) < 1e-10 and np.abs(g - g_ref).max() < 1e-8
    fd = O.central_fd_gradient(lambda p: U.energy(p, Hs, n, nes, ex_ops, pids), params, 1e-6)
    assert np.abs(g - fd).max() < 1e-6
    # the per-gate path (fused passes + tqb_grad_pair; what states beyond the L2-resident regime use) agrees
    sv2 = ucc.UCCStatevector(n, nes, ex_ops, pids, ham, device=cuda_device)
    assert sv.use_sweep
    sv2.use_sweep = False
    for graph in (False, True, True):
        e2, g2 = sv2.energy_and_grad(params, graph=graph)
        assert abs(e2 - e_ref) < 1e-10 and np.abs(g2 - g_ref).max() < 1e-8
    # complex64 state: 1e-5 tolerance
    import torch as _t
    sv3 = ucc.UCCStatevector(n, nes, ex_ops, pids, ham, device=cuda_device, dtype=_t.complex64)
    e3, g3 = sv3.energy_and_grad(params)
    assert abs(e3 - e_ref) < 1e-4 and np.abs(g3 - g_ref).max() < 1e-3


def test_ucc_h2o_shape_energy_grad(cuda_device):
    """Config 2 of BASELINE.json: 14 qubits, (5,5) electrons, 140 excitations / 75 parameters,
    synthetic integrals random_integral(7, 2077) (PySCF is not installable: parity with the real
    molecule is unpinned)."""
    from tyxonq_b200 import ucc
    n, nes, ex_ops, pids, i1, i2 = _ucc_problem(7, 10)
    assert n == 14 and len(ex_ops) == 140 and max(pids) + 1 == 75
    ham = ucc.hamiltonian_from_integral(i1, i2)
    Hs = U.hamiltonian_from_integral(i1, i2)
    np.random.seed(2077)
    params = np.random.rand(75) - 0.5
    sv = ucc.UCCStatevector(n, nes, ex_ops, pids, ham, device=cuda_device)
    e_ref, g_ref = U.energy_and_grad_adjoint(params, Hs, n, nes, ex_ops, pids)
    e, g = sv.energy_and_grad(params)
    assert abs(e - e_ref) < 1e-8
    assert np.abs(g - g_ref).max() < 1e-8
    # second call with other parameters reuses the compiled programs
    e2, g2 = sv.energy_and_grad(params * 0.5)
    e2_ref, g2_ref = U.energy_and_grad_adjoint(params * 0.5, Hs, n, nes, ex_ops, pids)
    assert abs(e2 - e2_ref) < 1e-8 and np.abs(g2 - g2_ref).max() < 1e-8
    # third call replays the captured CUDA graph with new matrices; the un-graphed path must agree too
    e3, g3 = sv.energy_and_grad(params * -0.3)
    e3_ref, g3_ref = U.energy_and_grad_adjoint(params * -0.3, Hs, n, nes, ex_ops, pids)
    assert sv._graph is not None
    assert abs(e3 - e3_ref) < 1e-8 and np.abs(g3 - g3_ref).max() < 1e-8
    e4, g4 = sv.energy_and_grad(params * -0.3, graph=False)
    assert abs(e4 - e3) < 1e-12 and np.abs(g4 - g3).max() < 1e-10


def test_ucc_batched_replicas(cuda_device):
    """energy_and_grad_batch: concurrent replicas (own buffers, stream, CUDA graph each) give the single-evaluation numbers."""
    from tyxonq_b200 import ucc
    n, nes, ex_ops, pids, i1, i2 = _ucc_problem(4, 4)
    sv = ucc.UCCStatevector(n, nes, ex_ops, pids, ucc.hamiltonian_from_integral(i1, i2), device=cuda_device)
    rng = np.random.default_rng(3)
    P_ = rng.uniform(-0.5, 0.5, (11, max(pids) + 1))
    es, gs = sv.energy_and_grad_batch(P_, replicas=4)
    for b in (0, 3, 10):
        e, g = sv.energy_and_grad(P_[b])
        assert abs(es[b] - e) < 1e-12 and np.abs(gs[b] - g).max() < 1e-12
    # the H2O-shaped problem, every member checked, for several replica counts (the energy is a two-stage reduction through
    # the library's workspace: concurrent replicas must reduce into their own slices, tqb_workspace_slot)
    i1, i2 = ucc.random_integral(7, 2077)
    ex_ops, pids = ucc.uccsd_ex_ops(5, 2)
    sv = ucc.UCCStatevector(14, (5, 5), ex_ops, pids, ucc.hamiltonian_from_integral(i1, i2), device=cuda_device)
    P_ = rng.uniform(-0.5, 0.5, (48, 75))
    ref = [sv.energy_and_grad(P_[b]) for b in range(48)]
    for R in (4, 12, 16):
        for _ in range(2):
            es, gs = sv.energy_and_grad_batch(P_, replicas=R)
            assert max(abs(es[b] - ref[b][0]) for b in range(48)) < 1e-12, R
            assert max(np.abs(gs[b] - ref[b][1]).max() for b in range(48)) < 1e-12, R


def test_b200_backend_value_and_grad(cuda_device):
    """Seam B3: tq.set_backend(B200Backend()) style use -- arrays live on the device, value_and_grad runs the
    adjoint sweep (reference pytorch_backend.py:446-564 returns (float, ndarray))."""
    import torch
    from tyxonq_b200 import B200Backend, StatevectorEngine
    K = B200Backend(cuda_device)
    w = K.array(np.arange(1, 9, dtype=np.float64))
    assert w.is_cuda and K.to_numpy(K.kron(K.eye(2), K.ones((2, 2)))).shape == (4, 4)
    eng = StatevectorEngine(K, device=cuda_device)
    assert eng.backend_name == "b200"

    def loss(theta):
        psi = eng.state(_torch_circuit(3, theta))
        assert psi.is_cuda
        return K.sum(w * K.abs(psi) ** 2) + K.real(psi[1] * K.conj(psi[6]))

    theta0 = np.array([0.3, -0.7, 1.1, 0.5, -0.2, 0.9])
    val, grad = K.value_and_grad(loss)(theta0)
    assert isinstance(val, float) and isinstance(grad, np.ndarray) and grad.shape == (6,)
    assert abs(val - _loss_np(theta0)) < 1e-10
    assert np.abs(grad - O.central_fd_gradient(_loss_np, theta0, 1e-6)).max() < 1e-7
    # sampling helper: numpy Generator.choice with explicit uniforms
    p = np.abs(O.evolve_ops(3, [("h", 0), ("rx", 1, 0.4), ("cx", 0, 2)])[0]) ** 2
    got = K.to_numpy(K.choice(np.random.default_rng(5), 8, size=64, p=K.array(p)))
    want = np.random.default_rng(5).choice(8, size=64, p=p / p.sum())
    assert np.array_equal(got, want)


def test_ucc_pair_gates_match_reference_fixture(cuda_device):
    """The device form of exp(theta G) -- ONE PAIR gate per excitation (ucc.excitation_gate) -- against the fixture
    generated from the reference's own constants and apply_kqubit_unitary (tests/golden/make_golden_ucc.py; mode
    'qubit': statevector_ops.py:42-43, 140-168).  With tests/test_ucc_host.py this pins the UCC evolution to reference
    output in everything but the Jordan-Wigner sign vector (openfermion is not installable here)."""
    import torch
    from pathlib import Path
    from tyxonq_b200 import program as P
    from tyxonq_b200 import ucc
    d = np.load(Path(__file__).resolve().parent / "golden" / "reference_ucc.npz")
    n, psi0 = int(d["n"]), d["psi0"]
    ex = [tuple(int(x) for x in r if x >= 0) for r in d["ex_ops"]]
    for k, (f, t) in enumerate(zip(ex, d["thetas"])):
        st = torch.from_numpy(psi0.copy()).to(cuda_device)
        P.apply_gates(st, [ucc.excitation_gate(f, float(t), n, mode="qubit")])
        assert np.abs(st.cpu().numpy() - d["each"][k]).max() < 1e-12, f
    st = torch.from_numpy(psi0.copy()).to(cuda_device)
    P.apply_gates(st, [ucc.excitation_gate(f, float(t), n, mode="qubit") for f, t in zip(ex, d["thetas"])])
    assert np.abs(st.cpu().numpy() - d["sequence"]).max() < 1e-12


def test_resident_vqe_kernel(cuda_device):
    """The CTA-resident evaluation (csrc/tqb_vqe.cu: forward, H|psi>, energy, reverse sweep in one CTA per parameter vector)
    against the oracle's energy and central differences, against the fused-pass / CUDA-graph path, and as a batch --
    the TFIM ansatz of examples/vqetfim_benchmark.py and a template with fixed gates, fixed angles, scaled and shared
    parameters."""
    from tyxonq_b200.pauli import PauliSum
    from tyxonq_b200.vqe import AdjointEnergy, Param, ResidentVQE, TFIMVqe
    rng = np.random.default_rng(12)
    # (1) TFIM-10, one layer
    v = TFIMVqe(10, 1, device=cuda_device)
    assert v._resident is not None
    p = rng.normal(size=(2, 10))
    e, g = v.energy_and_grad(p)
    e2, g2 = v.energy_and_grad(p, resident=False)
    assert abs(e - e2) < 1e-10 and np.abs(g - g2).max() < 1e-10
    B = 37
    pb = rng.normal(size=(B, 20))
    eb, gb = v.energy_and_grad_batch(pb)
    for b in (0, 5, B - 1):
        e1, g1 = v.energy_and_grad(pb[b].reshape(2, 10), resident=False)
        assert abs(eb[b] - e1) < 1e-10 and np.abs(gb[b] - g1.reshape(-1)).max() < 1e-10
    # (2) a template with fixed gates / angles, scaled and shared parameters, Y strings in H
    n = 6
    tmpl = [("h", q) for q in range(n)] + [("cx", 0, 1), ("ry", 1, Param(0)), ("rzz", 1, 2, Param(1, 2.0)), ("cz", 2, 3), ("rx", 3, 0.3),
            ("ryy", 3, 4, Param(0, -1.0)), ("s", 4), ("swap", 4, 5), ("rz", 5, Param(2)), ("rxx", 0, 5, Param(3)), ("sdg", 0),
            ("rx", 2, Param(4)), ("x", 3), ("cx", 5, 2)]
    ham_list = [(0.7, [("Z", 0), ("Z", 1)]), (-0.4, [("X", 1), ("Y", 3)]), (0.25, [("Y", 2)]), (0.6, [("X", 0), ("X", 5)]), (0.1, []),
                (-0.3, [("Z", 4)]), (0.45, [("Y", 1), ("Y", 4), ("Z", 5)])]
    ham = PauliSum.from_pauli_list(n, ham_list)
    assert ResidentVQE.supports(n, tmpl) and not ResidentVQE.supports(n, tmpl + [("t", 0)])
    ae = AdjointEnergy(n, tmpl, ham, device=cuda_device)
    th = rng.uniform(-1, 1, 5)

    def build(t):
        out = []
        for op in tmpl:
            out.append(tuple(a.scale * t[a.index] if isinstance(a, Param) else a for a in op))
        return out

    codes = {"X": 1, "Y": 2, "Z": 3}
    terms = []
    for c, ops in ham_list:
        ps = [0] * n
        for pch, q in ops:
            ps[q] = codes[pch]
        terms.append(ps)
    w = [c for c, _ in ham_list]

    def energy(t):
        psi, _ = O.evolve_ops(n, build(t), mode="state")
        return O.expect_pauli_sum(psi, terms, w)

    e, g = ae.energy_and_grad(th)
    assert abs(e - energy(th)) < 1e-10
    assert np.abs(g - O.central_fd_gradient(energy, th, 1e-6)).max() < 1e-7
    e2, g2 = ae.energy_and_grad(th, resident=False)
    assert abs(e - e2) < 1e-10 and np.abs(g - g2).max() < 1e-9


def test_layered_adjoint_matches_per_gate_path(cuda_device):
    """energy_and_grad_layered (tqb_transition_1q: all single-qubit gradients of a layer from one pair of states, fused
    un-apply passes between layers) against the per-gate adjoint path and central differences of the oracle: the
    hardware-efficient ansatz (rz.rx runs + cx ladder), the RY ansatz, a circuit with parametrised two-qubit gates and
    scaled / shared parameters; complex128 at n = 14 (one tile) and n = 16 (streaming tiles), complex64."""
    import torch
    from tyxonq_b200.pauli import PauliSum
    from tyxonq_b200.vqe import AdjointEnergy, Param
    rng = np.random.default_rng(21)

    def hea_template(n, layers):
        ops = [("h", q) for q in range(n)]
        k = 0
        for _ in range(layers):
            ops += [("cx", q, q + 1) for q in range(n - 1)]
            for q in range(n):
                ops.append(("rz", q, Param(k))); k += 1
                ops.append(("rx", q, Param(k))); k += 1
        return ops, k

    for n, dt, tol in ((14, torch.complex128, 1e-9), (16, torch.complex128, 1e-9), (15, torch.complex64, 2e-4)):
        tmpl, npar = hea_template(n, 2)
        ham = PauliSum.from_pauli_list(n, [(0.8, [("Z", i), ("Z", i + 1)]) for i in range(n - 1)] + [(-0.6, [("X", i)]) for i in range(n)]
                                       + [(0.3, [("Y", 0), ("Y", n - 1)])])
        ae = AdjointEnergy(n, tmpl, ham, device=cuda_device, dtype=dt)
        th = rng.uniform(-1.5, 1.5, npar)
        e1, g1 = ae.energy_and_grad_layered(th)
        e2, g2 = ae.energy_and_grad(th, resident=False) if n < AdjointEnergy.LAYERED_MIN_QUBITS else (e1, g1)
        assert abs(e1 - e2) < tol * 10 and np.abs(g1 - g2).max() < tol * 10, (n, dt, abs(e1 - e2), np.abs(g1 - g2).max())
        assert np.abs(g1).max() > 1e-3
    # two-qubit parametrised gates, scaled and shared parameters, fixed gates in between
    n = 12
    tmpl = [("h", q) for q in range(n)] + [("rzz", q, q + 1, Param(0, 2.0)) for q in range(0, n - 1, 2)] + [("rx", q, Param(1)) for q in range(n)] + \
           [("rxx", 1, 2, Param(2)), ("cx", 3, 4), ("ry", 4, Param(3)), ("rz", 4, Param(4, -1.0)), ("s", 5), ("rx", 5, Param(4)), ("ryy", 6, 9, Param(5))]
    ham = PauliSum.from_pauli_list(n, [(1.0, [("Z", i), ("Z", (i + 3) % n)]) for i in range(n)] + [(0.5, [("X", 4)]), (0.25, [("Y", 5), ("Z", 6)])])
    ae = AdjointEnergy(n, tmpl, ham, device=cuda_device)
    th = rng.uniform(-1, 1, 6)
    e1, g1 = ae.energy_and_grad_layered(th)
    e2, g2 = ae.energy_and_grad(th, resident=False)
    assert abs(e1 - e2) < 1e-10 and np.abs(g1 - g2).max() < 1e-9
