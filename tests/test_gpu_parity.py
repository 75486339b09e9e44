"""Parity of the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Tolerances (BASELINE.json north_star): amplitudes / energies within 1e-10 in complex128,
1e-5 in complex64; sampled indices bit-exact for the same host-supplied uniforms."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import sv_oracle as O
from tests.conftest import FakeCircuit, random_ops

pytestmark = pytest.mark.gpu

TOL128 = 1e-10
TOL64 = 1e-5


def _engine(dev, backend="numpy", dtype=None, tile=None):
    import torch
    from tyxonq_b200 import StatevectorEngine
    return StatevectorEngine(backend, device=dev, dtype=dtype or torch.complex128, tile=tile)


@pytest.mark.parametrize("tag", ["rand5", "rand9", "rand12"])
def test_golden_states_and_expectations(cuda_device, golden, tag):
    c = golden["circuits"][tag]
    eng = _engine(cuda_device)
    psi = eng.state(FakeCircuit(c["n"], c["ops"]))
    assert psi.dtype == np.complex128
    assert np.abs(psi - golden[f"{tag}_state"]).max() < TOL128
    ops_m = list(c["ops"]) + [("measure_z", q) for q in range(c["n"])]
    r = eng.run(FakeCircuit(c["n"], ops_m), shots=0)
    assert set(r.keys()) == {"expectations", "metadata"}
    got = np.array([r["expectations"][f"Z{q}"] for q in range(c["n"])])
    # NOTE run() executes cry, state() does not: the golden expz comes from engine.run
    assert np.abs(got - golden[f"{tag}_expz"]).max() < TOL128


def test_golden_builders(cuda_device, golden):
    eng = _engine(cuda_device)
    for name, n, ops in [
        ("hea12", 12, O.hea_ops(12, 3, golden["hea12_params"])),
        ("hwe10", 10, O.hwe_ry_ops(10, 4, golden["hwe10_params"])),
        ("qaoa10", 10, O.qaoa_ring_ops(10, 3, golden["qaoa10_params"])),
        ("trot8", 8, O.trotter_ops(*O.tfim_terms(8, 1.0, 1.0), 1.0, 3)),
    ]:
        psi = eng.state(FakeCircuit(n, ops))
        assert np.abs(psi - golden[f"{name}_state"]).max() < TOL128, name


@pytest.mark.parametrize("n,m,L", [(1, 1, 0), (2, 2, 1), (4, 3, 1), (9, 5, 2), (12, 8, 4), (14, 11, 5), (16, 11, 5), (18, 13, 5)])
@pytest.mark.parametrize("dtype", ["c128", "c64"])
def test_random_circuits_vs_oracle(cuda_device, n, m, L, dtype):
    import torch
    from tyxonq_b200.planner import TileConfig
    rng = np.random.default_rng(1000 + n)
    ops = random_ops(rng, n, 150)
    ref, _ = O.evolve_ops(n, ops, mode="run")
    td = torch.complex128 if dtype == "c128" else torch.complex64
    eng = _engine(cuda_device, "b200", td, TileConfig(m=m, L=L))
    psi, _, _ = eng._evolve(FakeCircuit(n, ops), "run")
    got = psi.cpu().numpy()
    assert np.abs(got - ref).max() < (TOL128 if dtype == "c128" else TOL64)
    assert eng.last_passes <= eng.last_gates


def test_default_tiles_streaming_regime(cuda_device):
    """n = 20..22: default streaming tile (m=11/12), HEA + QAOA generators, both dtypes."""
    import torch
    for n, ops in [(20, O.hea_ops(20, 3, np.random.default_rng(1).uniform(-3, 3, 120))),
                   (22, O.qaoa_ring_ops(22, 2, np.random.default_rng(2).uniform(-3, 3, 4)))]:
        ref, _ = O.evolve_ops(n, ops)
        for td, tol in ((torch.complex128, TOL128), (torch.complex64, TOL64)):
            eng = _engine(cuda_device, "b200", td)
            psi, _, _ = eng._evolve(FakeCircuit(n, ops), "state")
            assert np.abs(psi.cpu().numpy() - ref).max() < tol
            assert eng.last_passes < eng.last_gates / 2


def test_kernel_function_shims(cuda_device, golden):
    from tyxonq_b200 import kernels as K
    psi = golden["k7_psi"]
    for k in (1, 2, 3, 4):
        out = K.apply_kqubit_unitary(psi, golden[f"k7_U{k}"], golden[f"k7_q{k}"].tolist(), 7)
        assert isinstance(out, np.ndarray)
        assert np.abs(out - golden[f"k7_out{k}"]).max() < TOL128
    out = K.apply_1q_statevector(None, psi, O.gate_rx(0.3), 4, 7)
    assert np.abs(out - O.apply_1q(psi, O.gate_rx(0.3), 4, 7)).max() < TOL128
    out = K.apply_2q_statevector(None, psi, O.gate_rxx(0.7), 5, 2, 7)
    assert np.abs(out - O.apply_2q(psi, O.gate_rxx(0.7), 5, 2, 7)).max() < TOL128
    assert K.apply_2q_statevector(None, psi, O.gate_rxx(0.7), 3, 3, 7) is psi
    for q in range(7):
        assert abs(float(K.expect_z_statevector(psi, q, 7)) - O.expect_z(psi, q, 7)) < TOL128
    for s, ref in zip(golden["k7_kraus_status"], golden["k7_kraus_out"]):
        out = K.apply_kraus_statevector(psi, list(golden["k7_kraus_ops"]), 3, 7, float(s))
        assert np.abs(out - ref).max() < TOL128
    z = K.init_statevector(5)
    assert z.is_cuda and z.shape == (32,) and complex(z[0].cpu()) == 1.0 and float(z.abs().sum().cpu()) == 1.0


def test_project_reset_and_initial_state(cuda_device, golden):
    eng = _engine(cuda_device)
    psi = golden["k7_psi"]
    a = eng.state(FakeCircuit(7, [("project_z", 2, 0)], inputs=psi))
    assert np.abs(a - golden["k7_proj_out"][0]).max() < TOL128
    b = eng.state(FakeCircuit(7, [("project_z", 5, 1)], inputs=psi))
    assert np.abs(b - golden["k7_proj_out"][1]).max() < TOL128
    # run() ignores the initial state (engine.py:46); mid-circuit measurement KAT
    r = eng.run(FakeCircuit(2, [("h", 0), ("cx", 0, 1), ("project_z", 0, 0), ("reset", 1), ("measure_z", 0), ("measure_z", 1)],
                            inputs=np.array([0, 1, 0, 0], dtype=complex)), shots=0)
    assert abs(r["expectations"]["Z0"] - 1) < TOL128 and abs(r["expectations"]["Z1"] - 1) < TOL128


def test_unitary_ops_kat(cuda_device):
    eng = _engine(cuda_device)
    sx = 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]])
    assert np.allclose(eng.state(FakeCircuit(1, [("unitary", 0, "k")], unitary_cache={"k": sx})), [0.5 + 0.5j, 0.5 - 0.5j], atol=1e-12)
    isw = O.gate_iswap_4x4()
    assert np.allclose(eng.state(FakeCircuit(2, [("x", 0), ("unitary", 0, 1, "k")], unitary_cache={"k": isw})), [0, 1j, 0, 0], atol=1e-12)
    assert np.allclose(eng.state(FakeCircuit(2, [("h", 0), ("h", 1), ("unitary", 0, 1, "k")], unitary_cache={"k": isw})),
                       [0.5, 0.5j, 0.5j, 0.5], atol=1e-12)
    rng = np.random.default_rng(42)
    ops, cache = [], {}
    for i in range(5):
        cache[f"u{i}"] = np.linalg.qr(rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2)))[0]
        ops.append(("unitary", int(rng.integers(3)), f"u{i}"))
    psi = eng.state(FakeCircuit(3, ops, unitary_cache=cache))
    ref, _ = O.evolve_ops(3, ops, unitary_cache=cache)
    assert np.abs(psi - ref).max() < TOL128 and abs(np.linalg.norm(psi) - 1) < 1e-12


def test_probability_amplitude_perfect_sampling(cuda_device):
    eng = _engine(cuda_device)
    c = FakeCircuit(2, [("h", 0)])
    p = eng.probability(c)
    assert abs(p[0] + p[2] - 1.0) < 1e-12 and p[1] == 0 and p[3] == 0
    assert eng.amplitude(c, "01") == 0 and abs(eng.amplitude(c, "10") - 2 ** -0.5) < 1e-12
    with pytest.raises(ValueError):
        eng.amplitude(c, "0")
    bits, prob = eng.perfect_sampling(c, rng=np.random.default_rng(0))
    assert bits in ("00", "10") and abs(prob - 0.5) < 1e-12


@pytest.mark.parametrize("tag", ["rand9", "rand12"])
def test_sampling_bit_exact(cuda_device, golden, tag):
    """Same uniforms -> the same indices as the oracle's blocked CDF and (on these inputs) as numpy's
    Generator.choice run on the reference's probabilities (golden fixture)."""
    import torch
    from tyxonq_b200 import program as P
    st = golden[f"{tag}_state"]
    u = np.random.default_rng(golden["meta"]["sample_seed"]).random(golden["meta"]["sample_shots"])
    dev_state = torch.from_numpy(st).to(cuda_device)
    idx = P.sample(dev_state, torch.from_numpy(u)).cpu().numpy()
    assert np.array_equal(idx, O.sample_indices(O.probabilities(st), u))
    assert np.array_equal(idx, golden[f"{tag}_sample_idx"])
    # engine.run(shots) with host-supplied uniforms: counts over ALL n qubits, big-endian keys
    c = golden["circuits"][tag]
    eng = _engine(cuda_device)
    ops = [o for o in c["ops"]] + [("measure_z", 0)]
    psi_run, _ = O.evolve_ops(c["n"], ops, mode="run")
    r = eng.run(FakeCircuit(c["n"], ops), shots=len(u), uniforms=u)
    assert r["metadata"]["shots"] == len(u) and set(r.keys()) == {"result", "metadata"}
    assert r["result"] == O.counts_from_indices(O.sample_indices(O.probabilities(psi_run), u), c["n"])


def test_sampling_large_and_batched(cuda_device):
    """n = 18 (64 scan chunks) and a batch of states: bit-exact against the oracle, c128 and c64."""
    import torch
    from tyxonq_b200 import program as P
    rng = np.random.default_rng(7)
    n = 18
    st = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    st /= np.linalg.norm(st)
    u = rng.random(8192)
    idx = P.sample(torch.from_numpy(st).to(cuda_device), torch.from_numpy(u)).cpu().numpy()
    assert np.array_equal(idx, O.sample_indices(O.probabilities(st), u))
    st64 = st.astype(np.complex64)
    idx = P.sample(torch.from_numpy(st64).to(cuda_device), torch.from_numpy(u)).cpu().numpy()
    assert np.array_equal(idx, O.sample_indices(O.probabilities(st64), u))
    B, n2 = 5, 13
    sb = rng.normal(size=(B, 1 << n2)) + 1j * rng.normal(size=(B, 1 << n2))
    ub = rng.random((B, 1000))
    idx = P.sample(torch.from_numpy(sb).to(cuda_device), torch.from_numpy(ub)).cpu().numpy()
    for b in range(B):
        assert np.array_equal(idx[b], O.sample_indices(O.probabilities(sb[b]), ub[b]))


def test_reductions(cuda_device):
    import torch
    from tyxonq_b200 import program as P
    rng = np.random.default_rng(11)
    for n, B in [(3, 1), (9, 4), (15, 2), (20, 1)]:
        st = rng.normal(size=(B, 1 << n)) + 1j * rng.normal(size=(B, 1 << n))
        for dt, tol in ((np.complex128, 1e-10), (np.complex64, 1e-4)):
            s = st.astype(dt)
            d = torch.from_numpy(s).to(cuda_device)
            scale = float(np.sum(np.abs(s[0]) ** 2))
            nrm = P.norm2(d).cpu().numpy()
            assert np.abs(nrm - np.sum(np.abs(s.astype(np.complex128)) ** 2, axis=1)).max() < tol * scale
            z = P.expect_z_bits(d).cpu().numpy()
            for b in range(B):
                ref = np.array([O.expect_z(s[b].astype(np.complex128), n - 1 - p, n) for p in range(n)])
                assert np.abs(z[b] - ref).max() < tol * scale
            masks = [int(rng.integers(1, 1 << n)) for _ in range(37)]
            zm = P.expect_zmasks(d, masks).cpu().numpy()
            p = np.abs(s[0].astype(np.complex128)) ** 2
            i = np.arange(1 << n)
            for t, mk in enumerate(masks):
                par = np.array([bin(v).count("1") & 1 for v in (i & mk)]) if n <= 15 else None
                if par is not None:
                    assert abs(zm[0, t] - np.sum(p * (1 - 2 * par))) < tol * scale
            ip = P.inner(d, torch.flip(d, dims=[1]).contiguous()).cpu().numpy()
            assert np.abs(ip - np.sum(np.conj(s) * s[:, ::-1], axis=1)).max() < (tol * scale * 10)


def test_pauli_sum(cuda_device, golden):
    import torch
    from tyxonq_b200 import PauliSum
    terms, w = golden["pauli6_terms"].tolist(), golden["pauli6_w"].tolist()
    ham = PauliSum.from_codes(terms, w)
    psi = torch.from_numpy(golden["pauli6_psi"]).to(cuda_device)
    e = ham.expectation(psi).cpu().numpy()[0]
    assert abs(e.real - float(golden["pauli6_energy"])) < TOL128 and abs(e.imag) < TOL128
    assert np.abs(ham.apply(psi).cpu().numpy() - golden["pauli6_hpsi"]).max() < TOL128
    # larger, random weights incl. Y strings; list-of-(coeff, ops) constructor; engine.expval
    rng = np.random.default_rng(5)
    n = 11
    terms = [[int(c) for c in rng.integers(0, 4, n)] for _ in range(40)]
    w = rng.normal(size=40).tolist()
    st = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    st /= np.linalg.norm(st)
    ham = PauliSum.from_codes(terms, w)
    d = torch.from_numpy(st).to(cuda_device)
    assert abs(ham.expectation(d).cpu().numpy()[0].real - O.expect_pauli_sum(st, terms, w)) < TOL128
    assert np.abs(ham.apply(d).cpu().numpy() - O.apply_pauli_sum(st, terms, w)).max() < TOL128
    lst = [(w[t], [("IXYZ"[c], q) for q, c in enumerate(terms[t]) if c]) for t in range(40)]
    eng = _engine(cuda_device)
    assert abs(eng.expval(FakeCircuit(n, [], inputs=st), lst) - O.expect_pauli_sum(st, terms, w)) < TOL128


def test_pauli_sum_tiled_kernel(cuda_device):
    """Tile-staged Pauli sums (expect_pauli_tiled_kernel) against the oracle and against the gather kernel: several tile
    layouts (streaming 64 KiB tiles at n = 16 / 17), Hermitian and non-Hermitian (complex-weight) sums, X/Y strings spread
    over the register, batched states, complex64."""
    import torch
    from tyxonq_b200 import PauliSum
    rng = np.random.default_rng(23)
    for n, dt, tol, herm in ((16, torch.complex128, TOL128, True), (16, torch.complex128, TOL128, False),
                             (17, torch.complex64, 2e-5, True), (9, torch.complex128, TOL128, False)):
        terms = []
        for _ in range(60):
            t = [0] * n
            for q in rng.choice(n, size=int(rng.integers(1, 7)), replace=False):
                t[int(q)] = int(rng.integers(1, 4))
            terms.append(t)
        w = rng.normal(size=60) + (0 if herm else 1j * rng.normal(size=60))
        ham = PauliSum.from_codes(terms, w.tolist())
        assert ham.hermitian == herm
        B = 3
        st = rng.normal(size=(B, 1 << n)) + 1j * rng.normal(size=(B, 1 << n))
        st /= np.linalg.norm(st, axis=1, keepdims=True)
        d = torch.from_numpy(st).to(cuda_device).to(dt)
        plan = ham._tiled_plan(n, d.element_size(), d.device)
        assert plan is not None and (n < 14 or plan[1] >= 2), "expected several tile layouts"
        got = ham.expectation(d, tiled=True).cpu().numpy()
        old = ham.expectation(d, tiled=False).cpu().numpy()
        for b in range(B):
            ref = O.expect_pauli_sum(st[b], terms, w.tolist()) if herm else complex(np.vdot(st[b], O.apply_pauli_sum(st[b], terms, w.tolist())))
            assert abs(got[b] - ref) < tol, (n, dt, herm, b, got[b], ref)
            assert abs(old[b] - ref) < tol


def test_pauli_sum_tiled_many_terms_per_group(cuda_device):
    """Groups of many terms with one x-mask (sign-sum tables of the tiled kernel: low-bit table, high-bit table, straddling
    terms): Z-only strings of 1..6 random qubits plus a few X/Y patterns that each carry many Z decorations, real and
    complex weights, one-tile (n = 12) and streaming (n = 16, 17) layouts, complex64."""
    import torch
    from tyxonq_b200 import PauliSum
    rng = np.random.default_rng(29)
    for n, dt, tol, herm in ((16, torch.complex128, TOL128, True), (17, torch.complex64, 2e-5, True), (12, torch.complex128, TOL128, False),
                             (16, torch.complex128, TOL128, False)):
        terms = []
        for _ in range(40):     # diagonal group
            t = [0] * n
            for q in rng.choice(n, size=int(rng.integers(1, 7)), replace=False):
                t[int(q)] = 3
            terms.append(t)
        for xs in ((0, 1), (3, n - 1), (5,), (2, 7, 11)):   # four off-diagonal groups with 12 terms each
            for _ in range(12):
                t = [0] * n
                for q in rng.choice(n, size=int(rng.integers(0, 5)), replace=False):
                    t[int(q)] = 3
                for q in xs:
                    t[q] = int(rng.integers(1, 3)) if t[q] == 0 else 2
                terms.append(t)
        w = rng.normal(size=len(terms)) + (0 if herm else 1j * rng.normal(size=len(terms)))
        ham = PauliSum.from_codes(terms, w.tolist())
        B = 2
        st = rng.normal(size=(B, 1 << n)) + 1j * rng.normal(size=(B, 1 << n))
        st /= np.linalg.norm(st, axis=1, keepdims=True)
        d = torch.from_numpy(st).to(cuda_device).to(dt)
        got = ham.expectation(d, tiled=True).cpu().numpy()
        for b in range(B):
            ref = complex(np.vdot(st[b], O.apply_pauli_sum(st[b], terms, w.tolist())))
            assert abs(got[b] - ref) < tol * 4, (n, dt, herm, b, got[b], ref)


def test_pauli_sum_tiled_vanishing_half(cuda_device):
    """Two-term groups whose sign sum vanishes on half of the pairs (XX + YY and XX - YY bonds, hopping terms X Z..Z X +
    Y Z..Z Y, and pairs whose Z decorations reach OUTSIDE the tile so that the vanishing half changes from tile to tile),
    next to pairs with unequal weights that must take the general loop."""
    import torch
    from tyxonq_b200 import PauliSum
    rng = np.random.default_rng(31)
    for n, dt, tol in ((12, torch.complex128, TOL128), (17, torch.complex128, TOL128), (18, torch.complex64, 2e-5)):
        terms, w = [], []

        def add(ops, c):
            t = [0] * n
            for q, p in ops:
                t[q] = p
            terms.append(t)
            w.append(c)

        for i in range(n - 1):                       # Heisenberg-like bonds, equal and opposite weights
            c = float(rng.normal())
            add([(i, 1), (i + 1, 1)], c)
            add([(i, 2), (i + 1, 2)], c if i % 3 else -c)
        for i, j in ((0, 3), (2, n - 1), (5, 9), (1, n - 2)):   # hopping terms with a Z string in between
            c = float(rng.normal())
            zs = [(q, 3) for q in range(i + 1, j)]
            add([(i, 1), (j, 1)] + zs, c)
            add([(i, 2), (j, 2)] + zs, c)
        for i, j, k in ((0, 1, n - 1), (3, 4, n - 2), (6, 2, n - 1)):   # the two terms differ by a Z far away (outside most tiles)
            c = float(rng.normal())
            add([(i, 1), (j, 1), (k, 3)], c)
            add([(i, 2), (j, 2)], c)
        for i in range(0, n - 1, 2):                 # unequal weights: nothing vanishes
            add([(i, 1), (i + 1, 2)], float(rng.normal()))
            add([(i, 2), (i + 1, 1)], float(rng.normal()))
        ham = PauliSum.from_codes(terms, w)
        B = 2
        st = rng.normal(size=(B, 1 << n)) + 1j * rng.normal(size=(B, 1 << n))
        st /= np.linalg.norm(st, axis=1, keepdims=True)
        d = torch.from_numpy(st).to(cuda_device).to(dt)
        got = ham.expectation(d, tiled=True).cpu().numpy()
        old = ham.expectation(d, tiled=False).cpu().numpy()
        for b in range(B):
            ref = O.expect_pauli_sum(st[b], terms, w)
            assert abs(got[b] - ref) < tol * 4, (n, dt, b, got[b], ref)
            assert abs(old[b] - ref) < tol * 4


def test_tfim_vqe_energy(cuda_device):
    """examples/vqetfim_benchmark.py exact_energy on the device vs the oracle restatement."""
    from tyxonq_b200.vqe import TFIMVqe
    rng = np.random.default_rng(0)
    p = rng.normal(size=(2, 10))
    v = TFIMVqe(10, 1, device=cuda_device)
    assert abs(v.energy(p) - O.tfim_vqe_energy(10, 1, p)) < TOL128
    assert abs(v.energy(np.zeros((2, 10))) + 10.0) < TOL128
    e, g = v.energy_and_grad(p)
    fd = O.central_fd_gradient(lambda x: O.tfim_vqe_energy(10, 1, x.reshape(2, 10)), p.reshape(-1), 1e-6)
    assert abs(e - O.tfim_vqe_energy(10, 1, p)) < TOL128
    assert np.abs(g.reshape(-1) - fd).max() < 1e-7
    # (that call ran the CTA-resident kernel, csrc/tqb_vqe.cu); the fused-pass path: later calls replay one CUDA graph
    # with rewritten matrices
    assert v._resident is not None
    for scale in (1.0, 0.5, -1.3):
        q = p * scale
        e2, g2 = v.energy_and_grad(q, resident=False)
        assert abs(e2 - O.tfim_vqe_energy(10, 1, q)) < TOL128
        fd2 = O.central_fd_gradient(lambda x: O.tfim_vqe_energy(10, 1, x.reshape(2, 10)), q.reshape(-1), 1e-6)
        assert np.abs(g2.reshape(-1) - fd2).max() < 1e-7
    assert v._graph is not None
    # two layers, complex64 state
    import torch
    v2 = TFIMVqe(8, 2, device=cuda_device, dtype=torch.complex64)
    p2 = rng.normal(size=(4, 8))
    e3, g3 = v2.energy_and_grad(p2)
    assert abs(e3 - O.tfim_vqe_energy(8, 2, p2)) < 1e-4
    fd3 = O.central_fd_gradient(lambda x: O.tfim_vqe_energy(8, 2, x.reshape(4, 8)), p2.reshape(-1), 1e-6)
    assert np.abs(g3.reshape(-1) - fd3).max() < 1e-3


def test_lean_and_general_kernels_agree_with_oracle(cuda_device):
    """Every pass of HEA / hwe-ry / Trotter / QAOA circuits is lean-eligible (tile_pass_lean_kernel); the same plans must
    give the oracle's state through the general kernel too (lean switched off) and with chains limited to 3 layers.  Angles in (-pi, pi): both scaled-rotation modes occur."""
    import torch
    from tyxonq_b200 import _lib
    from tyxonq_b200.planner import TileConfig
    lib = _lib.load()
    rng = np.random.default_rng(77)
    cases = [
        (20, O.hea_ops(20, 8, rng.uniform(-np.pi, np.pi, 2 * 8 * 20))),
        (20, O.hwe_ry_ops(20, 5, rng.uniform(-np.pi, np.pi, 6 * 20))),
        (20, O.trotter_ops(*O.tfim_terms(20, 1.0, 0.7), 0.9, 4)),
        (20, O.qaoa_ring_ops(20, 4, rng.uniform(-np.pi, np.pi, 8))),
    ]
    try:
        for n, ops in cases:
            ref, _ = O.evolve_ops(n, ops)
            for td, tol, m, L in ((torch.complex128, TOL128, 11, 5), (torch.complex64, TOL64, 12, 6)):
                for lean, tile in ((1, TileConfig(m=m, L=L, threads=128)), (0, TileConfig(m=m, L=L, threads=128)),
                                   (1, TileConfig(m=m, L=L, threads=128, rot_layers=3))):
                    lib.tqb_set_tma(512 + lean)
                    eng = _engine(cuda_device, "b200", td, tile)
                    psi, _, _ = eng._evolve(FakeCircuit(n, ops), "state")
                    err = np.abs(psi.cpu().numpy() - ref).max()
                    assert err < tol, (n, ops[0], td, lean, tile.threads, err)
    finally:
        lib.tqb_set_tma(512 + 1)


def test_specialised_kernels_match_oracle(cuda_device):
    """The NVRTC-specialised pass kernels (csrc/tqb_spec.cuh through tqb_run_passes2, synchronous mode so that every
    eligible pass really runs its own kernel) against the oracle, complex128 and complex64, plain and padded layouts,
    outside-the-tile controls, batched states."""
    import torch
    from tyxonq_b200 import _lib
    from tyxonq_b200 import program as P
    from tyxonq_b200.fuse import fuse
    from tyxonq_b200.gates import lower_op
    from tyxonq_b200.planner import TileConfig, compile_program
    lib = _lib.load()
    rng = np.random.default_rng(78)
    cases = [
        (20, O.hea_ops(20, 8, rng.uniform(-np.pi, np.pi, 2 * 8 * 20))),
        (19, O.hwe_ry_ops(19, 5, rng.uniform(-np.pi, np.pi, 6 * 19))),
        (18, O.trotter_ops(*O.tfim_terms(18, 1.0, 0.7), 0.9, 4)),
        (20, O.qaoa_ring_ops(20, 4, rng.uniform(-np.pi, np.pi, 8))),
    ]
    old = lib.tqb_set_jit(2)
    try:
        for n, ops in cases:
            ref, _ = O.evolve_ops(n, ops)
            for td, tol, m, L in ((torch.complex128, TOL128, 11, 5), (torch.complex64, TOL64, 12, 6), (torch.complex64, TOL64, 11, 6)):
                for tensor_tma in (1, 0):   # tiles staged by one tensor copy (cp.async.bulk.tensor) / one bulk copy per run
                    lib.tqb_set_jit(512 + tensor_tma)
                    before = _lib.jit_stats()["spec_launches"]
                    eng = _engine(cuda_device, "b200", td, TileConfig(m=m, L=L, threads=128))
                    psi, _, _ = eng._evolve(FakeCircuit(n, ops), "state")
                    err = np.abs(psi.cpu().numpy() - ref).max()
                    assert err < tol, (n, ops[0], td, m, tensor_tma, err)
                    assert _lib.jit_stats()["spec_launches"] > before, "no specialised kernel ran"
        lib.tqb_set_jit(512 + 1)
        # batched states: the batch index is more tile-index bits
        n, ops = 14, O.qaoa_ring_ops(14, 3, rng.uniform(-np.pi, np.pi, 6))
        ref, _ = O.evolve_ops(n, ops)
        lg = fuse([g for g in (lower_op(o, n, mode="run") for o in ops) if g is not None])
        prog = compile_program(lg, n, TileConfig(m=11, L=5, threads=128), itemsize=16)
        st = P.new_state(n, batch=5, dtype=torch.complex128, device=cuda_device)
        before = _lib.jit_stats()["spec_launches"]
        P.DeviceProgram(prog, cuda_device, torch.complex128).run(st)
        assert _lib.jit_stats()["spec_launches"] - before == prog.n_passes
        assert np.abs(st.cpu().numpy() - ref[None, :]).max() < TOL128
        # one matrix set per batch member (config 5's batched ansatz): the kernel re-stages the matrices member by member;
        # both code shapes (unrolled / one body per gate class) for both dtypes
        from tyxonq_b200.batched import BatchedAnsatz
        nq, layers, B = 15, 3, 24
        params = rng.uniform(-np.pi, np.pi, (B, (layers + 1) * nq))
        refs = np.stack([O.evolve_ops(nq, O.hwe_ry_ops(nq, layers, params[b]))[0] for b in range(B)])
        for td, tol in ((torch.complex64, TOL64), (torch.complex128, TOL128)):
            for shape in (0, 1):
                lib.tqb_set_jit(1024 + shape)
                ba = BatchedAnsatz(nq, layers, B, device=cuda_device, dtype=td,
                                   tile=TileConfig(m=12 if td == torch.complex64 else 11, L=6 if td == torch.complex64 else 5, threads=128))
                before = _lib.jit_stats()["spec_launches"]
                out = ba.run(params).cpu().numpy()
                assert _lib.jit_stats()["spec_launches"] - before == ba.passes, "per-member passes must run specialised kernels"
                assert np.abs(out - refs).max() < tol, (td, shape, np.abs(out - refs).max())
                if shape == 1:   # later calls rewrite the matrix buffer of the resident program (batched.HweRyRefill)
                    params2 = rng.uniform(-3.0, 3.0, params.shape)
                    refs2 = np.stack([O.evolve_ops(nq, O.hwe_ry_ops(nq, layers, params2[b]))[0] for b in range(B)])
                    for pp, rr in ((params2, refs2), (params, refs), (params2, refs2)):
                        assert np.abs(ba.run(pp).cpu().numpy() - rr).max() < tol
                    assert ba.fast_calls == 3, ba.fast_calls
                    params3 = params2.copy()
                    params3[3, :] = np.pi - 1e-7    # a member outside the recipe's domain: that call is lowered generically
                    refs3 = np.stack([O.evolve_ops(nq, O.hwe_ry_ops(nq, layers, params3[b]))[0] for b in range(B)])
                    assert np.abs(ba.run(params3).cpu().numpy() - refs3).max() < tol * 10 and ba.fast_calls == 3
                    assert np.abs(ba.run(params2).cpu().numpy() - refs2).max() < tol and ba.fast_calls == 4
    finally:
        lib.tqb_set_jit(1024 + 2)
        lib.tqb_set_jit(512 + 1)
        lib.tqb_set_jit(old)


def test_large_state_invariants(cuda_device):
    """n = 30 (16 GiB complex128, the bench size): GHZ amplitudes, norm, reversibility -- size-independent properties."""
    import torch
    from tyxonq_b200 import program as P
    n = 30
    eng = _engine(cuda_device, "b200")
    ghz = [("h", 0)] + [("cx", q, q + 1) for q in range(n - 1)]
    psi, _, _ = eng._evolve(FakeCircuit(n, ghz), "state")
    assert abs(complex(psi[0].cpu()) - 2 ** -0.5) < 1e-12 and abs(complex(psi[-1].cpu()) - 2 ** -0.5) < 1e-12
    assert abs(float(P.norm2(psi).cpu()[0]) - 1.0) < 1e-12
    z = P.expect_z_bits(psi).cpu().numpy()[0]
    assert np.abs(z).max() < 1e-12
    del psi
    ops = O.hea_ops(n, 2, np.random.default_rng(3).uniform(-3, 3, 4 * n))
    inv = []
    for op in reversed(ops):
        if op[0] in ("rz", "rx"):
            inv.append((op[0], op[1], -op[2]))
        else:
            inv.append(op)  # h and cx are self-inverse
    psi, _, _ = eng._evolve(FakeCircuit(n, ops + inv), "state")
    assert abs(complex(psi[0].cpu()) - 1.0) < 1e-10
    assert abs(float(P.norm2(psi).cpu()[0]) - 1.0) < 1e-10


def test_batched_ansatz_expval_and_sampling(cuda_device):
    """Config 5 at test size: B parameter sets of the HWE-RY ansatz, Heisenberg Pauli sum, shots per state."""
    import torch
    from tyxonq_b200 import PauliSum
    from tyxonq_b200.batched import BatchedAnsatz
    n, L, B, shots = 12, 3, 6, 512
    rng = np.random.default_rng(7)
    params = rng.random((B, (L + 1) * n))
    terms, w = O.heisenberg_terms(n, [(i, i + 1) for i in range(n - 1)], hzz=1.0, hxx=0.5, hyy=0.5, hx=-0.3)
    ham = PauliSum.from_codes(terms, w)
    u = np.random.default_rng(99).random((B, shots))
    for dt, tol in ((torch.complex128, 1e-10), (torch.complex64, 2e-5)):
        ba = BatchedAnsatz(n, L, B, device=cuda_device, dtype=dt)
        st = ba.run(params).cpu().numpy()
        ev = ba.expvals(ham).cpu().numpy()
        idx = ba.sample(torch.from_numpy(u)).cpu().numpy()
        for b in range(B):
            ref, _ = O.evolve_ops(n, O.hwe_ry_ops(n, L, params[b]))
            assert np.abs(st[b] - ref).max() < tol
            assert abs(ev[b] - O.expect_pauli_sum(ref, terms, w)) < tol * 50
            assert np.array_equal(idx[b], O.sample_indices(O.probabilities(st[b]), u[b]))


def test_kqubit_unitary_above_four_qubits(cuda_device):
    """apply_kqubit_unitary has no size limit in the reference (statevector.py:71-129): k = 5, 6 on scattered qubits."""
    from scipy.stats import unitary_group
    from tyxonq_b200 import kernels as K
    rng = np.random.default_rng(41)
    n = 9
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi /= np.linalg.norm(psi)
    for k, qs in ((5, [7, 0, 3, 8, 2]), (6, [1, 5, 0, 4, 8, 6])):
        U = unitary_group.rvs(1 << k, random_state=int(rng.integers(1 << 30)))
        got = K.apply_kqubit_unitary(psi, U, qs, n)
        assert isinstance(got, np.ndarray)
        assert np.abs(got - O.apply_kq(psi, U, qs, n)).max() < TOL128


def test_batched_config5_full_size(cuda_device):
    """Config 5 at its own size: 20-qubit HWE-RY ansatz (4 layers), 8 parameter sets, the 57-term Heisenberg chain,
    8192 shots per state -- states, expectation values (tiled kernel, two layouts) and sampled indices against the oracle."""
    import torch
    from tyxonq_b200 import PauliSum
    from tyxonq_b200.batched import BatchedAnsatz
    n, L, B, shots = 20, 4, 8, 8192
    rng = np.random.default_rng(7)
    params = rng.random((B, (L + 1) * n))
    terms, w = O.heisenberg_terms(n, [(i, i + 1) for i in range(n - 1)], hzz=1.0, hxx=1.0, hyy=1.0)
    ham = PauliSum.from_codes(terms, w)
    assert ham.n_terms == 57
    u = np.random.default_rng(99).random((B, shots))
    ba = BatchedAnsatz(n, L, B, device=cuda_device, dtype=torch.complex64)
    st = ba.run(params).cpu().numpy()
    ev = ba.expvals(ham).cpu().numpy()
    idx = ba.sample(torch.from_numpy(u)).cpu().numpy()
    for b in range(B):
        ref, _ = O.evolve_ops(n, O.hwe_ry_ops(n, L, params[b]))
        assert np.abs(st[b] - ref).max() < 1e-5
        assert abs(ev[b] - O.expect_pauli_sum(ref, terms, w)) < 2e-4
        assert np.array_equal(idx[b], O.sample_indices(O.probabilities(st[b]), u[b]))


def test_pauli_sum_on_shards_of_one_state(cuda_device):
    """The per-rank call of ShardedState.expect_pauli_sum, on one GPU: the 2^g slices of a state are evaluated one
    after the other with global_base = rank << n_local (xmasks local, Z factors anywhere) and summed; then the
    world-size-1 ShardedState path itself (layout logic on CPU: tests/test_sharded_cpu.py)."""
    import torch
    from tyxonq_b200 import PauliSum
    from tyxonq_b200.sharded import ShardedState, lower_and_fuse, plan_sharded
    rng = np.random.default_rng(17)
    n, g = 12, 2
    n_local = n - g
    terms = [[int(c) for c in rng.integers(0, 4, n)] for _ in range(30)]
    for t in terms:                       # X / Y only on local qubits (qubit q = bit n-1-q: qubits >= g), Z / I on the rank bits
        for q in range(g):
            t[q] = 3 if t[q] in (2, 3) else 0
    w = rng.normal(size=30).tolist()
    st = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    st /= np.linalg.norm(st)
    full = PauliSum.from_codes(terms, w)
    sub = PauliSum(n_local, [(int(x), int(z), complex(c)) for gi, x in enumerate(full.group_x)
                             for z, c in zip(full.term_z[full.group_ptr[gi]:full.group_ptr[gi + 1]],
                                             full.term_coef[full.group_ptr[gi]:full.group_ptr[gi + 1]])])
    for dt, tol in ((torch.complex128, TOL128), (torch.complex64, 1e-5)):
        d = torch.from_numpy(st).to(cuda_device).to(dt)
        tot = 0j
        for r in range(1 << g):
            shard = d[r << n_local:(r + 1) << n_local].contiguous()
            tot += complex(sub.expectation(shard, global_base=r << n_local).cpu().numpy()[0])
        assert abs(tot.real - O.expect_pauli_sum(st, terms, w)) < tol and abs(tot.imag) < tol
    ops = O.hea_ops(n, 2, rng.uniform(-3, 3, 4 * n))
    ss = ShardedState(n, torch.complex128, cuda_device)
    ss.init_zero()
    ss.run(plan_sharded(lower_and_fuse(ops, n), n, ss.g))
    terms2 = [[int(c) for c in rng.integers(0, 4, n)] for _ in range(20)] + [[1] * n, [2] * n]
    w2 = rng.normal(size=22).tolist()
    ref, _ = O.evolve_ops(n, ops, mode="run")
    e = complex(ss.expect_pauli_sum(PauliSum.from_codes(terms2, w2)))
    assert abs(e.real - O.expect_pauli_sum(ref, terms2, w2)) < TOL128 and abs(e.imag) < TOL128
