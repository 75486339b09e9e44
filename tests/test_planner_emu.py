"""Host logic: gate lowering + pass planner + the tile kernel's index arithmetic, run through the
host emulator of tile_pass_kernel (tests/emu, same __host__ __device__ code as the GPU kernel)
and compared with the oracle.  CPU only."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import sv_oracle as O
from oracle import ucc_oracle as U
from tests.conftest import random_ops
from tests.emu.emu import run_program_emulated
from tyxonq_b200 import gates as G
from tyxonq_b200 import planner as P


def _lower(ops, n, mode="run", cache=None):
    lg = [G.lower_op(op, n, mode=mode, unitary_cache=cache) for op in ops]
    return [g for g in lg if g is not None]


@pytest.mark.parametrize("n,m,L,dtype", [
    (1, 1, 0, np.complex128), (2, 2, 1, np.complex64), (3, 3, 2, np.complex128), (6, 4, 2, np.complex128),
    (9, 5, 2, np.complex128), (9, 6, 3, np.complex64), (10, 10, 5, np.complex128), (11, 7, 0, np.complex64),
    (12, 8, 4, np.complex128), (12, 11, 5, np.complex128),
])
def test_random_circuits(n, m, L, dtype):
    rng = np.random.default_rng(n * 100 + m)
    ops = random_ops(rng, n, 120)
    ref, _ = O.evolve_ops(n, ops, mode="run")
    prog = P.compile_program(_lower(ops, n), n, P.TileConfig(m=m, L=L), itemsize=np.dtype(dtype).itemsize)
    psi0 = np.zeros(1 << n, dtype=dtype)
    psi0[0] = 1
    out = run_program_emulated(prog, psi0)
    tol = 1e-12 if dtype == np.complex128 else 2e-5
    assert np.abs(out - ref).max() < tol
    assert sorted(prog.order) == list(range(len(prog.order)))  # every gate scheduled exactly once


def test_golden_circuits(golden):
    for tag in ("rand5", "rand9", "rand12"):
        c = golden["circuits"][tag]
        n = c["n"]
        prog = P.compile_program(_lower(c["ops"], n, mode="state"), n, P.TileConfig(m=min(n, 7), L=3))
        psi0 = np.zeros(1 << n, dtype=np.complex128)
        psi0[0] = 1
        assert np.abs(run_program_emulated(prog, psi0) - golden[f"{tag}_state"]).max() < 1e-12


def test_dense_k3_k4_and_classifier(golden):
    psi = golden["k7_psi"]
    for k in (1, 2, 3, 4):
        g = G.classify_unitary(golden[f"k7_U{k}"], golden[f"k7_q{k}"].tolist(), 7)
        assert g.kind == G.DENSE
        for m, L in ((7, 3), (5, 2), (4, 0)):
            prog = P.compile_program([g], 7, P.TileConfig(m=m, L=L))
            assert np.abs(run_program_emulated(prog, psi) - golden[f"k7_out{k}"]).max() < 1e-13
    assert G.classify_unitary(O.gate_cx_4x4(), [0, 1], 3).kind == G.SWAP
    assert G.classify_unitary(O.gate_cry_4x4(0.4), [0, 1], 3).kind == G.PAIR
    assert G.classify_unitary(O.gate_cz_4x4(), [0, 1], 3).kind == G.DIAG
    assert G.classify_unitary(O.gate_iswap_4x4(), [0, 1], 3).kind == G.PAIR
    assert G.classify_unitary(O.gate_rxx(0.3), [0, 1], 3).kind == G.DENSE


def test_batched_matrices_and_global_base():
    rng = np.random.default_rng(5)
    n, B = 6, 3
    thetas = rng.uniform(-3, 3, B)
    st = rng.normal(size=(B, 1 << n)) + 1j * rng.normal(size=(B, 1 << n))
    g = G.dense_gate(np.stack([G.rx_mat(t) for t in thetas]), [2], n)
    prog = P.compile_program([g], n, P.TileConfig(m=4, L=2), batch_mats=B)
    out = run_program_emulated(prog, st.reshape(-1), batch=B).reshape(B, -1)
    for b in range(B):
        assert np.abs(out[b] - O.apply_1q(st[b], O.gate_rx(thetas[b]), 2, n)).max() < 1e-13
    # a sharded state: 2 local bits hold the low part of a 4-qubit register; diag gates see rank bits
    N, nl = 4, 2
    full = rng.normal(size=1 << N) + 1j * rng.normal(size=1 << N)
    ops = [("rz", 0, 0.7), ("rzz", 1, 3, -0.4), ("cz", 0, 2), ("h", 3), ("cx", 2, 3)]
    ref = full.copy()
    for op in ops:
        ref, _ = O.evolve_ops(N, [op], mode="state", initial=ref)
    out = np.empty_like(full)
    for rank in range(1 << (N - nl)):
        lg = []
        for op in ops:
            g = G.lower_op(op, N, mode="run")
            lg.append(g)
        prog = P.compile_program(lg, nl, P.TileConfig(m=2, L=1))
        shard = full[rank << nl:(rank + 1) << nl]
        out[rank << nl:(rank + 1) << nl] = run_program_emulated(prog, shard, global_base=rank << nl)
    assert np.abs(out - ref).max() < 1e-13


def test_ucc_pair_gates_with_parity():
    """UCC excitation = PAIR rotation selected by the JW parity mask (tyxonq_b200.ucc) vs the oracle."""
    from tyxonq_b200 import ucc
    n = 8
    rng = np.random.default_rng(3)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi /= np.linalg.norm(psi)
    for f_idx in [(3, 0), (0, 3), (6, 2), (4, 0, 3, 7), (1, 3, 2, 0), (0, 1, 2, 3), (7, 5, 2, 0)]:
        theta = float(rng.uniform(-2, 2))
        g = ucc.excitation_gate(f_idx, theta, n)
        for m, L in ((8, 3), (6, 2)):
            prog = P.compile_program([g], n, P.TileConfig(m=m, L=L))
            out = run_program_emulated(prog, psi)
            assert np.abs(out - U.evolve_excitation(psi, f_idx, theta, n)).max() < 1e-13, f_idx


@pytest.mark.parametrize("n,seed", [(5, 0), (8, 1), (10, 2)])
def test_gate_fusion_preserves_the_circuit(n, seed):
    from tyxonq_b200.fuse import fuse
    rng = np.random.default_rng(seed)
    for ops in (random_ops(rng, n, 200), O.hea_ops(n, 3, rng.uniform(-3, 3, 6 * n)), O.qaoa_ring_ops(n, 3, rng.uniform(-3, 3, 6)),
                O.trotter_ops(*O.tfim_terms(n), 1.0, 2), O.tfim_vqe_ops(n, 2, rng.normal(size=(4, n)))):
        ref, _ = O.evolve_ops(n, ops, mode="run")
        lg = _lower(ops, n)
        fg = fuse(lg)
        assert len(fg) <= len(lg)
        prog = P.compile_program(fg, n, P.TileConfig(m=min(n, 6), L=2))
        psi0 = np.zeros(1 << n, dtype=np.complex128)
        psi0[0] = 1
        assert np.abs(run_program_emulated(prog, psi0) - ref).max() < 1e-12


def test_planner_fuses_layers():
    n = 20
    ops = O.hea_ops(n, 6, np.random.default_rng(0).uniform(-3, 3, 2 * 6 * n))
    lg = _lower(ops, n)
    prog = P.compile_program(lg, n, P.TileConfig(m=11, L=5))
    assert prog.n_passes < len(lg) / 3  # many gates per state sweep
    assert sorted(prog.order) == list(range(len(lg)))


@pytest.mark.parametrize("seed", range(6))
def test_chains_grouped_after_scheduling_only_take_tile_local_targets(seed):
    """Diagonal 1-qubit gates ride along in any pass (no tile-local bit needed); when chains are grouped per pass such
    a gate must not become a chain layer unless its bit IS in the tile (regression: KeyError in compile_program)."""
    from tyxonq_b200.fuse import fuse
    rng = np.random.default_rng(100 + seed)
    n = 11
    ops = random_ops(rng, n, 160)
    ref, _ = O.evolve_ops(n, ops, mode="run")
    prog = P.compile_program(fuse(_lower(ops, n)), n, P.TileConfig(m=7, L=3))
    psi0 = np.zeros(1 << n, dtype=np.complex128)
    psi0[0] = 1
    assert np.abs(run_program_emulated(prog, psi0) - ref).max() < 1e-12
    kinds = {(int(g["kind"]), int(g["k"]), int(g["off_a"])) for g in prog.gates if int(g["kind"]) == 5}
    assert kinds  # the random circuits do produce chains


@pytest.mark.parametrize("dtype,m,L", [(np.complex128, 7, 3), (np.complex64, 8, 4), (np.complex128, 6, 1)])
def test_lean_passes_padded_layout(dtype, m, L):
    """HEA / hwe-ry / QAOA / Trotter circuits compile to lean-eligible passes only (max_dense_k = -1): the emulator
    runs them like tile_pass_lean_kernel does -- padded tile layout (pidx), chains decoded into RotDesc with padded
    strides, low-bit targets included -- and must reproduce the oracle."""
    from tyxonq_b200.fuse import fuse
    rng = np.random.default_rng(5)
    n = 12
    cases = [
        O.hea_ops(n, 5, rng.uniform(-np.pi, np.pi, 2 * 5 * n)),
        O.hwe_ry_ops(n, 4, rng.uniform(-np.pi, np.pi, 5 * n)),
        O.qaoa_ring_ops(n, 3, rng.uniform(-np.pi, np.pi, 6)),
        O.trotter_ops(*O.tfim_terms(n, 1.0, 0.8), 0.7, 3),
    ]
    for ops in cases:
        ref, _ = O.evolve_ops(n, ops)
        prog = P.compile_program(fuse(_lower([o for o in ops if o[0] != "measure_z"], n, mode="state")), n,
                                 P.TileConfig(m=m, L=L), itemsize=np.dtype(dtype).itemsize)
        assert (prog.passes["max_dense_k"] < 0).all(), "these circuits must be lean-eligible"
        psi0 = np.zeros(1 << n, dtype=dtype)
        psi0[0] = 1
        out = run_program_emulated(prog, psi0)
        assert np.abs(out - ref).max() < (1e-12 if dtype == np.complex128 else 3e-5)


def test_streamed_chunks_are_the_same_program():
    """compile_program_stream (what StatevectorEngine launches chunk by chunk) yields, chunk after chunk, exactly the
    passes, gate descriptors and matrices of compile_program -- only the offsets restart at 0 in every chunk."""
    from tests.conftest import random_ops
    from tyxonq_b200.fuse import fuse
    from tyxonq_b200.gates import lower_op
    from tyxonq_b200.planner import TileConfig, compile_program, compile_program_stream, default_tile
    rng = np.random.default_rng(21)
    cases = [(14, O.hea_ops(14, 6, rng.uniform(-3, 3, 12 * 14))), (12, O.qaoa_ring_ops(12, 5, rng.uniform(-3, 3, 10))),
             (10, random_ops(rng, 10, 250))]
    for n, ops in cases:
        lowered = [g for g in (lower_op(o, n, mode="run") for o in ops) if g is not None]
        for itemsize in (16, 8):
            for chain in (None, False):
                for tile in (default_tile(n, itemsize, 1), TileConfig(m=min(n, 6), L=2)):
                    full = compile_program(fuse(list(lowered)), n, tile, itemsize=itemsize, chain=chain)
                    for first, chunk in ((8, 32), (1, 1), (2, 3)):
                        ps_all, g_all, m_all, go, mo = [], [], [], 0, 0
                        for pr in compile_program_stream(fuse(list(lowered)), n, tile, itemsize=itemsize, chain=chain, first=first, chunk=chunk):
                            ps, ga = pr.passes.copy(), pr.gates.copy()
                            assert int(ps["gate_begin"][0]) == 0 and int(ps["mat_begin"][0]) == 0
                            ps["gate_begin"] += go; ps["mat_begin"] += mo; ga["mat_off"] += mo
                            ps_all.append(ps); g_all.append(ga); m_all.append(pr.mats)
                            go += pr.gates.size; mo += pr.mats.size
                            assert pr.tile == full.tile
                        assert np.concatenate(ps_all).tobytes() == full.passes.tobytes()
                        assert np.concatenate(g_all).tobytes() == full.gates.tobytes()
                        assert np.concatenate(m_all).tobytes() == full.mats.tobytes()


@pytest.mark.parametrize("kind,n", [("hea", 11), ("random", 9)])
def test_streamed_chunks_run_one_after_the_other_equal_the_oracle(kind, n):
    """Every chunk of compile_program_stream is a self-contained Program: the emulated kernel runs them in sequence on
    the same state (what StatevectorEngine._flush does on the device) and must land on the oracle's state."""
    from tyxonq_b200.fuse import fuse
    from tyxonq_b200.gates import lower_op
    from tyxonq_b200.planner import TileConfig, compile_program_stream
    rng = np.random.default_rng(33)
    ops = O.hea_ops(n, 5, rng.uniform(-3, 3, 10 * n)) if kind == "hea" else random_ops(rng, n, 200)
    lowered = [g for g in (lower_op(o, n, mode="run") for o in ops) if g is not None]
    state = np.zeros(1 << n, dtype=np.complex128)
    state[0] = 1
    chunks = 0
    for prog in compile_program_stream(fuse(lowered), n, TileConfig(m=6, L=2), first=2, chunk=3):
        state = run_program_emulated(prog, state)
        chunks += 1
    assert chunks >= 3
    ref, _ = O.evolve_ops(n, ops, mode="run")
    assert np.abs(state - ref).max() < 1e-12
