"""Specialised pass kernels on the CPU tier: the library's generator (tqb_spec_source) + the kernel template's own
per-thread code (csrc/tqb_spec.cuh, compiled by g++ per pass shape) against the oracle, and an NVRTC compile of the
same shapes for sm_100a (needs no GPU)."""
from __future__ import annotations

import ctypes as C
import re

import numpy as np
import pytest

from oracle import sv_oracle as O
from tests.emu.spec_emu import run_program_spec_emulated, spec_header
from tyxonq_b200 import _lib
from tyxonq_b200.circuits import hea_ops, hwe_ry_ops, qaoa_ring_ops, tfim_terms, trotter_ops
from tyxonq_b200.fuse import fuse
from tyxonq_b200.gates import lower_op
from tyxonq_b200.planner import TileConfig, compile_program


def _workload(name: str, n: int, layers: int, rng):
    if name == "hea":
        return hea_ops(n, layers, rng.uniform(-np.pi, np.pi, 2 * layers * n))
    if name == "hwe":
        return hwe_ry_ops(n, layers, rng.uniform(-np.pi, np.pi, (layers + 1) * n))
    if name == "qaoa":
        return qaoa_ring_ops(n, layers, rng.uniform(-np.pi, np.pi, 2 * layers))
    terms, w = tfim_terms(n, 1.0, 0.7)
    return [op for op in trotter_ops(terms, w, 0.9, layers) if op[0] != "measure_z"]


def _compile(ops, n, m, L, itemsize):
    lg = fuse([g for g in (lower_op(o, n, mode="run") for o in ops) if g is not None])
    return compile_program(lg, n, TileConfig(m=m, L=L, threads=128), itemsize=itemsize)


@pytest.mark.parametrize("name,n,layers,m,L,dtype", [
    ("hea", 16, 4, 11, 5, np.complex128),
    ("hwe", 15, 3, 11, 5, np.complex128),
    ("qaoa", 15, 3, 11, 5, np.complex128),
    ("trotter", 15, 2, 11, 5, np.complex128),
    ("hea", 15, 3, 12, 6, np.complex64),
    ("qaoa", 14, 2, 11, 6, np.complex64),
])
def test_spec_passes_match_oracle(name, n, layers, m, L, dtype):
    rng = np.random.default_rng(n * 31 + layers)
    ops = _workload(name, n, layers, rng)
    ref, _ = O.evolve_ops(n, ops, mode="run")
    prog = _compile(ops, n, m, L, np.dtype(dtype).itemsize)
    psi0 = np.zeros(1 << n, dtype=dtype)
    psi0[0] = 1
    out, n_spec = run_program_spec_emulated(prog, psi0)
    assert n_spec == prog.n_passes, "every pass of these circuits is lean-eligible"
    assert np.abs(out - ref).max() < (1e-12 if dtype == np.complex128 else 3e-5)


@pytest.mark.parametrize("dtype,m,L", [(np.complex128, 11, 5), (np.complex64, 12, 6)])
def test_spec_scaled_layer_forms(dtype, m, L):
    """Angles at and next to pi: scaled rotation layers in the c form (compiled in: tqb_gate.off_b bits 8..11 -> GateC.inv)
    and in the t form with |t| up to gates.ROT_T_MAX."""
    from tyxonq_b200 import gates as G
    n = 14
    rng = np.random.default_rng(77)
    special = [np.pi, -np.pi, np.pi - 1e-9, np.pi - 1.0 / 600, np.pi - 1.0 / 400, 0.0, np.pi / 2, 3.0]
    th = rng.choice(special, size=2 * 3 * n)
    ops = hea_ops(n, 3, th)
    ref, _ = O.evolve_ops(n, ops, mode="run")
    prog = _compile(ops, n, m, L, np.dtype(dtype).itemsize)
    chains = prog.gates[(prog.gates["kind"] == 5) & (prog.gates["off_a"] >= 4)]
    forms = {int(g["off_b"]) >> 8 for g in chains}
    assert all(f & 8 for f in forms) and any(f & 7 for f in forms) and any(not (f & 7) for f in forms), forms
    code = 1 if dtype == np.complex128 else 0
    hdrs = "".join(spec_header(np.ascontiguousarray(prog.passes), i, np.ascontiguousarray(prog.gates), code) for i in range(prog.n_passes))
    rot = re.findall(r"^  \{6, .*\}, (-?\d+), \d+, \d+, \d+\},$", hdrs, flags=re.M)    # GateC.inv of the rotation-form chains
    assert rot and all(int(v) >= 0 for v in rot) and any(int(v) > 0 for v in rot), rot
    psi0 = np.zeros(1 << n, dtype=dtype)
    psi0[0] = 1
    out, n_spec = run_program_spec_emulated(prog, psi0)
    assert n_spec == prog.n_passes
    assert np.abs(out - ref).max() < (1e-12 if dtype == np.complex128 else 3e-5)
    assert G.ROT_T_MAX >= 512


def test_spec_sharded_base_and_batch():
    """global_base feeds outside-the-tile controls / table bits; the batch index is just more tile index bits."""
    n, g = 13, 2
    rng = np.random.default_rng(5)
    ops = _workload("hea", n + g, 2, rng)
    ref, _ = O.evolve_ops(n + g, ops, mode="run")
    # run the first pass-set on the shards of a state whose top g bits are rank bits: only gates local to the shard
    # may appear, so use a circuit on n qubits embedded at the LOW qubits' end instead: compare per-batch members
    ops_n = _workload("qaoa", n, 2, rng)
    refn, _ = O.evolve_ops(n, ops_n, mode="run")
    prog = _compile(ops_n, n, 11, 5, 16)
    psi0 = np.zeros((3, 1 << n), dtype=np.complex128)
    psi0[:, 0] = 1
    out, n_spec = run_program_spec_emulated(prog, psi0.reshape(-1), batch=3)
    assert n_spec == prog.n_passes
    assert np.abs(out.reshape(3, -1) - refn[None, :]).max() < 1e-12


@pytest.mark.parametrize("dtype,m,L", [(np.complex64, 12, 6), (np.complex128, 11, 5)])
def test_spec_per_member_matrices(dtype, m, L):
    """One matrix set per batch member (tqb_gate.mat_bstride != 0, the batched ansatz of config 5): the specialised
    kernel stages the member's matrices gate by gate and re-stages them when its tile range reaches the next member."""
    from tyxonq_b200.batched import hwe_ry_gates
    n, layers, B = 14, 2, 3
    rng = np.random.default_rng(11)
    params = rng.uniform(-np.pi, np.pi, (B, (layers + 1) * n))
    prog = compile_program(fuse(hwe_ry_gates(n, layers, params)), n, TileConfig(m=m, L=L, threads=128), batch_mats=B,
                           itemsize=np.dtype(dtype).itemsize)
    assert (prog.gates["mat_bstride"] != 0).any()
    psi0 = np.zeros((B, 1 << n), dtype=dtype)
    psi0[:, 0] = 1
    out, n_spec = run_program_spec_emulated(prog, psi0.reshape(-1), batch=B)
    assert n_spec == prog.n_passes
    for b in range(B):
        ref, _ = O.evolve_ops(n, hwe_ry_ops(n, layers, params[b]), mode="run")
        assert np.abs(out.reshape(B, -1)[b] - ref).max() < (1e-12 if dtype == np.complex128 else 3e-5), b


def test_batched_ansatz_refill_equals_generic_lowering():
    """batched.HweRyRefill (the matrix buffer of the batched HWE-RY program written directly from the parameters, recipe
    derived by probing the generic pipeline) against the generic lowering: equal buffers for angles all over (-3, 3), in
    place into a complex64 staging buffer, and a refusal when a scaled layer needs the c form."""
    from tyxonq_b200.batched import HweRyRefill, hwe_ry_gates
    from tyxonq_b200.planner import default_tile
    for n, layers, B, itemsize in ((14, 2, 5, 16), (20, 4, 16, 8)):
        tile = default_tile(n, itemsize, B)
        rng = np.random.default_rng(n)
        p0 = rng.random((B, (layers + 1) * n))
        prog = compile_program(fuse(hwe_ry_gates(n, layers, p0)), n, tile, batch_mats=B, itemsize=itemsize)
        rf = HweRyRefill(n, layers, prog, B, tile, itemsize)
        assert rf.ok
        buf = rf.mats(p0).astype(np.complex64 if itemsize == 8 else np.complex128)
        assert np.abs(rf.mats(p0) - prog.mats).max() < 1e-13
        for _ in range(3):
            p = rng.uniform(-3.0, 3.0, p0.shape)
            ref = compile_program(fuse(hwe_ry_gates(n, layers, p)), n, tile, batch_mats=B, itemsize=itemsize)
            assert np.array_equal(ref.gates, prog.gates) and np.array_equal(ref.passes, prog.passes), "the plan must not depend on the angles"
            assert np.abs(rf.mats(p) - ref.mats).max() < 1e-13
            assert rf.mats(p, out=buf) is buf
            assert np.abs(buf - ref.mats.astype(buf.dtype)).max() < (1e-6 if itemsize == 8 else 1e-13)
        p = rng.uniform(-3.0, 3.0, p0.shape)
        p[B - 1, :] = np.pi - 1e-7              # |tan(theta / 2)| > gates.ROT_T_MAX in every scaled layer of one member
        before = buf.copy()
        assert rf.mats(p, out=buf) is None and np.array_equal(before, buf), "a refused call must leave the buffer alone"


def test_generator_invariants():
    """Register bits and thread bits partition the tile; a warp-level sync is only used between gates that share the
    warp bits; 'no sync' only between gates with the same mapping; the first gate waits for the tile."""
    rng = np.random.default_rng(11)
    n = 20
    ops = _workload("hea", n, 6, rng)
    prog = _compile(ops, n, 11, 5, 16)
    passes, gates = np.ascontiguousarray(prog.passes), np.ascontiguousarray(prog.gates)
    seen = 0
    for pi in range(len(passes)):
        h = spec_header(passes, pi, gates, 1)
        assert h is not None
        rows = re.findall(r"^\s*\{(.*)\},\s*$", h, flags=re.M)
        m = int(re.search(r"M = (\d+)", h).group(1))
        rbits = int(re.search(r"RBITS = (\d+)", h).group(1))
        prev = None
        for r in rows:
            nums = [int(x) for x in re.findall(r"-?\d+", r)]
            sync = nums[8]
            rb, tb = nums[17:22][:rbits], nums[22:29]
            assert sorted(rb + tb) == list(range(m))
            if prev is None:
                assert sync == 0
            else:
                prb, ptb = prev
                if sync == -1:
                    assert sorted(rb) == sorted(prb) and tb == ptb
                elif sync == 1:
                    assert tb[5:] == ptb[5:]
                else:
                    assert sync == 2
            prev = (rb, tb)
            seen += 1
    assert seen >= len(prog.gates)


def test_nvrtc_compiles_for_sm100a():
    """NVRTC (no GPU needed) accepts the template for a padded and an unpadded shape; cubins carry sm_100 code."""
    lib = _lib.load()
    rng = np.random.default_rng(3)
    ops = _workload("hea", 18, 2, rng)
    prog = _compile(ops, 18, 11, 5, 16)
    passes, gates = np.ascontiguousarray(prog.passes), np.ascontiguousarray(prog.gates)
    sizes = []
    for pi in range(min(2, len(passes))):
        one = np.ascontiguousarray(passes[pi:pi + 1])
        rc = lib.tqb_spec_compile(one.ctypes.data, gates.ctypes.data, 1)
        assert rc > 0, lib.tqb_last_error().decode()
        sizes.append(rc)
    assert all(s > 10000 for s in sizes)
